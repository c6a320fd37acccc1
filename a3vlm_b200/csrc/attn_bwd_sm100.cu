// FlashAttention-style backward on tcgen05 / TMEM, fed by TMA (sm_100a), head_dim 128.
//
// One CTA = 128 keys of one (batch, head), walking the 64-query tiles that can see them.  18 warps:
//
//   warps  0-15  workers: threads t, t+128, t+256, t+384 share key row t (TMEM lane t) and split the 64 query
//                columns of a tile into four 16-column parts.  P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - delta)
//                scale -> shared memory (bf16, SW128); dQ^T read-out -> fp32 staging tile; dK / dV epilogue
//   warp   16    MMA issuer (warp-uniform loop, one elected lane issues)
//   warp   17    TMA producer: K, V once; Q_i / dO_i through a three-stage ring (with two stages the L2 / HBM
//                latency of the next tile sat in the critical path of every tile); issues the dQ reduce and
//                recycles the staging tile
//
// Per tile of 64 queries (all five contractions on tcgen05, transposes are free through the UMMA majorness):
//     S^T  = K Q^T            A = K  (K-major)   B = Q   (K-major)    -> TMEM [  0, 64)
//     dP^T = V dO^T           A = V  (K-major)   B = dO  (K-major)    -> TMEM [ 64,128)
//     dV  += P^T  dO          A = P^T (K-major)  B = dO  (MN-major)   -> TMEM [256,384)
//     dK  += dS^T Q           A = dS^T (K-major) B = Q   (MN-major)   -> TMEM [384,512)
//     dQ^T = K^T dS           A = K  (MN-major)  B = dS^T (MN-major)  -> TMEM [128,192)
// dQ^T has head-dim on lanes and queries on columns; each thread transposes its row into a
// [64 queries][128 dims] fp32 staging tile in shared memory (conflict-free: a warp writes 32 consecutive
// floats per query; the tile aliases the P^T / dS^T buffers, which are dead once the gradient MMAs retired)
// and ONE bulk-tensor reduce-add (cp.reduce.async.bulk.tensor, fp32 add in L2) per tile accumulates it into
// the global dQ accumulator.  The roles talk through mbarriers only (no block-wide barrier in the loop).
// Software pipeline: as soon as the workers have turned S^T / dP^T of tile i into P^T / dS^T (shared memory,
// double-buffered), the issuer sends S^T / dP^T of tile i+1 down the tensor pipe FIRST and the gradient MMAs of
// tile i behind them, so the workers' element-wise phase of tile i+1 runs while the tensor pipe computes the
// gradients of tile i; dQ^T of tile i is read out one iteration later.  The tensor pipe (whose M 128 x N 64
// MMAs are bound by shared-memory operand bandwidth, ~2100 cycles per tile) is then the limiter.
//
// Replaces the autograd backward of flash_attn_func / SDPA (reference
// model/accessory/model/LLM/llama_ens5.py:142-167) on the fine-tune path
// (model/accessory/engine_finetune.py:60-68).
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace a3 {

constexpr int BW_KEYS = 128;
constexpr int BW_Q = 64;
constexpr int BW_HD = 128;
constexpr int BW_STAGES = 3;                 // Q / dO ring
constexpr int BW_WORKERS = 512;
constexpr int BW_THREADS = BW_WORKERS + 64;  // + issuer warp + producer warp

struct BwdSmem {
  static constexpr int KV_BYTES = BW_KEYS * BW_HD * 2;  // 32 KB
  static constexpr int QT_BYTES = BW_Q * BW_HD * 2;     // 16 KB
  static constexpr int PT_BYTES = BW_KEYS * BW_Q * 2;   // 16 KB
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_Q = OFF_V + KV_BYTES;                 // [BW_STAGES]
  static constexpr int OFF_DO = OFF_Q + BW_STAGES * QT_BYTES;    // [BW_STAGES]
  // P^T | dS^T of a tile (16 + 16 KB), double-buffered by tile parity; once the tile's gradient MMAs have retired the
  // same 32 KB become the fp32 [64 q][128 d] staging tile of its dQ reduce
  static constexpr int PDS_BYTES = 2 * PT_BYTES;
  static constexpr int OFF_PDS = OFF_DO + BW_STAGES * QT_BYTES;  // [2]
  static constexpr int OFF_STAT = OFF_PDS + 2 * PDS_BYTES;       // lse[2][64], delta[2][64] fp32
  static constexpr int OFF_BAR = OFF_STAT + 4 * BW_Q * 4;
  static constexpr int TOTAL = OFF_BAR + 256;  // the dynamic smem base is 1024-byte aligned (probed by the host)
};
static_assert(BW_Q * BW_HD * 4 == BwdSmem::PDS_BYTES, "dQ staging must fit the P^T | dS^T buffers");

struct BwdParams {
  const float* lse;    // [B][H][T]
  const float* delta;  // [B][H][T]
  float* dq_acc;       // [B*T][H*128] fp32
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  const float2* rope;  // [T][64] (cos, sin) or null: fold the backward of apply_rotary_emb into dq / dk
  long long lddkv;
  int B, H, T;
  float scale;
  int causal;
  int debug_skip;  // profiling builds only (always 0 in the library): 1 = no dQ reduce
  long long* trace;  // debug (A3_BW_TRACE_PTR): clock64 stamps [cta][role: worker, issuer, producer][128]
};

__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const BwdParams p) {
  using S = BwdSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) return;  // SW128 tiles need a 1024-byte aligned base (host probe: A3_ERR_UNSUPPORTED)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* kv_full = bars;                // [1] producer -> issuer
  uint64_t* q_full = bars + 1;             // [4] producer -> issuer: Q_i, dO_i landed
  uint64_t* q_empty = bars + 5;            // [4] issuer (commit) -> producer: every MMA reading the stage retired
  uint64_t* sdp_bar = bars + 9;            // issuer (commit) -> workers: S^T / dP^T of the tile ready
  uint64_t* grad_bar = bars + 10;          // issuer (commit) -> workers: dV / dK / dQ^T MMAs of the tile retired
  uint64_t* ps_ready = bars + 11;          // workers (16 warps) -> issuer: P^T / dS^T of the tile in shared memory
  uint64_t* dq_staged = bars + 12;         // workers (16 warps) -> issuer + producer: dQ^T read out of TMEM, staged
  uint64_t* dq_free = bars + 13;           // producer -> workers: the reduce has read the staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* s_lse = reinterpret_cast<float*>(smem + S::OFF_STAT);   // [2][64]
  float* s_delta = s_lse + 2 * BW_Q;                             // [2][64]

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int key0 = blockIdx.x * BW_KEYS, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  const int n_q_tiles = (T + BW_Q - 1) / BW_Q;
  const int i_start = p.causal ? key0 / BW_Q : 0;  // first query tile that can see these keys
  const int n_iter = n_q_tiles - i_start;

  if (tid == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
    mbar_init(kv_full, 1);
    for (int i = 0; i < BW_STAGES; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(sdp_bar, 1);
    mbar_init(grad_bar, 1);
    mbar_init(ps_ready, BW_WORKERS / 32);
    mbar_init(dq_staged, BW_WORKERS / 32);
    mbar_init(dq_free, 1);
    fence_mbar_init();
  }
  if (warp_id == 16) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t C_S = 0, C_DP = 64, C_DQ = 128, C_DV = 256, C_DK = 384;
  const long long stat_base = (static_cast<long long>(b) * p.H + h) * T;
  long long* trace = p.trace ? p.trace + ((static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 384
                             : nullptr;
  int tr_n = 0;
  auto TR = [&](int role) {
    if (trace && tr_n < 128) trace[role * 128 + tr_n] = clock64();
    ++tr_n;
  };

  if (warp_id == 17) {
    // =========================================================== TMA producer + dQ reduce
    if (n_iter > 0) {  // warp-uniform; an elected lane issues (no waterfall loop around the TMA instructions)
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t b_qfull = __shfl_sync(0xffffffffu, smem_u32(q_full), 0);
      const uint32_t b_kvfull = __shfl_sync(0xffffffffu, smem_u32(kv_full), 0);
      auto load_q = [&](int it) {  // Q_i and dO_i -> stage it % BW_STAGES
        const int st = it % BW_STAGES, i = i_start + it;
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[st], 2 * S::QT_BYTES);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tma_load_3d_s(sb + S::OFF_Q + st * S::QT_BYTES + c * (BW_Q * 128), &tmQ, b_qfull + st * 8,
                          h * BW_HD + c * 64, i * BW_Q, b);
            tma_load_3d_s(sb + S::OFF_DO + st * S::QT_BYTES + c * (BW_Q * 128), &tmDO, b_qfull + st * 8,
                          h * BW_HD + c * 64, i * BW_Q, b);
          }
        }
        __syncwarp();
      };
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * S::KV_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d_s(sb + S::OFF_K + c * (BW_KEYS * 128), &tmK, b_kvfull, h * BW_HD + c * 64, key0, b);
          tma_load_3d_s(sb + S::OFF_V + c * (BW_KEYS * 128), &tmV, b_kvfull, h * BW_HD + c * 64, key0, b);
        }
      }
      __syncwarp();
      for (int it = 0; it < min(n_iter, BW_STAGES); ++it) load_q(it);
      for (int it = 0; it < n_iter; ++it) {
        if (it + BW_STAGES < n_iter) {  // refill the stage of tile `it` once its gradient MMAs have retired
          mbar_wait(&q_empty[it % BW_STAGES], (it / BW_STAGES) & 1);
          load_q(it + BW_STAGES);
        }
        TR(2);  // refill issued
        mbar_wait(dq_staged, it & 1);
        TR(2);  // staged seen
        // the reduce, its bulk group and the release belong to ONE thread (bulk groups are per thread); elect.sync
        // picks the same lane for the same member mask every time
        if (elect_one()) {
          if (!(p.debug_skip & 1)) {
            tma_reduce_add_3d_s(&tmDQ, sb + S::OFF_PDS + (it & 1) * S::PDS_BYTES, h * BW_HD, (i_start + it) * BW_Q,
                                b);  // rows >= T are clipped
            bulk_commit_group();
            bulk_wait_group_read0();  // the staging tile (= P^T | dS^T buffers) may be overwritten
          }
          mbar_arrive(dq_free);
        }
        __syncwarp();
        TR(2);  // reduce read done
      }
      if (elect_one()) bulk_wait_group0();
      __syncwarp();
    }
  } else if (warp_id == 16) {
    // =========================================================== MMA issuer
    if (n_iter > 0) {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t sK = sb + S::OFF_K, sV = sb + S::OFF_V, sQ = sb + S::OFF_Q, sDO = sb + S::OFF_DO,
                     sPDS = sb + S::OFF_PDS;
      constexpr uint32_t idesc_sdp = make_idesc_bf16(128, BW_Q, 0, 0);   // [keys x q]
      constexpr uint32_t idesc_dkv = make_idesc_bf16(128, BW_HD, 0, 1);  // [keys x hd], B MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, BW_Q, 1, 1);    // [hd x q], both MN-major
      auto issue_sdp = [&](int it) {
        const int st = it % BW_STAGES;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < BW_HD / 16; ++kk) {
            const uint32_t ch = kk >> 2, inner = (kk & 3) * 32;
            const uint64_t ak = make_smem_desc_sw128(sK + ch * (BW_KEYS * 128) + inner, 16, 1024);
            const uint64_t bq = make_smem_desc_sw128(sQ + st * S::QT_BYTES + ch * (BW_Q * 128) + inner, 16, 1024);
            umma_bf16(tm + C_S, ak, bq, idesc_sdp, kk > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < BW_HD / 16; ++kk) {
            const uint32_t ch = kk >> 2, inner = (kk & 3) * 32;
            const uint64_t av = make_smem_desc_sw128(sV + ch * (BW_KEYS * 128) + inner, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(sDO + st * S::QT_BYTES + ch * (BW_Q * 128) + inner, 16, 1024);
            umma_bf16(tm + C_DP, av, bd, idesc_sdp, kk > 0 ? 1u : 0u);
          }
          umma_commit(sdp_bar);
        }
        __syncwarp();
      };
      auto issue_grads = [&](int it) {
        const int st = it % BW_STAGES;
        const uint32_t sP = sPDS + (it & 1) * S::PDS_BYTES, sDS = sP + S::PT_BYTES;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < BW_Q / 16; ++kk) {  // reduce over the 64 queries
            const uint64_t ap = make_smem_desc_sw128(sP + kk * 32, 16, 1024);
            const uint64_t ads = make_smem_desc_sw128(sDS + kk * 32, 16, 1024);
            const uint64_t bdo = make_smem_desc_sw128(sDO + st * S::QT_BYTES + kk * (16 * 128), BW_Q * 128, 1024);
            const uint64_t bq = make_smem_desc_sw128(sQ + st * S::QT_BYTES + kk * (16 * 128), BW_Q * 128, 1024);
            const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
            umma_bf16(tm + C_DV, ap, bdo, idesc_dkv, acc);
            umma_bf16(tm + C_DK, ads, bq, idesc_dkv, acc);
          }
#pragma unroll
          for (int kk = 0; kk < BW_KEYS / 16; ++kk) {  // reduce over the 128 keys
            const uint64_t ak = make_smem_desc_sw128(sK + kk * (16 * 128), BW_KEYS * 128, 1024);
            const uint64_t bds = make_smem_desc_sw128(sDS + kk * (16 * 128), BW_KEYS * 128, 1024);
            umma_bf16(tm + C_DQ, ak, bds, idesc_dq, kk > 0 ? 1u : 0u);
          }
          umma_commit(grad_bar);
          umma_commit(&q_empty[st]);  // Q_i / dO_i (and P^T / dS^T) are no longer read
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(ps_ready, it & 1);  // P^T / dS^T of tile it are in shared memory, S^T / dP^T TMEM is consumed
        tc_fence_after();
        if (lane == 0) TR(1);  // ps_ready seen
        if (it + 1 < n_iter) {  // next tile's scores first: the workers' element-wise phase overlaps the gradients
          mbar_wait(&q_full[(it + 1) % BW_STAGES], ((it + 1) / BW_STAGES) & 1);
          tc_fence_after();
          if (lane == 0) TR(1);  // q landed
          issue_sdp(it + 1);
          if (lane == 0) TR(1);  // sdp issued
        }
        if (it > 0) mbar_wait(dq_staged, (it - 1) & 1);  // dQ^T of tile it-1 has been read out of TMEM
        tc_fence_after();
        if (lane == 0) TR(1);  // dq_staged seen
        issue_grads(it);
        if (lane == 0) TR(1);  // grads issued
      }
    }
  } else {
    // =========================================================== workers
    const int warp = warp_id & 3;   // TMEM lane quarter
    const int part = tid >> 7;      // which 16 of the 64 query columns this thread handles
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    auto load_stats = [&](int it) {  // lse / delta of query tile it -> smem stage it&1
      const int st = it & 1;
      const int qi = (i_start + it) * BW_Q + (tid & 63);
      // lse is stored pre-multiplied by log2(e): P = exp2(S * scale * log2e - lse * log2e)
      if (tid < 64) s_lse[st * BW_Q + tid] = qi < T ? p.lse[stat_base + qi] * 1.4426950408889634f : INFINITY;
      else if (tid < 128) s_delta[st * BW_Q + (tid - 64)] = qi < T ? p.delta[stat_base + qi] : 0.f;
    };
    if (n_iter > 0) load_stats(0);
    named_bar_sync(1, BW_WORKERS);

    const int key = key0 + warp * 32 + lane;
    const bool key_ok = key < T;
    const float c2 = p.scale * 1.4426950408889634f;
    const int sw = lane & 7;
    const int row_off = (warp * 32 + lane) * 128;
    // dQ^T of tile `jt`: TMEM -> fp32 staging (the tile's own P^T | dS^T buffer, dead once its gradient MMAs retired)
    auto stage_dq = [&](int jt) {
      mbar_wait(grad_bar, jt & 1);
      tc_fence_after();
      if (tid == 0) TR(0);  // grads done
      float* stage = reinterpret_cast<float*>(smem + S::OFF_PDS + (jt & 1) * S::PDS_BYTES) + warp * 32 + lane;  // column d
      uint32_t dq[16];
      tmem_ld_32x16(t_lane + C_DQ + part * 16, dq);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) stage[(part * 16 + c) * BW_HD] = __uint_as_float(dq[c]);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_staged);
      if (tid == 0) TR(0);  // dQ staged
    };

    for (int it = 0; it < n_iter; ++it) {
      const int i = i_start + it;
      const int st = it & 1;
      // prefetch the statistics of the next tile (consumed after this iteration's worker barrier)
      if (it + 1 < n_iter) load_stats(it + 1);
      mbar_wait(sdp_bar, it & 1);
      tc_fence_after();
      if (tid == 0) TR(0);  // S / dP ready
      const int qbase = i * BW_Q;
      uint4 pw[2], dw[2];
      {
        uint32_t s[16], dp[16];
        tmem_ld_32x16(t_lane + C_S + part * 16, s);
        tmem_ld_32x16(t_lane + C_DP + part * 16, dp);
        const float4* l4 = reinterpret_cast<const float4*>(s_lse + st * BW_Q + part * 16);
        const float4* d4 = reinterpret_cast<const float4*>(s_delta + st * BW_Q + part * 16);
        // masking is only needed on tiles that touch the diagonal or the end of the sequence (block-uniform)
        const bool edge = (p.causal && qbase < key0 + BW_KEYS) || key0 + BW_KEYS > T;
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 la = l4[2 * u], lb = l4[2 * u + 1], da = d4[2 * u], db = d4[2 * u + 1];
          const float l2[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
          const float dl[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
          float pv[8], dv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int idx = u * 8 + e;
            float pr = ex2_approx(fmaf(__uint_as_float(s[idx]), c2, -l2[e]));  // lse = +inf beyond T -> 0
            if (edge) {
              const int qi = qbase + part * 16 + idx;
              if (!(key_ok && (!p.causal || key <= qi))) pr = 0.f;
            }
            pv[e] = pr;
            dv[e] = pr * (__uint_as_float(dp[idx]) - dl[e]) * p.scale;
          }
          pw[u] = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                             pack_bf16x2(pv[6], pv[7]));
          dw[u] = make_uint4(pack_bf16x2(dv[0], dv[1]), pack_bf16x2(dv[2], dv[3]), pack_bf16x2(dv[4], dv[5]),
                             pack_bf16x2(dv[6], dv[7]));
        }
      }
      if (tid == 0) TR(0);  // element-wise done
      // this buffer last held the staging tile of tile it-2: its reduce must have read it
      if (it > 1) mbar_wait(dq_free, (it - 2) & 1);
      {
        uint8_t* p_row = smem + S::OFF_PDS + st * S::PDS_BYTES + row_off;
        uint8_t* ds_row = p_row + S::PT_BYTES;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int unit = part * 2 + u;
          *reinterpret_cast<uint4*>(p_row + ((unit ^ sw) << 4)) = pw[u];
          *reinterpret_cast<uint4*>(ds_row + ((unit ^ sw) << 4)) = dw[u];
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ps_ready);
      named_bar_sync(1, BW_WORKERS);  // statistics of tile it+1 are visible; everyone is done with those of tile it
      if (tid == 0) TR(0);  // P / dS published + worker barrier
      // dQ^T of the PREVIOUS tile (its gradient MMAs ran underneath this tile's element-wise phase)
      if (it > 0) stage_dq(it - 1);
    }
    if (n_iter > 0) stage_dq(n_iter - 1);

    // ---- epilogue: dV, dK -> bf16
    if (n_iter > 0) {
      // grad_bar of the last iteration has been waited on above by every worker
      __nv_bfloat16* dvr = p.dv + (static_cast<long long>(b) * T + key) * p.lddkv + h * BW_HD;
      __nv_bfloat16* dkr = p.dk + (static_cast<long long>(b) * T + key) * p.lddkv + h * BW_HD;
      {
        const int c = part;  // one 32-column chunk of dV and dK per thread
        uint32_t a[32], d[32];
        tmem_ld_32x32(t_lane + C_DV + c * 32, a);
        tmem_ld_32x32(t_lane + C_DK + c * 32, d);
        tmem_ld_wait();
        if (key_ok) {
          if (p.rope) {  // dK was taken w.r.t. the rotated keys: rotate back by -angle(key position)
            const float2* tp = p.rope + (long long)key * (BW_HD / 2) + c * 16;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 cs = __ldg(tp + (i >> 1));
              const float x = __uint_as_float(d[i]), y = __uint_as_float(d[i + 1]);
              d[i] = __float_as_uint(x * cs.x + y * cs.y);
              d[i + 1] = __float_as_uint(y * cs.x - x * cs.y);
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
            q.y = pack_bf16x2(__uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]));
            q.z = pack_bf16x2(__uint_as_float(a[i + 4]), __uint_as_float(a[i + 5]));
            q.w = pack_bf16x2(__uint_as_float(a[i + 6]), __uint_as_float(a[i + 7]));
            *reinterpret_cast<uint4*>(dvr + c * 32 + i) = q;
            q.x = pack_bf16x2(__uint_as_float(d[i]), __uint_as_float(d[i + 1]));
            q.y = pack_bf16x2(__uint_as_float(d[i + 2]), __uint_as_float(d[i + 3]));
            q.z = pack_bf16x2(__uint_as_float(d[i + 4]), __uint_as_float(d[i + 5]));
            q.w = pack_bf16x2(__uint_as_float(d[i + 6]), __uint_as_float(d[i + 7]));
            *reinterpret_cast<uint4*>(dkr + c * 32 + i) = q;
          }
        }
      }
    } else if (key_ok && part == 0) {
      // keys no query can see (cannot happen for causal Tq == Tk, kept for safety): zero gradient
      __nv_bfloat16* dvr = p.dv + (static_cast<long long>(b) * T + key) * p.lddkv + h * BW_HD;
      __nv_bfloat16* dkr = p.dk + (static_cast<long long>(b) * T + key) * p.lddkv + h * BW_HD;
      for (int c = 0; c < BW_HD / 8; ++c) {
        reinterpret_cast<uint4*>(dvr)[c] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(dkr)[c] = make_uint4(0, 0, 0, 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b][h][t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   (one warp per (b,t,h))
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                  const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                                  int B, int H, int T, long long ldo) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * T * H) return;
  const int hh = (int)(w % H);
  const long long bt = w / H;
  const long long off = bt * ldo + hh * BW_HD + lane * 4;
  const uint2 a = *reinterpret_cast<const uint2*>(o + off);
  const uint2 d = *reinterpret_cast<const uint2*>(dout + off);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y);
  const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y);
  float s = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) {
    const int bb = (int)(bt / T), t = (int)(bt % T);
    delta[((long long)bb * H + hh) * T + t] = s;
  }
}

// dq (bf16, strided) = dq_acc (fp32, dense [B*T][H*128])
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq,
                                       long long rows, int width, long long lddq,
                                       const float2* __restrict__ rope, int T) {
  const int cpr = width / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * cpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int c = (int)(i - r * cpr);
    float4 a = *reinterpret_cast<const float4*>(acc + r * width + c * 8);
    float4 b = *reinterpret_cast<const float4*>(acc + r * width + c * 8 + 4);
    if (rope) {  // inverse rotation at position r % T, pairs ((c*8) % 128) / 2 ...
      const float2* tp = rope + (r % T) * (BW_HD / 2) + (((c * 8) & (BW_HD - 1)) >> 1);
      const float2 c0 = __ldg(tp), c1 = __ldg(tp + 1), c2 = __ldg(tp + 2), c3 = __ldg(tp + 3);
      float t;
      t = a.x * c0.x + a.y * c0.y; a.y = a.y * c0.x - a.x * c0.y; a.x = t;
      t = a.z * c1.x + a.w * c1.y; a.w = a.w * c1.x - a.z * c1.y; a.z = t;
      t = b.x * c2.x + b.y * c2.y; b.y = b.y * c2.x - b.x * c2.y; b.x = t;
      t = b.z * c3.x + b.w * c3.y; b.w = b.w * c3.x - b.z * c3.y; b.z = t;
    }
    uint4 q;
    q.x = pack_bf16x2(a.x, a.y);
    q.y = pack_bf16x2(a.z, a.w);
    q.z = pack_bf16x2(b.x, b.y);
    q.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dq + r * lddq + c * 8) = q;
  }
}

}  // namespace a3

extern "C" int a3_attn_bwd(const void* q, const void* k, const void* v, const void* o,
                           const void* dout, const float* lse, float* delta, float* dq_acc, void* dq,
                           void* dk, void* dv, int B, int H, int T, int head_dim, long long ldqkv,
                           long long ldo, long long lddqkv, float scale, int causal,
                           const float* rope_table, void* stream) {
  using namespace a3;
  if (B <= 0 || H <= 0 || T <= 0) return A3_OK;
  if (head_dim != BW_HD) return set_error(A3_ERR_UNSUPPORTED, "a3_attn_bwd: head_dim must be 128");
  if ((ldqkv & 7) || (ldo & 7) || (lddqkv & 7))
    return set_error(A3_ERR_ARG, "a3_attn_bwd: strides must be multiples of 8 elements");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int width = H * BW_HD;
  {
    const long long warps = (long long)B * T * H;
    const long long blocks = (warps * 32 + 255) / 256;
    attn_delta_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                       reinterpret_cast<const __nv_bfloat16*>(dout),
                                                       delta, B, H, T, ldo);
    int rc = check_launch("a3_attn_bwd(delta)");
    if (rc) return rc;
  }
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * T * width * sizeof(float), s);
  if (e != cudaSuccess) return set_error(A3_ERR_CUDA, cudaGetErrorString(e));

  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ;
  int rc = make_tmap_3d_bf16(&tmQ, q, (uint64_t)width, (uint64_t)T, (uint64_t)B, (uint64_t)ldqkv * 2,
                             (uint64_t)T * ldqkv * 2, 64, BW_Q, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmK, k, (uint64_t)width, (uint64_t)T, (uint64_t)B, (uint64_t)ldqkv * 2,
                         (uint64_t)T * ldqkv * 2, 64, BW_KEYS, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmV, v, (uint64_t)width, (uint64_t)T, (uint64_t)B, (uint64_t)ldqkv * 2,
                         (uint64_t)T * ldqkv * 2, 64, BW_KEYS, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmDO, dout, (uint64_t)width, (uint64_t)T, (uint64_t)B, (uint64_t)ldo * 2,
                         (uint64_t)T * ldo * 2, 64, BW_Q, 1);
  if (rc) return rc;

  rc = make_tmap_3d_f32_linear(&tmDQ, dq_acc, (uint64_t)width, (uint64_t)T, (uint64_t)B,
                               (uint64_t)width * 4, (uint64_t)T * width * 4, BW_HD, BW_Q, 1);
  if (rc) return rc;

  rc = check_dyn_smem_alignment("a3_attn_bwd");
  if (rc) return rc;
  rc = ensure_dyn_smem(reinterpret_cast<const void*>(&attn_bwd_kernel), BwdSmem::TOTAL);
  if (rc) return rc;
  BwdParams p;
  p.lse = lse;
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddkv = lddqkv;
  p.rope = reinterpret_cast<const float2*>(rope_table);
  p.B = B; p.H = H; p.T = T;
  p.scale = scale;
  p.causal = causal ? 1 : 0;
  p.debug_skip = 0;
  static long long* const trace_ptr =
      getenv("A3_BW_TRACE_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("A3_BW_TRACE_PTR"), nullptr, 0)) : nullptr;
  p.trace = trace_ptr;
  dim3 grid((T + BW_KEYS - 1) / BW_KEYS, H, B);
  attn_bwd_kernel<<<grid, BW_THREADS, BwdSmem::TOTAL, s>>>(tmQ, tmK, tmV, tmDO, tmDQ, p);
  rc = check_launch("a3_attn_bwd");
  if (rc) return rc;
  const long long rows = (long long)B * T;
  long long blocks = (rows * (width / 8) + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  attn_dq_convert_kernel<<<(unsigned)blocks, 256, 0, s>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(dq),
                                                          rows, width, lddqkv,
                                                          reinterpret_cast<const float2*>(rope_table), T);
  return check_launch("a3_attn_bwd(dq)");
}

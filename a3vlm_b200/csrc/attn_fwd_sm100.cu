// Warp-specialised, persistent FlashAttention forward on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// One CTA per SM (512 threads) walks a list of work items; an item = TWO 128-query tiles ("stages") of one
// (batch, head) that share the same 128-key tiles:
//
//   warps  0-3   softmax group 0 (query tile 0): thread t owns query row t = TMEM lane t
//   warps  4-7   softmax group 1 (query tile 1)
//   warps  8-11  epilogue group: O_s / l -> bf16 -> global, lse, hands O_s back for the next item
//   warp   12    MMA issuer of stage 0: S_0 = Q_0 K_j^T and O_0 += P_0 V_j on tcgen05 (one elected lane)
//   warp   14    MMA issuer of stage 1
//   warp   13    K/V producer (TMA): K_j / V_j through a 4-slot ring shared by both stages
//   warp   15    Q producer (TMA), TMEM allocator
//
// Tensor memory (512 columns): S_0 [0,128), S_1 [128,256), O_0 [256,256+HD), O_1 [256+HD,256+2HD).
// P (bf16, two keys per 32-bit column) overwrites the first 64 columns of its S and is fed to the
// P V product as the TMEM A-operand of tcgen05.mma -- P never touches shared memory (the round-1
// kernel wrote it through smem with 1.07 M bank conflicts per launch).  While one softmax group
// exponentiates tile j of its queries the tensor core works for the other stage, so MUFU and the tensor
// pipe overlap; P is published in two halves so the first four k-steps of P V start while the second half
// is still being exponentiated.  The two stages are independent instruction streams (own issuer warp, own
// barriers) coupled only through the K/V ring, so a stage that has finished an item (a causal pair is n
// and n + 1 key tiles long) starts the next one at once, and the epilogue of an item overlaps the first
// tiles of the next.  Items are ordered heaviest first and dealt round-robin to the persistent CTAs.
// Rescaling is lazy: the running maximum a row normalises by only moves when the true maximum grew by
// more than 2^8 (exact: l and O use the same stale maximum), which takes the O read-modify-write off
// the steady state; when it is needed the softmax thread rescales its own O row.
// Registers are re-balanced with setmaxnreg (softmax 184 / epilogue 72 / issuers + producers 72 per thread:
// a whole 128-key row of scores lives in registers).
//
// Replaces flash_attn_func / F.scaled_dot_product_attention in reference
// model/accessory/model/LLM/llama_ens5.py:142-167 (causal, right-aligned when Tk > Tq, head_dim 128)
// and the attention core of open_clip's ResidualAttentionBlock (llama_ens5.py:366; no mask, head_dim 64).
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace a3 {

int attn_fwd_legacy(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq,
                    int Tk, int head_dim, long long ldq, long long ldk, long long ldv, long long ldo,
                    long long q_bstride, long long kv_bstride, float scale, int causal, void* stream);

constexpr int FA_BM = 128;       // queries per tile (two tiles per CTA)
constexpr int FA_BN = 128;       // keys per tile
constexpr int FA_THREADS = 512;
constexpr float FA_RESCALE_LOG2 = 8.f;
#ifndef FA_REGS_SOFTMAX  // per-thread register budgets after setmaxnreg (2 x 128 + 128 + 128 threads, 64 K registers)
#define FA_REGS_SOFTMAX 184
#define FA_REGS_EPILOGUE 72
#define FA_REGS_CONTROL 72
#endif
static_assert(2 * FA_REGS_SOFTMAX + FA_REGS_EPILOGUE + FA_REGS_CONTROL == 512, "register budget");


template <int HD>
struct FaSmem {
  static constexpr int Q_BYTES = FA_BM * HD * 2;
  static constexpr int KV_BYTES = FA_BN * HD * 2;
  // K/V ring (K_j, V_j alternate): five 32 KB slots at head_dim 128 -- everything the 227 KB of an SM hold next
  // to the two Q tiles -- so the producer stays ~1.5 key tiles ahead of the slower stage
  static constexpr int KV_SLOTS = HD == 128 ? 5 : 8;
  static constexpr int OFF_Q = 0;                                  // [2]
  static constexpr int OFF_KV = 2 * Q_BYTES;                       // [KV_SLOTS]
  static constexpr int OFF_STAT = OFF_KV + KV_SLOTS * KV_BYTES;    // l[2][128], m[2][128] fp32
  static constexpr int OFF_BAR = OFF_STAT + 2 * 2 * FA_BM * 4;
  static constexpr int TOTAL = OFF_BAR + 512;                      // the dynamic smem base is 1024-byte aligned (checked)
};

struct FaParams {
  long long* trace;  // debug (A3_FA_TRACE_PTR): per-CTA clock64 stamps of the roles' hand-offs, null in production
  __nv_bfloat16* o;
  long long ldo;
  float* lse;
  int n_items, BH, n_qt;  // work items (heaviest first), batch x heads, 128-query tiles per (batch, head)
  int B, H, Tq, Tk;
  float scale;
  int causal;
};

// One unit of work: up to two 128-query tiles of one (batch, head).  Items are numbered heaviest first
// (causal: the last query tiles see the most keys); persistent CTA c takes items c, c + grid, ... so every
// CTA receives a decreasing sequence of costs (LPT).  With an odd number of query tiles the unpaired one is
// tile 0 -- the cheapest (one key tile when causal) -- because a lone tile cannot overlap its softmax with
// its own MMAs.
struct FaItem {
  int b, h, q0_0, q0_1, nt_0, nt_1;
};

template <int HD>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  using S = FaSmem<HD>;
  constexpr int NCH = HD / 64;  // 64-column (128-byte) chunks per head row
  constexpr int NS = S::KV_SLOTS;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) return;  // SW128 tiles need a 1024-byte aligned base; the host probes this once per
                                       // device before the first launch and returns A3_ERR_UNSUPPORTED (fa_smem_probe)
  float* s_l = reinterpret_cast<float*>(smem + S::OFF_STAT);  // [2 stages][128]
  float* s_m = s_l + 2 * FA_BM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;             // [2]  Q producer -> MMA_s: Q tile of stage s landed
  uint64_t* q_free = bars + 2;         // [2]  MMA_s (commit) -> Q producer: the item's last S_s retired
  uint64_t* kv_full = bars + 28;       // [NS] K/V producer -> both MMA warps
  uint64_t* kv_empty = bars + 28 + NS; // [NS] both MMA warps (commit, count 2) -> K/V producer
  uint64_t* s_full = bars + 12;        // [2]  MMA_s (commit) -> softmax group s
  uint64_t* p_half = bars + 14;        // [2]  softmax group s (4 warps) -> MMA_s: keys [0,64) of P_s(j) are in TMEM
  uint64_t* p_full = bars + 16;        // [2]  softmax group s (4 warps) -> MMA_s: all of P_s(j)
  uint64_t* pv_done = bars + 18;       // [2]  MMA_s (commit) -> softmax group s (rescale) / epilogue: PV_s(j) retired
  uint64_t* fin = bars + 20;           // [2][2] softmax group s (128 threads) -> epilogue: l, m of the item final; by item
                                       //      parity, because a softmax group may finish the NEXT item (one key
                                       //      tile: its only S needs no O) before the epilogue has seen this one
  uint64_t* o_free = bars + 24;        // [2]  epilogue (128 threads) -> MMA_s / softmax s: O_s, l, m of the item have been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off = p.Tk - p.Tq;  // right-aligned causal offset
  // trace layout (debug): [cta][role 0..5 = softmax0, softmax1, mma0, mma1, K/V producer, epilogue][128 stamps]
  long long* trace = p.trace ? p.trace + static_cast<long long>(blockIdx.x) * 768 : nullptr;
  int tr_n = 0;
  auto TR = [&](int role) {
    if (trace && tr_n < 128 && (role < 2 || lane == 0)) trace[role * 128 + tr_n] = clock64();
    if (trace) ++tr_n;
  };
  auto tiles_of = [&](int q0) {
    const int kv_end = p.causal ? min(p.Tk, q0 + FA_BM + off) : p.Tk;
    return (q0 >= 0 && q0 < p.Tq && kv_end > 0) ? (kv_end + FA_BN - 1) / FA_BN : 0;
  };
  auto get_item = [&](int it) {
    FaItem w;
    const int rank = it / p.BH, bh = it - rank * p.BH;
    w.b = bh / p.H;
    w.h = bh - w.b * p.H;
    const int t0 = p.n_qt - 2 * (rank + 1);  // first query tile of the pair; -1: the lone tile 0
    w.q0_0 = (t0 < 0 ? 0 : t0) * FA_BM;
    w.q0_1 = t0 < 0 ? -1 : w.q0_0 + FA_BM;
    w.nt_0 = tiles_of(w.q0_0);
    w.nt_1 = tiles_of(w.q0_1);
    return w;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_half[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&fin[2 * i], FA_BM);      // every softmax thread releases its own l, m
      mbar_init(&fin[2 * i + 1], FA_BM);
      mbar_init(&o_free[i], FA_BM);       // every epilogue thread, after it has read l, m and its O row
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    fence_mbar_init();
  }
  if (warp == 15) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t C_S = 0, C_O = 256;

  if (warp < 8) {
    // =========================================================== softmax groups
    setmaxnreg_inc<FA_REGS_SOFTMAX>();
    const int s = warp >> 2;
    const int t = threadIdx.x & 127;
    const float c2 = p.scale * 1.4426950408889634f;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_S = t_lane + C_S + s * FA_BN, t_O = t_lane + C_O + s * HD;
    uint32_t c_tile = 0, c_item = 0;  // cumulative tiles / items of this stage (barrier phases)
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const FaItem w = get_item(it);
      const int n = s ? w.nt_1 : w.nt_0;
      if (n == 0) continue;
      const int row = (s ? w.q0_1 : w.q0_0) + t;
      const int limit = p.causal ? min(p.Tk - 1, row + off) : p.Tk - 1;  // last visible key of this row
      float m_use = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n; ++j, ++c_tile) {
        mbar_wait(&s_full[s], c_tile & 1);
        tc_fence_after();
        if (t == 0) TR(s);  // S ready
        uint32_t r[FA_BN];
        const int k0 = j * FA_BN;
        auto load_scores = [&]() {
          tmem_ld_32x32_at<0>(t_S, r);
          tmem_ld_32x32_at<32>(t_S + 32, r);
          tmem_ld_32x32_at<64>(t_S + 64, r);
          tmem_ld_32x32_at<96>(t_S + 96, r);
          tmem_ld_wait();
          if (k0 + FA_BN - 1 > limit) {
#pragma unroll
            for (int i = 0; i < FA_BN; ++i)
              if (k0 + i > limit) r[i] = 0xff800000u;  // -inf
          }
        };
        load_scores();
        float mxv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mxv[i] = fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 8]));
#pragma unroll
        for (int i = 16; i < FA_BN; i += 16) {  // eight independent 3-input max chains
#pragma unroll
          for (int e = 0; e < 8; ++e)
            mxv[e] = fmaxf(mxv[e], fmaxf(__uint_as_float(r[i + e]), __uint_as_float(r[i + 8 + e])));
        }
        const float mx = fmaxf(fmaxf(fmaxf(mxv[0], mxv[1]), fmaxf(mxv[2], mxv[3])),
                               fmaxf(fmaxf(mxv[4], mxv[5]), fmaxf(mxv[6], mxv[7])));
        float resc = 1.f;
        if (j == 0) {
          m_use = mx;
        } else {
          // lazy rescale: keep normalising by the stale maximum unless the true one grew by more than 2^8
          // (exact: l and O carry the same stale maximum).  Rare after the first tiles, so the softmax
          // thread rescales its own O row; the scores are re-read from TMEM afterwards so that the rare
          // path does not add its registers to the 128 live scores.
          const bool grow = (mx - m_use) * c2 > FA_RESCALE_LOG2;
          if (__any_sync(0xffffffffu, grow)) {  // warp-uniform: tcgen05.ld / st are warp collectives
            if (grow) {
              resc = ex2_approx((m_use - mx) * c2);
              m_use = mx;
            }
            mbar_wait(&pv_done[s], (c_tile - 1) & 1);  // PV_s(j - 1) has retired: O_s is stable
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(t_O + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * resc);
              tmem_st_32x32(t_O + c * 32, o);
            }
            tmem_st_wait();
            load_scores();
          }
        }
        if (t == 0) TR(s);  // max known
        const float mc = (m_use == -INFINITY ? 0.f : m_use) * c2;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int c = 0; c < FA_BN / 32; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float x0, x1, x2, x3;  // packed fp32 FMAs: two scores per issue slot
            ffma2(x0, x1, __uint_as_float(r[c * 32 + 2 * i]), __uint_as_float(r[c * 32 + 2 * i + 1]), c2, -mc);
            ffma2(x2, x3, __uint_as_float(r[c * 32 + 2 * i + 2]), __uint_as_float(r[c * 32 + 2 * i + 3]), c2, -mc);
            // (a cubic on the FMA pipe for a fraction of these was measured SLOWER -- 97 -> 104 / 109 us at 25 / 50 %:
            // the softmax warps are issue- and latency-bound next to MUFU, not MUFU-bound alone)
            x0 = ex2_approx(x0);
            x1 = ex2_approx(x1);
            x2 = ex2_approx(x2);
            x3 = ex2_approx(x3);
            pk[i] = pack_bf16x2(x0, x1);
            pk[i + 1] = pack_bf16x2(x2, x3);
            fadd2(sum0, sum1, sum0, sum1, x0, x1);
            fadd2(sum2, sum3, sum2, sum3, x2, x3);
          }
          tmem_st_32x16(t_S + c * 16, pk);
          if (c == 1) {  // keys [0, 64) of P are stored: the first four k-steps of P V may start
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[s]);
          }
        }
        l_run = l_run * resc + ((sum0 + sum1) + (sum2 + sum3));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (t == 0) TR(s);  // P published
      }
      // the epilogue must have consumed the previous item's l, m (a one-tile item can be finished before that)
      if (c_item > 0) mbar_wait(&o_free[s], (c_item - 1) & 1);
      s_l[s * FA_BM + t] = l_run;
      s_m[s * FA_BM + t] = m_use;
      mbar_arrive(&fin[2 * s + (c_item & 1)]);
      ++c_item;
    }
  } else if (warp < 12) {
    // =========================================================== epilogue group
    // O_s / l -> bf16 -> global (each thread its own row), lse; hands O_s back to MMA_s for the next item
    setmaxnreg_dec<FA_REGS_EPILOGUE>();
    const int t = threadIdx.x & 127;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t c_pv0 = 0, c_pv1 = 0, c_item0 = 0, c_item1 = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const FaItem w = get_item(it);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int n = s ? w.nt_1 : w.nt_0;
        const int q0 = s ? w.q0_1 : w.q0_0;
        if (q0 < 0 || q0 >= p.Tq) continue;
        uint32_t& c_pv = s ? c_pv1 : c_pv0;
        uint32_t& c_item = s ? c_item1 : c_item0;
        float l = 0.f, m = -INFINITY;
        if (n > 0) {
          c_pv += n;
          mbar_wait(&fin[2 * s + (c_item & 1)], (c_item >> 1) & 1);
          mbar_wait(&pv_done[s], (c_pv - 1) & 1);  // the item's last PV_s
          tc_fence_after();
          if (t == 0) TR(5);  // epilogue of (item, s) starts
          l = s_l[s * FA_BM + t];
          m = s_m[s * FA_BM + t];
          ++c_item;
        }
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
        const int row = q0 + t;
        const bool row_ok = row < p.Tq;
        __nv_bfloat16* orow = p.o + (static_cast<long long>(w.b) * p.Tq + row) * p.ldo + w.h * HD;
        const uint32_t t_O = t_lane + C_O + s * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t o[32];
          if (n > 0) {
            tmem_ld_32x32(t_O + c * 32, o);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
          }
          if (c == HD / 32 - 1 && n > 0) {  // O_s is in registers: the next item's first PV_s may overwrite it
            tc_fence_before();
            mbar_arrive(&o_free[s]);
          }
          if (row_ok) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 q;
              q.x = pack_bf16x2(__uint_as_float(o[u * 8]) * inv_l, __uint_as_float(o[u * 8 + 1]) * inv_l);
              q.y = pack_bf16x2(__uint_as_float(o[u * 8 + 2]) * inv_l, __uint_as_float(o[u * 8 + 3]) * inv_l);
              q.z = pack_bf16x2(__uint_as_float(o[u * 8 + 4]) * inv_l, __uint_as_float(o[u * 8 + 5]) * inv_l);
              q.w = pack_bf16x2(__uint_as_float(o[u * 8 + 6]) * inv_l, __uint_as_float(o[u * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) = q;
            }
          }
        }
        if (row_ok && p.lse)
          p.lse[(static_cast<long long>(w.b) * p.H + w.h) * p.Tq + row] = l > 0.f ? m * p.scale + logf(l) : -INFINITY;
      }
    }
  } else {
    setmaxnreg_dec<FA_REGS_CONTROL>();
    if (warp == 12 || warp == 14) {
      // =========================================================== MMA issuer of stage s (warp 12: 0, warp 14: 1)
      // Each query tile has its own issuer, so the two stages are independent instruction streams that only
      // share the K/V ring: a stage that finishes an item (the causal pair is n and n + 1 key tiles long) goes
      // straight on to the next one.  The whole warp runs the (warp-uniform) control flow and the waits; one
      // elected lane issues the tcgen05 instructions.  Operand addresses are made provably uniform (broadcast
      // shuffles) and the descriptors are bases + immediates, so an MMA costs a couple of instructions to issue
      // -- with a lane-0-only branch the compiler wraps every UTCHMMA in an ELECT / R2UR waterfall loop, which
      // at 64 cycles per MMA made the issuing thread, not the tensor pipe, the limiter.
      const int s = (warp - 12) >> 1;
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sQ = __shfl_sync(0xffffffffu, smem_u32(smem + S::OFF_Q + s * S::Q_BYTES), 0);
      const uint32_t sKV = __shfl_sync(0xffffffffu, smem_u32(smem + S::OFF_KV), 0);
      constexpr uint32_t idesc_s = make_idesc_bf16(FA_BM, FA_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, HD, 0, 1);
      // descriptor = (constant high word, low word = start address >> 4 | LBO field): one add per k-step
      constexpr uint64_t hi_k = (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      constexpr uint32_t lbo_k = (16u >> 4) << 16, lbo_v = ((FA_BN * 128u) >> 4) << 16;
      auto desc = [&](uint32_t addr, uint32_t lbo) { return hi_k | lbo | ((addr >> 4) & 0x3fffu); };
      const uint32_t d_S = tm + C_S + s * FA_BN, d_O = tm + C_O + s * HD;
      const uint64_t qd = desc(sQ, lbo_k);
      // K/V ring cursor (K_j, V_j alternate; walked strictly in order, continues across items)
      uint32_t slot = 0, phase = 0;
      uint32_t c_q = 0, c_tile = 0, c_item = 0;
      auto kv_wait = [&]() {
        mbar_wait(&kv_full[slot], phase);
        tc_fence_after();
      };
      auto kv_release = [&]() {  // arrives once this thread's MMAs so far have retired (at once if none); advances
        if (leader) umma_commit(&kv_empty[slot]);
        __syncwarp();
        if (++slot == NS) {
          slot = 0;
          phase ^= 1;
        }
      };
      auto issue_s = [&]() {  // S_s = Q_s K^T, K in the current slot
        const uint64_t kd = desc(sKV + slot * S::KV_BYTES, lbo_k);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t o16 = ((kk >> 2) * (FA_BM * 128) + (kk & 3) * 32) >> 4;
            umma_bf16(d_S, qd + o16, kd + o16, idesc_s, kk > 0 ? 1u : 0u);
          }
        }
        __syncwarp();  // reconverge: the other lanes must not run ahead into the next wait loop
      };
      auto issue_pv = [&](bool acc, int half) {  // O_s (+)= P_s V (current slot), keys [64 half, 64 half + 64)
        const uint64_t vd = desc(sKV + slot * S::KV_BYTES, lbo_v);
        if (leader) {
#pragma unroll
          for (int k4 = 0; k4 < FA_BN / 32; ++k4) {
            const int kk = half * (FA_BN / 32) + k4;
            umma_bf16_ts(d_O, d_S + kk * 8, vd + ((kk * 16 * 128) >> 4), idesc_o, (acc || kk > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (leader) umma_commit(bar);
        __syncwarp();
      };
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const FaItem w = get_item(it);
        const int n = s ? w.nt_1 : w.nt_0, nmax = max(w.nt_0, w.nt_1);
        if (nmax == 0) continue;
        // ring order of this item: K_0, V_0, K_1, V_1, ... (walked by both issuers, used or not)
        TR(2 + s);  // item start
        kv_wait();
        TR(2 + s);  // K_0 landed
        if (n > 0) {
          mbar_wait(&q_full[s], c_q & 1);
          ++c_q;
          tc_fence_after();
          TR(2 + s);  // Q landed
          issue_s();
          commit(&s_full[s]);
          if (n == 1) commit(&q_free[s]);  // the item's only S_s
        }
        kv_release();
        for (int j = 0; j < nmax; ++j) {
          kv_wait();  // V_j landed long ago: take this wait before the ones that matter
          if (j < n) {
            if (j == 0) {
              TR(2 + s);  // V_0 landed
              mbar_wait(&o_free[s], (c_item & 1) ^ 1);  // the previous item's O_s has been read out
              ++c_item;
              TR(2 + s);  // O free
            }
            mbar_wait(&p_half[s], c_tile & 1);
            tc_fence_after();
            TR(2 + s);  // first half of P seen
            issue_pv(j > 0, 0);
            mbar_wait(&p_full[s], c_tile & 1);
            tc_fence_after();
            issue_pv(j > 0, 1);
            ++c_tile;
            commit(&pv_done[s]);
          }
          kv_release();
          if (j + 1 < nmax) {
            kv_wait();
            if (j + 1 < n) {
              TR(2 + s);  // PV issued, K landed
              issue_s();
              commit(&s_full[s]);
              if (j + 2 == n) commit(&q_free[s]);  // the item's last S_s
              TR(2 + s);  // S issued
            }
            kv_release();
          }
        }
      }
    } else if (warp == 13) {
      // =========================================================== K/V producer (warp-uniform, elected lane issues)
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t bfull = __shfl_sync(0xffffffffu, smem_u32(kv_full), 0);
      uint32_t slot = 0, phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const FaItem w = get_item(it);
        const int nmax = max(w.nt_0, w.nt_1);
        for (int i = 0; i < 2 * nmax; ++i) {
          const int j = i >> 1;
          mbar_wait(&kv_empty[slot], phase ^ 1);
          TR(4);  // slot free -> load issued
          const CUtensorMap* tm = (i & 1) ? &tmV : &tmK;
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[slot], S::KV_BYTES);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              tma_load_3d_s(sb + S::OFF_KV + slot * S::KV_BYTES + c * (FA_BN * 128), tm, bfull + slot * 8,
                            w.h * HD + c * 64, j * FA_BN, w.b);
          }
          __syncwarp();
          if (++slot == NS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 15) {
      // =========================================================== Q producer (warp-uniform, elected lane issues)
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t bfull = __shfl_sync(0xffffffffu, smem_u32(q_full), 0);
      uint32_t c_q0 = 0, c_q1 = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const FaItem w = get_item(it);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if ((s ? w.nt_1 : w.nt_0) == 0) continue;
          uint32_t& c_q = s ? c_q1 : c_q0;
          mbar_wait(&q_free[s], (c_q & 1) ^ 1);  // the previous item's S_s MMAs have retired
          ++c_q;
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[s], S::Q_BYTES);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              tma_load_3d_s(sb + S::OFF_Q + s * S::Q_BYTES + c * (FA_BM * 128), &tmQ, bfull + s * 8, w.h * HD + c * 64,
                            s ? w.q0_1 : w.q0_0, w.b);
          }
          __syncwarp();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 15) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
static int launch_attn_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                           const FaParams& p, cudaStream_t stream) {
  using S = FaSmem<HD>;
  if (int rc = check_dyn_smem_alignment("a3_attn_fwd")) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&attn_fwd_kernel<HD>), S::TOTAL)) return rc;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  attn_fwd_kernel<HD><<<grid, FA_THREADS, S::TOTAL, stream>>>(tmQ, tmK, tmV, p);
  return check_launch("a3_attn_fwd");
}

}  // namespace a3

extern "C" int a3_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                           int H, int Tq, int Tk, int head_dim, long long ldq, long long ldk,
                           long long ldv, long long ldo, long long q_bstride, long long kv_bstride,
                           float scale, int causal, void* stream) {
  using namespace a3;
  if (B <= 0 || H <= 0 || Tq <= 0) return A3_OK;
  if (Tk <= 0) return set_error(A3_ERR_ARG, "a3_attn_fwd: Tk must be positive");
  if (head_dim != 64 && head_dim != 128)
    return set_error(A3_ERR_UNSUPPORTED, "a3_attn_fwd: head_dim must be 64 or 128");
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7) || (q_bstride & 7) || (kv_bstride & 7) ||
      (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
    return set_error(A3_ERR_ARG, "a3_attn_fwd: strides % 8 == 0 and 16-byte aligned pointers required");
  static const int legacy = getenv("A3_ATTN_FWD_LEGACY") ? atoi(getenv("A3_ATTN_FWD_LEGACY")) : 0;  // A/B testing only
  if (legacy)
    return attn_fwd_legacy(q, k, v, o, lse, B, H, Tq, Tk, head_dim, ldq, ldk, ldv, ldo, q_bstride, kv_bstride, scale,
                           causal, stream);
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t width = (uint64_t)H * head_dim;
  int rc = make_tmap_3d_bf16(&tmQ, q, width, (uint64_t)Tq, (uint64_t)B, (uint64_t)ldq * 2,
                             (uint64_t)q_bstride * 2, 64, FA_BM, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmK, k, width, (uint64_t)Tk, (uint64_t)B, (uint64_t)ldk * 2,
                         (uint64_t)kv_bstride * 2, 64, FA_BN, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmV, v, width, (uint64_t)Tk, (uint64_t)B, (uint64_t)ldv * 2,
                         (uint64_t)kv_bstride * 2, 64, FA_BN, 1);
  if (rc) return rc;
  FaParams p;
  static long long* const trace_ptr =
      getenv("A3_FA_TRACE_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("A3_FA_TRACE_PTR"), nullptr, 0)) : nullptr;
  p.trace = trace_ptr;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  p.n_qt = (Tq + FA_BM - 1) / FA_BM;
  p.BH = B * H;
  p.n_items = ((p.n_qt + 1) / 2) * p.BH;  // pairs of query tiles, plus the lone tile 0 when their number is odd
  p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk;
  p.scale = scale;
  p.causal = causal ? 1 : 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return head_dim == 64 ? launch_attn_fwd<64>(tmQ, tmK, tmV, p, s) : launch_attn_fwd<128>(tmQ, tmK, tmV, p, s);
}

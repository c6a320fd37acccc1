// Decode-time projection on tcgen05 with swapped operands (sm_100a): for a batch of M <= 64 single-token
// rows the weight matrix is the 128-row MMA operand and the batch is the narrow N side,
//
//     Y^T[n][b] = sum_k W[n][k] * X[b][k]        D (TMEM) = A (W tile, 128 x 64) * B (X tile, NB x 64)
//
// so every byte TMA brings into shared memory is a weight byte the MMA uses (the plain GEMM would stage a
// 128-row activation tile of which M rows are live).  The kernel is HBM-bound by construction: what
// matters is bytes in flight: two CTAs per SM with 4-6 ring stages of 18-24 KB each, the grid is
// (N / 128) tiles x split-K sized to one resident wave, and the K splits of one tile form a thread-block cluster whose partial accumulators are
// reduced through distributed shared memory by the leader CTA.  Launched with programmatic dependent
// launch: the weight tiles of the first ring pass are requested BEFORE griddepcontrol.wait, the
// activations (which the previous kernel produced) after it.
//
// Replaces the decode-step (seqlen 1) projections of reference model/accessory/model/LLM/llama_ens5.py:112,
// 169,216-217,530 for batches above the CUDA-core weight-streaming kernel's range.
#include "common.cuh"
#include "ptx.cuh"

namespace a3 {

constexpr int DG_BM = 128;  // weight rows (output features) per tile
constexpr int DG_BK = 64;
constexpr int DG_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

template <int NB>
struct DgSmem {
  static constexpr int W_BYTES = DG_BM * DG_BK * 2;  // 16 KB
  static constexpr int X_BYTES = NB * DG_BK * 2;
  static constexpr int X_SLOT = X_BYTES < 1024 ? 1024 : X_BYTES;  // SW128 tiles start on 1024 B
  static constexpr int STAGE_BYTES = W_BYTES + X_SLOT;
  static constexpr int STAGES = (108 * 1024) / STAGE_BYTES;  // two CTAs per SM: 2 x 4-6 stages in flight
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = NB < 32 ? 32 : NB;
};

struct DgParams {
  void* Y;
  const __nv_bfloat16* residual;
  long long ldy, ldr;
  int M, N, K, out_fp32, splits;
};

__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

template <int NB>
__global__ void __launch_bounds__(DG_THREADS, 2)
decode_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const DgParams p) {
  using S = DgSmem<NB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + S::STAGES;
  uint64_t* acc_full = bars + 2 * S::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 1);
  float* partial = reinterpret_cast<float*>(smem);  // [NB][128] fp32, reuses the drained ring

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = p.splits > 1 ? (int)cluster_ctarank() : 0;
  const int nb = blockIdx.x / p.splits;
  const int n0 = nb * DG_BM;
  const int kb_total = (p.K + DG_BK - 1) / DG_BK;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int kb_begin = ks * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
  const int nkb = max(kb_end - kb_begin, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // producer and issuer warps run warp-uniform with one elected lane issuing, so UTMALDG / UTCHMMA take their
  // operands from uniform registers (no ELECT / R2UR waterfall loop per instruction; the MMAs here are 8-32 cycles)
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    // first ring pass: weights now, activations after the previous kernel has finished
    const int first = min(nkb, S::STAGES);
    if (leader) {
      for (int i = 0; i < first; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], S::W_BYTES + S::X_BYTES);
        tma_load_2d_s(smem0 + i * S::STAGE_BYTES, &tmW, &full_bar[i], (kb_begin + i) * DG_BK, n0);
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (leader) {
      for (int i = 0; i < first; ++i)
        tma_load_2d_s(smem0 + i * S::STAGE_BYTES + S::W_BYTES, &tmX, &full_bar[i], (kb_begin + i) * DG_BK, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 1;  // the first pass filled every slot once
    for (int i = first; i < nkb; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t sw = smem0 + stage * S::STAGE_BYTES;
      if (leader) {
        mbar_arrive_expect_tx(&full_bar[stage], S::W_BYTES + S::X_BYTES);
        tma_load_2d_s(sw, &tmW, &full_bar[stage], (kb_begin + i) * DG_BK, n0);
        tma_load_2d_s(sw + S::W_BYTES, &tmX, &full_bar[stage], (kb_begin + i) * DG_BK, 0);
      }
      __syncwarp();
      if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc = make_idesc_bf16(DG_BM, NB < 16 ? 16 : NB, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sw = smem0 + stage * S::STAGE_BYTES;
      const uint64_t adesc = make_smem_desc_sw128(sw, 16, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(sw + S::W_BYTES, 16, 1024);
      if (leader) {
#pragma unroll
        for (int k = 0; k < DG_BK / 16; ++k)
          umma_bf16(tm, adesc + k * 2, bdesc + k * 2, idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (i == nkb - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
    }
  }

  // ---- accumulator -> registers (epilogue warps: TMEM lane quarter = warp % 4)
  float acc[NB];
  const int q = warp & 3;
  const int feat = q * 32 + lane;  // feature row of this thread inside the tile
  if (warp >= 2) {
    if (nkb > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < (NB + 31) / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < NB) acc[c * 32 + j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] = 0.f;
    }
  }
  if (p.splits > 1) {
    // split-K: every CTA parks its partial tile in its own shared memory, the leader sums them
    if (warp >= 2) {
      // the ring is drained: all TMA writes were observed through full_bar by the MMA warp and the
      // MMAs reading them have retired (acc_full)
#pragma unroll
      for (int j = 0; j < NB; ++j) partial[j * DG_BM + feat] = acc[j];
    }
    cluster_sync_all();
    if (ks == 0 && warp >= 2) {
      for (int r = 1; r < p.splits; ++r) {
        const uint32_t remote = mapa_u32(partial, (uint32_t)r);
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] += ld_cluster_f32(remote + (uint32_t)(j * DG_BM + feat) * 4u);
      }
    }
  }
  if (ks == 0 && warp >= 2) {
    const int n = n0 + feat;
    if (n < p.N) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < p.M) {
          float v = acc[j];
          if (p.out_fp32) {
            v = bf16_round(v);  // logits: .float() of a bf16 GEMM result (llama_ens5.py:530-531)
            if (p.residual) v += __bfloat162float(p.residual[j * p.ldr + n]);
            reinterpret_cast<float*>(p.Y)[j * p.ldy + n] = v;
          } else {
            if (p.residual) v = bf16_round(v) + __bfloat162float(p.residual[j * p.ldr + n]);
            reinterpret_cast<__nv_bfloat16*>(p.Y)[j * p.ldy + n] = __float2bfloat16_rn(v);
          }
        }
      }
    }
  }
  if (p.splits > 1) cluster_sync_all();  // keep the partials alive until the leader has read them
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int NB>
static int launch_decode_gemm(const CUtensorMap& tmW, const CUtensorMap& tmX, const DgParams& p,
                              cudaStream_t stream) {
  using S = DgSmem<NB>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&decode_gemm_kernel<NB>), S::TOTAL)) return rc;
  const int n_blocks = (p.N + DG_BM - 1) / DG_BM;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_blocks * p.splits);
  cfg.blockDim = dim3(DG_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (p.splits > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.splits;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_gemm_kernel<NB>, tmW, tmX, p);
  if (e != cudaSuccess) return set_error(A3_ERR_CUDA, cudaGetErrorString(e));
  return check_launch("a3_decode_gemm");
}

}  // namespace a3

extern "C" int a3_decode_gemm(const void* X, const void* W, void* Y, int M, int N, int K, long long ldx,
                              long long ldw, long long ldy, const void* residual, long long ldr,
                              int out_fp32, void* stream) {
  using namespace a3;
  if (M <= 0 || N <= 0 || K <= 0) return A3_OK;
  if (M > 64) return set_error(A3_ERR_ARG, "a3_decode_gemm: M <= 64 (use a3_gemm_bf16)");
  if ((K & 7) || (ldx & 7) || (ldw & 7) ||
      ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15))
    return set_error(A3_ERR_ARG, "a3_decode_gemm: K, ldx, ldw % 8 == 0 and 16-byte aligned pointers");
  const int nbatch = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  CUtensorMap tmW, tmX;
  int rc = make_tmap_2d_bf16(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, 64, DG_BM);
  if (rc != A3_OK) return rc;
  rc = make_tmap_2d_bf16(&tmX, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, 64, (uint32_t)nbatch);
  if (rc != A3_OK) return rc;
  DgParams p;
  p.Y = Y;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldy = ldy; p.ldr = ldr; p.M = M; p.N = N; p.K = K; p.out_fp32 = out_fp32 ? 1 : 0;
  // one resident wave: two CTAs fit per SM, so split K while the grid stays within 2 x SMs slots
  // (each split keeps >= 8 k-blocks); more waves only add per-CTA prologue / epilogue time
  const int n_blocks = (N + DG_BM - 1) / DG_BM, kb_total = (K + DG_BK - 1) / DG_BK;
  int splits = 1;
  while (splits < 8 && n_blocks * splits * 2 <= 2 * num_sms() && kb_total / (splits * 2) >= 8) splits *= 2;
  p.splits = splits;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (nbatch == 16) return launch_decode_gemm<16>(tmW, tmX, p, s);
  if (nbatch == 32) return launch_decode_gemm<32>(tmW, tmX, p, s);
  return launch_decode_gemm<64>(tmW, tmX, p, s);
}

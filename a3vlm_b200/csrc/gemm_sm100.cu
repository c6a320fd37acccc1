// Persistent, warp-specialised bf16 GEMM for sm_100a:  TMA -> smem ring -> tcgen05.mma -> TMEM ->
// fused epilogue.  One kernel serves every GEMM-shaped op on the A3VLM hot path (ViT / projector /
// Llama projections and MLPs, LM head, and their dgrad / wgrad forms) through operand "majorness":
//
//   C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
//     A K-major : stored [M][K] (k contiguous, row stride lda)      -- activations in forward / dgrad
//     A MN-major: stored [K][M] (m contiguous, row stride lda)      -- dy^T in wgrad
//     B K-major : stored [N][K] (k contiguous, row stride ldb)      -- nn.Linear weight (out,in) in forward
//     B MN-major: stored [K][N] (n contiguous, row stride ldb)      -- weight in dgrad, x^T in wgrad
//
// Replaces the cuBLAS calls behind fairscale Column/RowParallelLinear and nn.Linear in
// reference model/accessory/model/LLM/llama_ens5.py:63-90,202-210,267-269,330-333 and the
// open_clip ViT linears (llama_ens5.py:354-370).
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warp 2 TMEM
// allocator, warps 4..11 epilogue (TMEM lane quarter = warp_idx % 4, two warps per quarter split the
// columns).  CTA = 384 threads.  Tile 128 x BN x 64, 4-8 smem
// stages, two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "ptx.cuh"
#include "a3vlm_b200.h"
#include <atomic>
#include <cstring>
#include "common.cuh"
#include <cstdlib>

namespace a3 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // max: warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
// The epilogue runs on 4 warps (256-thread launch) or 8 warps (384 threads, two per TMEM lane quarter,
// each taking half of the tile's columns).  Eight warps halve heavy epilogues (GELU, short-K ViT tiles:
// c_fc 137 -> 80 us) but cost the long LLM projections ~3 % (measured), so the host picks per call.

struct GemmParams {
  void* C;
  const __nv_bfloat16* bias;      // [N] bf16 or null
  const __nv_bfloat16* residual;  // [M][ldr] bf16 or null
  const float2* rope;             // [max_pos][64] (cos, sin) or null
  long long ldc, ldr;
  int rope_cols, rope_period, rope_pos0;
  int M, N, K;
  int a_mn, b_mn;  // 1 = MN-major operand
  int act;         // 0 none, 1 gelu (erf), 2 quick-gelu
  int out_fp32;
  int accumulate;
  float alpha;
  int group_m;  // M-blocks per rasterisation group
  unsigned long long hint_a, hint_b;  // L2 cache policies of the operand loads (CTA-pair kernel)
  unsigned int* tile_ctr;             // dynamic tile claim (CTA-pair kernel, A3_GEMM_DYNAMIC=1): zero at launch, self-resetting
  // fused gated-SiLU modes (llama_ens5.py:213-217):
  //   swiglu_F > 0 (CTA-pair kernel only): B = [w1; w3] ([2F][K]); CTA 0 of the pair loads 128 rows of
  //     w1, CTA 1 the same 128 rows of w3, so one accumulator tile holds g | u of 128 features; the
  //     epilogue writes gu (C, [M][2F]) and y = bf16(bf16(silu(g)) * u) (C2, [M][F]).
  //   dswiglu_gu != null: the GEMM computes dy = d W2 (N = F); the epilogue turns it into
  //     dgu = (dg | du) ([M][2F], C) from the saved gu -- dy never touches HBM.
  int swiglu_F;
  __nv_bfloat16* C2;
  long long ldc2;
  const __nv_bfloat16* dswiglu_gu;
  long long ldgu;
  // small-M (decode-time) mode of the single-CTA kernel: only a_rows (< 128) rows of A are staged per
  // k-block; the remaining rows of the 128-row MMA tile read stale shared memory and their
  // accumulator rows are never stored.
  int a_rows;
  // im2col-free patch embedding (single-CTA kernel, a3_vit_patch_embed): A is never materialised.  The "A" tensor
  // map is a 4-D view of the bf16 image whose rows were cast with every patch column on a 16-element pitch
  // (a3_vit_pad_cast), {x: 16 g, ky: patch, py: grid, channel-of-image: 3B}, 32-byte swizzle.  One TMA box
  // {16 x, 1 ky, 8 py, 1} lands as one 8-row x 32-byte atom of a K-major SW32 operand: row = patch (px fixed, 8 py),
  // 16 k = the kx of one kernel row (kx >= patch is zero padding in both operands; ky >= patch is out of bounds of
  // the map: zero fill).  (Measured: with a swizzle mode TMA puts every inner box row on its own swizzle-span line,
  // so a {16, 4, 8} box does NOT give 128-byte K rows; the 32-byte mode with one kernel row per box does.)  A stage
  // holds four such [128 rows][32 B] sub-tiles, one per k-step of the MMA, whose A descriptor is SW32 while B (the
  // weights) stays SW128.  A 128-row tile = 16 x 4 boxes per k-block;
  // k-block kb covers channel kb / 4, kernel rows 4 (kb % 4) .. +3.  Rows are ordered (image, px, py-group, py) and
  // the epilogue writes accumulator row r to output row image * g^2 + py * g + px.
  int patch_g;      // grid side g (0: off)
  int patch_p;      // patch size (<= 16)
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int group_m,
                                            int& mb, int& nb) {
  // Grouped rasterisation: walk all N-blocks for a band of `group_m` M-blocks before moving on,
  // so one wave of CTAs shares a small set of A rows and B rows in L2.
  const int per_group = group_m * n_blocks;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int rows = min(group_m, m_blocks - first_m);
  const int r = tile - g * per_group;
  mb = first_m + r % rows;
  nb = r / rows;
}

}  // namespace a3
// L2 policies of the operand loads.  The raster walks all N-blocks for a band of group_m M-blocks, so the band's A
// panel is re-read by every wave of the band while a B panel is read once per band: A3_GEMM_HINT=al_bf marks A
// evict-last / B evict-first, af_bl the opposite, unset / 0 leaves both evict-normal (the measured default).
static void set_cache_hints(a3::GemmParams& p) {
  constexpr unsigned long long NORMAL = 0x1000000000000000ull, FIRST = 0x12F0000000000000ull, LAST = 0x14F0000000000000ull;
  static const char* env = getenv("A3_GEMM_HINT");
  p.hint_a = p.hint_b = NORMAL;
  if (!env) return;
  if (!strcmp(env, "al_bf")) { p.hint_a = LAST; p.hint_b = FIRST; }
  else if (!strcmp(env, "af_bl")) { p.hint_a = FIRST; p.hint_b = LAST; }
  else if (!strcmp(env, "al_bn")) { p.hint_a = LAST; }
  else if (!strcmp(env, "an_bf")) { p.hint_b = FIRST; }
}
namespace a3 {
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (act == 2) return v / (1.0f + __expf(-1.702f * v));
  return v;
}


// Epilogue of one accumulator tile for the 32 rows a warp owns: TMEM -> registers -> fused
// alpha / bias / activation / RoPE / residual / accumulate -> global (bf16 or fp32).
__device__ __forceinline__ void load8f(const __nv_bfloat16* ptr, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(ptr);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = unpack_bf16x2(w[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ void store8f(__nv_bfloat16* ptr, const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(ptr) = o;
}

// Backward of the gated SiLU fused behind the w2 dgrad: acc = dy tile; dg = bf16(dy*u) * silu'(g),
// du = dy * bf16(silu(g))  (same rounding points as the stand-alone a3_swiglu_bwd).
template <int BN>
__device__ __forceinline__ void epilogue_dswiglu_rows(const GemmParams& p, uint32_t t_row, int row, int n0,
                                                      int c0, int c1) {
  const bool row_ok = row < p.M;
  const int F = p.N;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c * 32, r);
    tmem_ld_wait();
    const int col0 = n0 + c * 32;
    if (row_ok && col0 < F) {
      const __nv_bfloat16* gp = p.dswiglu_gu + static_cast<long long>(row) * p.ldgu + col0;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (col0 + j < F) {
          float g[8], u[8], dg[8], du[8];
          load8f(gp + j, g);
          load8f(gp + F + j, u);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dy = bf16_round(__uint_as_float(r[j + e]) * p.alpha);
            const float sig = 1.f / (1.f + __expf(-g[e]));
            dg[e] = bf16_round(dy * u[e]) * sig * (1.f + g[e] * (1.f - sig));
            du[e] = dy * bf16_round(g[e] * sig);
          }
          store8f(op + j, dg);
          store8f(op + F + j, du);
        }
      }
    }
  }
}

// Forward gated SiLU fused behind the w1 | w3 projection (CTA-pair kernel, swiglu_F mode): columns
// [0,128) of the accumulator tile are g, [128,256) are u of features f0 .. f0+127.
__device__ __forceinline__ void epilogue_swiglu_rows(const GemmParams& p, uint32_t t_row, int row, int f0,
                                                     int c0, int c1) {
  const bool row_ok = row < p.M;
  const int F = p.swiglu_F;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    uint32_t rg[32], ru[32];
    tmem_ld_32x32(t_row + c * 32, rg);
    tmem_ld_32x32(t_row + 128 + c * 32, ru);
    tmem_ld_wait();
    if (row_ok) {
      __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + f0 + c * 32;
      __nv_bfloat16* yp = p.C2 + static_cast<long long>(row) * p.ldc2 + f0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float g[8], u[8], y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          g[e] = bf16_round(__uint_as_float(rg[j + e]));
          u[e] = bf16_round(__uint_as_float(ru[j + e]));
          y[e] = bf16_round(g[e] / (1.f + __expf(-g[e]))) * u[e];
        }
        store8f(gp + j, g);
        store8f(gp + F + j, u);
        store8f(yp + j, y);
      }
    }
  }
}

// Two warps share a TMEM lane quarter: `half` selects which half of the tile's 32-column chunks a warp
// handles (c0..c1), doubling the threads that hide the epilogue's TMEM / global latencies and ALU work.
template <int BN>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, uint32_t t_row, int row, int n0, int half,
                                              int nhalves) {
  constexpr int NCH = BN / 32;
  const int per = (NCH + nhalves - 1) / nhalves;
  const int c0 = half * per;
  const int c1 = min(NCH, c0 + per);
  if (p.dswiglu_gu) {
    epilogue_dswiglu_rows<BN>(p, t_row, row, n0, c0, c1);
    return;
  }
  const bool row_ok = row < p.M;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c * 32, r);
    tmem_ld_wait();
    const int col0 = n0 + c * 32;
    if (row_ok && col0 < p.N) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
      const int ncols = min(32, p.N - col0);  // N % 8 == 0 is enforced on the host
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncols) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = unpack_bf16x2(w[t]);
              v[j + 2 * t] += f.x;
              v[j + 2 * t + 1] += f.y;
            }
          }
        }
      }
      if (!p.out_fp32) {
        // bf16 path mirrors the reference's rounding points: Linear output rounds to bf16,
        // the activation rounds, the residual add rounds.
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(bf16_round(v[j]), p.act);
        }
        if (col0 < p.rope_cols) {
          // apply_rotary_emb on the bf16-rounded projection, fp32 math, one rounding at the
          // store.  Heads are 128 wide, so pair index = (col % 128) / 2.
          const int pos = p.rope_pos0 + row % p.rope_period;
          const float2* tp = p.rope + static_cast<long long>(pos) * 64 + ((col0 & 127) >> 1);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 cs = __ldg(tp + (j >> 1));
            const float a = bf16_round(v[j]), b = bf16_round(v[j + 1]);
            v[j] = a * cs.x - b * cs.y;
            v[j + 1] = a * cs.y + b * cs.x;
          }
        }
        if (p.residual) {
          const __nv_bfloat16* rp = p.residual + static_cast<long long>(row) * p.ldr + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_bf16x2(w[t]);
                v[j + 2 * t] = bf16_round(v[j + 2 * t]) + f.x;
                v[j + 2 * t + 1] = bf16_round(v[j + 2 * t + 1]) + f.y;
              }
            }
          }
        }
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) +
                            static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncols) {
            if (p.accumulate) {
              const uint4 q = *reinterpret_cast<const uint4*>(cp + j);
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_bf16x2(w[t]);
                v[j + 2 * t] += f.x;
                v[j + 2 * t + 1] += f.y;
              }
            }
            uint4 o;
            o.x = pack_bf16x2(v[j], v[j + 1]);
            o.y = pack_bf16x2(v[j + 2], v[j + 3]);
            o.z = pack_bf16x2(v[j + 4], v[j + 5]);
            o.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(cp + j) = o;
          }
        }
      } else {
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
        }
        if (p.residual) {
          const __nv_bfloat16* rp = p.residual + static_cast<long long>(row) * p.ldr + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) v[j] += __bfloat162float(rp[j]);
        }
        float* cp = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < ncols) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.accumulate) {
              const float4 q = *reinterpret_cast<const float4*>(cp + j);
              o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            *reinterpret_cast<float4*>(cp + j) = o;
          }
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + S::STAGES;         // [STAGES]
  uint64_t* tmem_full = bars + 2 * S::STAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * S::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // patch mode: accumulator rows run over (image, px, py-group, 8 py), i.e. M is padded to whole groups
  const int m_rows = p.patch_g ? (p.M / (p.patch_g * p.patch_g)) * p.patch_g * ((p.patch_g + 7) >> 3) * 8 : p.M;
  const int m_blocks = (m_rows + BM - 1) / BM;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], (blockDim.x >> 5) - 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb;
        tile_coords(tile, m_blocks, n_blocks, p.group_m, mb, nb);
        const int m0 = mb * BM, n0 = nb * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sa = smem0 + stage * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          const int k0 = kb * BK;
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], p.a_rows * (BK * 2) + S::B_BYTES);
            if (p.patch_g) {
              const int gpp = (p.patch_g + 7) >> 3;         // 8-row py groups per patch column
              const int gpi = p.patch_g * gpp;              // groups per image
#pragma unroll 2
              for (int r8 = 0; r8 < BM / 8; ++r8) {
                const int G = (m0 >> 3) + r8;
                const int img = G / gpi, rem = G - img * gpi;
                const int px = rem / gpp, pyg = rem - px * gpp;
                // groups past the last image have channel coordinate >= 3B: zero fill, byte count unchanged
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)  // sub-tile kk = kernel row 4 (kb % 4) + kk: [128 rows][32 B], atom r8
                  tma_load_4d_s(sa + kk * (BM * 32) + r8 * 256, &tmA, &full_bar[stage], px * 16, (kb & 3) * 4 + kk,
                                pyg * 8, img * 3 + (kb >> 2));
              }
            } else if (!p.a_mn) {
              tma_load_2d_s(sa, &tmA, &full_bar[stage], k0, m0);  // box {64 k, a_rows (128) rows}
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)  // box {64 m, 64 k-rows} per 64-wide chunk
                tma_load_2d_s(sa + c * (BK * 128), &tmA, &full_bar[stage], m0 + c * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_s(sb, &tmB, &full_bar[stage], k0, n0);  // box {64 k, BN rows}
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d_s(sb + c * (BK * 128), &tmB, &full_bar[stage], n0 + c * 64, k0);
            }
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) loop and the waits; one elected lane issues the tcgen05
    // instructions.  With a `lane == 0` branch around the loop every operand of UTCHMMA / UTCBAR lives in
    // per-thread registers and the compiler wraps each instruction in an ELECT / R2UR / BRA.U.ANY waterfall
    // loop (~80-100 issue cycles per MMA, close to the 128 cycles the MMA itself takes); here the operands
    // are provably uniform (kernel parameters, loop counters, broadcast shuffles) and descriptors are a base
    // plus an immediate.
    {
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t idesc = make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      // Descriptor template per operand; only the start-address field changes per stage / k-step.
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tm + acc_stage * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem0 + stage * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          // patch mode: A is four SW32 sub-tiles of [128 rows][32 bytes], one per k-step
          const uint64_t adesc = p.patch_g ? make_smem_desc_sw32(sa, 256) : make_smem_desc_sw128(sa, a_lbo, 1024);
          const uint32_t a_step = p.patch_g ? (BM * 32) >> 4 : a_kstep;
          const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_bf16(d_tmem, adesc + k * a_step, bdesc + k * b_kstep, idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
            if (kb == k_blocks - 1) umma_commit(&tmem_full[acc_stage]);  // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> regs -> global =====================
    const int ew = (warp - 4) & 3;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int eh = (warp - 4) >> 2;  // column half of the tile
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb, nb;
      tile_coords(tile, m_blocks, n_blocks, p.group_m, mb, nb);
      int row = mb * BM + ew * 32 + lane;
      if (p.patch_g) {  // (image, px, py-group, py) -> image * g^2 + py * g + px; rows of padded groups are dropped
        const int g = p.patch_g, gpp = (g + 7) >> 3, gpi = g * gpp;
        const int G = row >> 3, img = G / gpi, rem = G - img * gpi;
        const int px = rem / gpp, py = (rem - px * gpp) * 8 + (row & 7);
        row = py < g ? (img * g + py) * g + px : p.M;
      }
      const int n0 = nb * BN;
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc_stage * BN;
      epilogue_rows<BN>(p, t_row, row, n0, eh, ((int)blockDim.x >> 5) / 4 - 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each
// CTA stages its own 128 rows of A and its own 128 of the 256 B rows, so a k-block costs 32 KB of
// shared memory per CTA instead of 48 KB -- six pipeline stages instead of four and half the B
// traffic per SM.  The leader CTA's single thread issues tcgen05.mma.cta_group::2 (M = 256); the
// accumulator rows of each CTA land in that CTA's own TMEM, so both CTAs run the same epilogue.
// Barriers: full[] lives in the leader (both CTAs' TMA bytes are signalled there); empty[] and
// tmem_full[] are arrived on in both CTAs by a multicast tcgen05.commit; tmem_empty[] lives in the
// leader and counts the 16 epilogue warps of the pair.
// ------------------------------------------------------------------------------------------------
struct Gemm2Smem {
  static constexpr int BN = 256;
  static constexpr int BNH = 128;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
};

// DYN: tiles are claimed at run time instead of walked in a fixed stride -- a CTA pair that shares its SMs with a
// collective or optimiser CTA simply takes fewer tiles.  The leader's producer warp draws tile ids from a global
// counter (one atomicAdd per tile, issued a whole tile ahead of its use) and publishes them through a 4-entry ring in
// BOTH CTAs' shared memory (st.shared::cluster + remote mbarrier arrive); the peer's producer, the MMA issuer and the
// epilogue warps read the ring in order and release each entry on the leader's `sched_empty` barrier; id -1 ends the
// stream.  The leader that draws the last sentinel resets the counter for the next launch that uses the same word.
constexpr int SCHED_RING = 4;
__device__ unsigned int g_tile_ctr[64];

template <bool DYN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmParams p) {
  using S = Gemm2Smem;
  constexpr int BN = S::BN, BNH = S::BNH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + S::STAGES;
  uint64_t* tmem_full = bars + 2 * S::STAGES;
  uint64_t* tmem_empty = bars + 2 * S::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
  uint64_t* sched_full = bars + 2 * S::STAGES + 6;               // [SCHED_RING], one arrival (the claimer) per entry
  uint64_t* sched_empty = sched_full + SCHED_RING;              // [SCHED_RING], leader's copy counts every reader of the pair
  volatile int* sched_tile = reinterpret_cast<volatile int*>(sched_empty + SCHED_RING);  // [SCHED_RING]
  static_assert((2 * S::STAGES + 6 + 2 * SCHED_RING) * 8 + SCHED_RING * 4 <= S::BAR_BYTES, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int m_blocks = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = p.swiglu_F ? p.swiglu_F / BNH : (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks;
  const int first_tile = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * ((blockDim.x >> 5) - 4));  // epilogue warps of both CTAs
    }
    if (DYN) {
      for (int s = 0; s < SCHED_RING; ++s) {
        mbar_init(&sched_full[s], 1);
        // readers of an entry: the peer's producer warp, the leader's MMA warp, the epilogue warps of both CTAs
        mbar_init(&sched_empty[s], 2 + 2 * ((blockDim.x >> 5) - 4));
      }
    }
    fence_mbar_init();
  }
  cluster_sync_all();  // peer barriers are initialised before anyone signals them
  // ring reader (every lane of the calling warp): next tile id of the stream, -1 at its end
  auto ring_next = [&](int& slot, uint32_t& ph) -> int {
    mbar_wait_cluster(&sched_full[slot], ph);
    const int t = sched_tile[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(mapa_u32(&sched_empty[slot], 0));
    if (++slot == SCHED_RING) { slot = 0; ph ^= 1; }
    return t;
  };
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 2 * BN);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; whole warp walks the loop, one elected lane issues) =====================
    {
      const bool issuer = elect_one();
      const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      // tile stream of this warp: fixed stride, or (DYN) claimed by the leader / read from the ring by the peer
      int rslot = 0;
      uint32_t rphase = 0;
      const int n_pairs = gridDim.x >> 1;
      auto claim = [&]() -> int {  // warp-uniform; the counter word returns to 0 after the last sentinel draw
        unsigned int t = 0;
        if (lane == 0) {
          t = atomicAdd(p.tile_ctr, 1u);
          if (t == (unsigned int)(num_tiles + n_pairs - 1)) *p.tile_ctr = 0u;
        }
        t = __shfl_sync(0xffffffffu, t, 0);
        return t < (unsigned int)num_tiles ? (int)t : -1;
      };
      int tile = first_tile, next_claim = -1;
      if (DYN) tile = leader ? claim() : ring_next(rslot, rphase);
      for (;;) {
        if (DYN) {
          if (leader) {  // publish the tile (or the end marker) to every reader of the pair, then draw the next one
            mbar_wait(&sched_empty[rslot], rphase ^ 1);
            if (lane == 0) {
              sched_tile[rslot] = tile;
              st_shared_cluster_u32(mapa_u32(const_cast<int*>(&sched_tile[rslot]), 1), (uint32_t)tile);
              mbar_arrive(&sched_full[rslot]);
              mbar_arrive_cluster(mapa_u32(&sched_full[rslot], 1));
            }
            __syncwarp();
            if (++rslot == SCHED_RING) { rslot = 0; rphase ^= 1; }
            if (tile >= 0) next_claim = claim();  // in flight while this tile's operands are loaded
          }
          if (tile < 0) break;
        } else if (tile >= num_tiles) {
          break;
        }
        int mb, nb;
        tile_coords(tile, m_blocks, n_blocks, p.group_m, mb, nb);
        const int m0 = mb * (2 * BM) + rank * BM;
        // B rows of this CTA: its half of the 256-wide tile, or (gated-SiLU mode) 128 rows of w1 / w3
        const int n0 = p.swiglu_F ? nb * BNH + rank * p.swiglu_F : nb * BN + rank * BNH;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sa = smem0 + stage * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          const uint32_t fb = mapa_u32(&full_bar[stage], 0);
          const int k0 = kb * BK;
          if (issuer) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d_2sm_h(sa, &tmA, fb, k0, m0, p.hint_a);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)
                tma_load_2d_2sm_h(sa + c * (BK * 128), &tmA, fb, m0 + c * 64, k0, p.hint_a);
            }
            if (!p.b_mn) {
              tma_load_2d_2sm_h(sb, &tmB, fb, k0, n0, p.hint_b);
            } else {
#pragma unroll
              for (int c = 0; c < BNH / 64; ++c)
                tma_load_2d_2sm_h(sb + c * (BK * 128), &tmB, fb, n0 + c * 64, k0, p.hint_b);
            }
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        if (DYN) tile = leader ? next_claim : ring_next(rslot, rphase);
        else tile += tile_step;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp, one elected lane issues) =====================
    if (leader) {
      const bool issuer = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t idesc = make_idesc_bf16(2 * BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      int rslot = 0;
      uint32_t rphase = 0;
      for (int tile = first_tile;; tile += tile_step) {
        if (DYN) {
          if (ring_next(rslot, rphase) < 0) break;
        } else if (tile >= num_tiles) {
          break;
        }
        mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tm + acc_stage * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem0 + stage * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_bf16_2sm(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 3);  // frees this stage in both CTAs
            if (kb == k_blocks - 1) umma_commit_2sm(&tmem_full[acc_stage], 3);
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of the 256-row tile =====================
    const int ew = (warp - 4) & 3, eh = (warp - 4) >> 2;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    int rslot = 0;
    uint32_t rphase = 0;
    for (int tile = first_tile;; tile += tile_step) {
      if (DYN) {
        tile = ring_next(rslot, rphase);
        if (tile < 0) break;
      } else if (tile >= num_tiles) {
        break;
      }
      int mb, nb;
      tile_coords(tile, m_blocks, n_blocks, p.group_m, mb, nb);
      const int row = mb * (2 * BM) + rank * BM + ew * 32 + lane;
      const int n0 = nb * BN;
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc_stage * BN;
      const int nhalves = ((int)blockDim.x >> 5) / 4 - 1;
      if (p.swiglu_F) epilogue_swiglu_rows(p, t_row, row, nb * BNH, eh * (4 / nhalves), (eh + 1) * (4 / nhalves));
      else epilogue_rows<BN>(p, t_row, row, n0, eh, nhalves);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc_stage], 0));
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ host side
// 384 threads (8 epilogue warps) when the epilogue is heavy relative to the main loop: activations,
// residual adds (wo +res 190 -> 170 us), the fused gate backward, or short reductions (K <= 2048, the
// ViT tower); 256 threads otherwise.
static int epilogue_threads(const GemmParams& p) {
  static const int force = getenv("A3_GEMM_EPI_WARPS") ? atoi(getenv("A3_GEMM_EPI_WARPS")) : 0;  // 4 / 8: testing
  if (force == 4) return 256;
  if (force == 8) return 384;
  if (p.swiglu_F) return 256;  // the fused gate forward measured faster on 4 warps (978 vs 1097 us)
  return (p.act || p.dswiglu_gu || p.residual || p.K <= 2048) ? 384 : 256;
}

static int encode_operand_map(CUtensorMap* tm, const void* ptr, int rows_out, int k_red,
                              long long ld, int mn_major, int block_rows) {
  // K-major: memory [rows_out][k_red] -> dims {k_red, rows_out}, box {64, block_rows}
  // MN-major: memory [k_red][rows_out] -> dims {rows_out, k_red}, box {64, 64}
  if (!mn_major)
    return make_tmap_2d_bf16(tm, ptr, (uint64_t)k_red, (uint64_t)rows_out, (uint64_t)ld * 2, 64,
                             (uint32_t)block_rows);
  return make_tmap_2d_bf16(tm, ptr, (uint64_t)rows_out, (uint64_t)k_red, (uint64_t)ld * 2, 64, BK);
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t stream) {
  using S = GemmSmem<BN>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_bf16_kernel<BN>), S::TOTAL)) return rc;
  const int m_blocks = (p.M + BM - 1) / BM, n_blocks = (p.N + BN - 1) / BN;
  const int tiles = m_blocks * n_blocks;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_kernel<BN><<<grid, epilogue_threads(p), S::TOTAL, stream>>>(tmA, tmB, p);
  return check_launch("a3_gemm_bf16");
}

static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                        cudaStream_t stream) {
  using S = Gemm2Smem;
  static const bool dynamic = getenv("A3_GEMM_DYNAMIC") && atoi(getenv("A3_GEMM_DYNAMIC")) == 1;
  const int m_blocks = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = p.swiglu_F ? p.swiglu_F / S::BNH : (p.N + S::BN - 1) / S::BN;
  const int tiles = m_blocks * n_blocks;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  if (dynamic) {
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm2_bf16_kernel<true>), S::TOTAL)) return rc;
    // one of 64 self-resetting counter words, round-robin: a word is reused 64 launches later, long after its reset
    static std::atomic<unsigned int> seq{0};
    unsigned int* base = nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_tile_ctr) != cudaSuccess)
      return set_error(A3_ERR_CUDA, "a3_gemm_bf16: tile counter symbol");
    GemmParams q = p;
    q.tile_ctr = base + (seq.fetch_add(1, std::memory_order_relaxed) & 63u);
    gemm2_bf16_kernel<true><<<grid, epilogue_threads(p), S::TOTAL, stream>>>(tmA, tmB, q);
    return check_launch("a3_gemm_bf16(2cta, dynamic)");
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm2_bf16_kernel<false>), S::TOTAL)) return rc;
  gemm2_bf16_kernel<false><<<grid, epilogue_threads(p), S::TOTAL, stream>>>(tmA, tmB, p);
  return check_launch("a3_gemm_bf16(2cta)");
}

}  // namespace a3

extern "C" int a3_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K,
                            long long lda, long long ldb, long long ldc, int a_mn_major,
                            int b_mn_major, const void* bias, const void* residual,
                            long long ldr, int act, int out_fp32, int accumulate, float alpha,
                            const void* rope_table, int rope_cols, int rope_period, int rope_pos0,
                            void* stream) {
  using namespace a3;
  if (M <= 0 || N <= 0 || K <= 0) return A3_OK;  // empty problem: nothing to do
  if (!A || !B || !C) return set_error(A3_ERR_ARG, "a3_gemm_bf16: null operand");
  if ((N & 7) || (lda & 7) || (ldb & 7) || (ldc & (out_fp32 ? 3 : 7)) ||
      (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(C) & 15))
    return set_error(A3_ERR_ARG,
                     "a3_gemm_bf16: N, lda, ldb, ldc must be multiples of 8 elements and "
                     "pointers 16-byte aligned");
  if (residual && ((ldr & 7) || (reinterpret_cast<uintptr_t>(residual) & 15)))
    return set_error(A3_ERR_ARG, "a3_gemm_bf16: residual must be 16-byte aligned, ldr % 8 == 0");
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15))
    return set_error(A3_ERR_ARG, "a3_gemm_bf16: bias must be 16-byte aligned");
  if (act < 0 || act > 2) return set_error(A3_ERR_ARG, "a3_gemm_bf16: act must be 0, 1 or 2");
  if (rope_table && rope_cols > 0 && ((rope_cols & 127) || out_fp32))
    return set_error(A3_ERR_ARG, "a3_gemm_bf16: rope needs rope_cols % 128 == 0 and bf16 output");

  // Kernel choice by wave quantisation: both kernels retire one 128 x 256 output block per SM per
  // "tile time"; the CTA-pair kernel (256 x 256 per SM pair) runs ~8 % faster per tile (six smem
  // stages, half the B traffic) but has half as many scheduling slots, so narrow / short problems
  // (ViT at batch 8) stay on the single-CTA kernel when that needs fewer waves.
  static const int force = getenv("A3_GEMM_KERNEL") ? atoi(getenv("A3_GEMM_KERNEL")) : 0;  // 1 / 2: testing
  bool pair;
  if (force) {
    pair = force == 2;
  } else if (N < 256 || M < 256) {
    pair = false;
  } else {
    const long long sms = num_sms();
    const long long tp = (long long)((M + 255) / 256) * ((N + 255) / 256);
    const long long ts = (long long)((M + 127) / 128) * ((N + 255) / 256);
    const double t_pair = (double)((tp + sms / 2 - 1) / (sms / 2)) / 1.08;
    const double t_single = (double)((ts + sms - 1) / sms);
    pair = t_pair < t_single;
  }
  int bn = pair ? 128 : ((N <= 128 || (N % 256 != 0 && N % 128 == 0 && N < 2048)) ? 128 : 256);
  int a_rows = BM;
  if (!pair && M <= 64 && !a_mn_major && !b_mn_major) {
    // decode-time shapes (one short M block): HBM-bound on the weight stream, so what matters is
    // enough CTAs pulling on HBM -> narrow N tiles, and only the live rows of A staged per k-block
    const long long sms = num_sms();
    bn = (N / 64 >= 2 * sms) ? 64 : 32;
    a_rows = (M + 7) & ~7;
  }
  CUtensorMap tmA, tmB;
  int rc = encode_operand_map(&tmA, A, M, K, lda, a_mn_major, a_rows);
  if (rc != A3_OK) return rc;
  rc = encode_operand_map(&tmB, B, N, K, ldb, b_mn_major, bn);  // pair: each CTA loads 128 B rows
  if (rc != A3_OK) return rc;

  GemmParams p;
  p.C = C;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.rope = reinterpret_cast<const float2*>(rope_table);
  p.rope_cols = rope_table ? rope_cols : 0;
  p.rope_period = rope_period > 0 ? rope_period : 1;
  p.rope_pos0 = rope_pos0;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldc = ldc;
  p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn_major ? 1 : 0;
  p.b_mn = b_mn_major ? 1 : 0;
  p.act = act;
  p.out_fp32 = out_fp32 ? 1 : 0;
  p.accumulate = accumulate ? 1 : 0;
  p.alpha = alpha;
  static const int gm_env = getenv("A3_GEMM_GROUP_M") ? atoi(getenv("A3_GEMM_GROUP_M")) : 0;  // tuning experiments
  p.group_m = gm_env > 0 ? gm_env : 16;
  set_cache_hints(p);
  p.tile_ctr = nullptr;
  p.swiglu_F = 0; p.C2 = nullptr; p.ldc2 = 0; p.dswiglu_gu = nullptr; p.ldgu = 0;
  p.a_rows = a_rows;
  p.patch_g = 0; p.patch_p = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pair) return launch_gemm2(tmA, tmB, p, s);
  if (bn == 32) return launch_gemm<32>(tmA, tmB, p, s);
  if (bn == 64) return launch_gemm<64>(tmA, tmB, p, s);
  return bn == 128 ? launch_gemm<128>(tmA, tmB, p, s) : launch_gemm<256>(tmA, tmB, p, s);
}

static void init_params(a3::GemmParams& p, int M, int N, int K) {
  p.C = nullptr; p.bias = nullptr; p.residual = nullptr; p.rope = nullptr; p.ldc = 0; p.ldr = 0;
  p.rope_cols = 0; p.rope_period = 1; p.rope_pos0 = 0; p.M = M; p.N = N; p.K = K; p.a_mn = 0; p.b_mn = 0;
  p.act = 0; p.out_fp32 = 0; p.accumulate = 0; p.alpha = 1.f; p.group_m = 16; set_cache_hints(p); p.tile_ctr = nullptr; p.swiglu_F = 0; p.C2 = nullptr;
  p.ldc2 = 0; p.dswiglu_gu = nullptr; p.ldgu = 0; p.a_rows = a3::BM; p.patch_g = 0; p.patch_p = 0;
}

extern "C" int a3_gemm_swiglu_bf16(const void* X, const void* W13, void* GU, void* Y, int M, int F, int K,
                                   long long ldx, long long ldw, long long ldgu, long long ldy,
                                   void* stream) {
  using namespace a3;
  if (M <= 0 || F <= 0 || K <= 0) return A3_OK;
  if (!X || !W13 || !GU || !Y) return set_error(A3_ERR_ARG, "a3_gemm_swiglu_bf16: null operand");
  if ((F & 127) || (ldx & 7) || (ldw & 7) || (ldgu & 7) || (ldy & 7) ||
      ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W13) | reinterpret_cast<uintptr_t>(GU) |
        reinterpret_cast<uintptr_t>(Y)) & 15))
    return set_error(A3_ERR_ARG, "a3_gemm_swiglu_bf16: F % 128 == 0, strides % 8 == 0, 16-byte aligned pointers");
  CUtensorMap tmA, tmB;
  int rc = encode_operand_map(&tmA, X, M, K, ldx, 0, BM);
  if (rc != A3_OK) return rc;
  rc = encode_operand_map(&tmB, W13, 2 * F, K, ldw, 0, 128);
  if (rc != A3_OK) return rc;
  GemmParams p;
  init_params(p, M, 2 * F, K);
  p.C = GU; p.ldc = ldgu; p.C2 = reinterpret_cast<__nv_bfloat16*>(Y); p.ldc2 = ldy; p.swiglu_F = F;
  return launch_gemm2(tmA, tmB, p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int a3_gemm_dswiglu_bf16(const void* D, const void* W2, const void* GU, void* DGU, int M, int F,
                                    int Dm, long long ldd, long long ldw, long long ldgu, long long lddgu,
                                    void* stream) {
  using namespace a3;
  if (M <= 0 || F <= 0 || Dm <= 0) return A3_OK;
  if (!D || !W2 || !GU || !DGU) return set_error(A3_ERR_ARG, "a3_gemm_dswiglu_bf16: null operand");
  if ((F & 7) || (ldd & 7) || (ldw & 7) || (ldgu & 7) || (lddgu & 7) ||
      ((reinterpret_cast<uintptr_t>(D) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(GU) |
        reinterpret_cast<uintptr_t>(DGU)) & 15))
    return set_error(A3_ERR_ARG, "a3_gemm_dswiglu_bf16: F % 8 == 0, strides % 8 == 0, 16-byte aligned pointers");
  // dy[M][F] = D[M][Dm] * W2[Dm][F]: A K-major, B MN-major (W2 is stored [K = Dm][N = F])
  const bool pair = M >= 256 && F >= 256;
  CUtensorMap tmA, tmB;
  int rc = encode_operand_map(&tmA, D, M, Dm, ldd, 0, BM);
  if (rc != A3_OK) return rc;
  rc = encode_operand_map(&tmB, W2, F, Dm, ldw, 1, pair ? 128 : 256);
  if (rc != A3_OK) return rc;
  GemmParams p;
  init_params(p, M, F, Dm);
  p.b_mn = 1; p.C = DGU; p.ldc = lddgu;
  p.dswiglu_gu = reinterpret_cast<const __nv_bfloat16*>(GU); p.ldgu = ldgu;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return pair ? launch_gemm2(tmA, tmB, p, s) : launch_gemm<256>(tmA, tmB, p, s);
}

/* im2col-free patch embedding: out[B * g^2][D] = conv(image, w) with kernel = stride = patch, no bias
 * (open_clip conv1, reference LLM/llama_ens5.py:354-357).  image: the output of a3_vit_pad_cast, bf16
 * [B][3][S][g][16]; w: [D][3 * 16 * 16] bf16, the (D, 3, patch, patch) kernel zero-padded to 16 x 16 windows
 * (k = c * 256 + ky * 16 + kx). */
extern "C" int a3_vit_patch_embed(const void* image, const void* w, void* out, int B, int S, int patch, int D,
                                  long long ldo, void* stream) {
  using namespace a3;
  if (B <= 0) return A3_OK;
  if (!image || !w || !out) return set_error(A3_ERR_ARG, "a3_vit_patch_embed: null operand");
  if (patch <= 0 || patch > 16 || S % patch || (D & 7) || (ldo & 7) ||
      ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15))
    return set_error(A3_ERR_ARG, "a3_vit_patch_embed: patch <= 16, S % patch == 0, D % 8 == 0, 16-byte aligned pointers");
  const int g = S / patch;
  CUtensorMap tmA, tmB;
  const uint64_t rowb = (uint64_t)g * 16 * 2;  // bytes of one padded pixel row
  const uint64_t dims[4] = {(uint64_t)g * 16, (uint64_t)patch, (uint64_t)g, (uint64_t)3 * B};
  const uint64_t strides[3] = {rowb, (uint64_t)patch * rowb, (uint64_t)S * rowb};
  const uint32_t box[4] = {16, 1, 8, 1};
  int rc = make_tmap_4d_bf16(&tmA, image, dims, strides, box, /*swizzle32=*/1);
  if (rc != A3_OK) return rc;
  const int K = 3 * 256;
  const int bn = D <= 128 ? 128 : 256;
  rc = encode_operand_map(&tmB, w, D, K, K, 0, bn);
  if (rc != A3_OK) return rc;
  GemmParams p;
  init_params(p, B * g * g, D, K);
  p.C = out; p.ldc = ldo;
  p.patch_g = g; p.patch_p = patch;
  // grid: tiles over the group-padded rows
  const int m_rows = B * g * ((g + 7) / 8) * 8;
  const int tiles = ((m_rows + BM - 1) / BM) * ((D + bn - 1) / bn);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 128) {
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_bf16_kernel<128>), GemmSmem<128>::TOTAL)) return e;
    gemm_bf16_kernel<128><<<grid, 384, GemmSmem<128>::TOTAL, st>>>(tmA, tmB, p);
  } else {
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_bf16_kernel<256>), GemmSmem<256>::TOTAL)) return e;
    gemm_bf16_kernel<256><<<grid, 384, GemmSmem<256>::TOTAL, st>>>(tmA, tmB, p);
  }
  return check_launch("a3_vit_patch_embed");
}

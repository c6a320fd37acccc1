// Decode-step kernels (q_len = 1): weight-streaming small-M linear and split-KV single-query
// attention over the KV cache.  Both are HBM-bound (13.2 GB of weights per token for
// 7B): 16-byte loads, >= 8 loads in flight per lane, grids several CTAs per SM.
#include "common.cuh"
#include "ptx.cuh"

namespace a3 {

constexpr int SM_MAXM = 16;
constexpr int LIN_WARPS = 8;
constexpr int LIN_COLS_PER_WARP = 2;

// Weight rows are read exactly once per token: L1 no-allocate.  (An L2 evict-first cache-hint policy on these loads was
// measured: 2.99 -> 5.59 ms/token at batch 1 -- the hinted loads lose more in the memory pipeline than the policy saves.)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 x = unpack_bf16x2(aw[t]), y = unpack_bf16x2(bw[t]);
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
  }
  return s;
}

// Y[m][n] = sum_k X[m][k] W[n][k].  A warp owns LIN_COLS_PER_WARP consecutive output columns and
// streams their weight rows once; X (M <= 16 rows) is re-read through L1.
template <int M>
__global__ void __launch_bounds__(LIN_WARPS * 32)
linear_small_m_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ W,
                      void* __restrict__ Y, int N, int K, long long ldx, long long ldw,
                      long long ldy, const __nv_bfloat16* __restrict__ residual, long long ldr,
                      int out_fp32) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * LIN_WARPS + warp) * LIN_COLS_PER_WARP;
  if (n0 >= N) return;
  const int nchunks = K >> 3;
  float acc[LIN_COLS_PER_WARP][M];
#pragma unroll
  for (int c = 0; c < LIN_COLS_PER_WARP; ++c)
#pragma unroll
    for (int m = 0; m < M; ++m) acc[c][m] = 0.f;
  const __nv_bfloat16* w0 = W + (long long)n0 * ldw;
  const bool has1 = n0 + 1 < N;
  const __nv_bfloat16* w1 = has1 ? w0 + ldw : w0;
#pragma unroll 4
  for (int ch = lane; ch < nchunks; ch += 32) {
    const uint4 a = ldg_stream(w0 + ch * 8);
    const uint4 b = ldg_stream(w1 + ch * 8);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(X + m * ldx) + ch);
      acc[0][m] += dot8(a, x);
      acc[1][m] += dot8(b, x);
    }
  }
#pragma unroll
  for (int c = 0; c < LIN_COLS_PER_WARP; ++c)
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c][m] += __shfl_xor_sync(0xffffffffu, acc[c][m], o);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < LIN_COLS_PER_WARP; ++c) {
      const int n = n0 + c;
      if (n >= N) break;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        float v = acc[c][m];
        if (out_fp32) {
          // the reference's logits are a bf16 GEMM result cast with .float() (llama_ens5.py:530-531)
          v = bf16_round(v);
          if (residual) v += __bfloat162float(residual[m * ldr + n]);
          reinterpret_cast<float*>(Y)[m * ldy + n] = v;
        } else {
          if (residual) v = bf16_round(v) + __bfloat162float(residual[m * ldr + n]);
          reinterpret_cast<__nv_bfloat16*>(Y)[m * ldy + n] = __float2bfloat16_rn(v);
        }
      }
    }
  }
}


// ------------------------------------------------------------------ fused decode linear
// Y[m][n] = sum_k Xeff[m][k] W[n][k] for M <= 8 rows (decode, q_len = 1), one launch per projection:
//   prologue (once per CTA, into shared memory):  Xeff = X                       (plain)
//                                                 Xeff = RMSNorm(X) * norm_w     (components.py:39-53)
//                                                 Xeff = bf16(silu(g)) * u       (llama_ens5.py:213-214)
//   main loop: a warp owns two consecutive output columns and streams their weight rows once
//              (ld.global.nc, 16-byte, L1 no-allocate); HBM-bound
//   epilogue:  + residual | fp32 logits | RoPE on the (2j, 2j+1) pair the warp owns + KV-cache append
//              (llama_ens5.py:118-127) for the QKV projection.
// Programmatic dependent launch: the first weight chunks are fetched BEFORE griddepcontrol.wait, so the
// weight stream of projection i+1 starts while projection i drains; everything that depends on the
// previous kernel (X, the position) is read after the wait.
struct DecLinParams {
  const __nv_bfloat16* X; const __nv_bfloat16* W; void* Y;
  const __nv_bfloat16* norm_w; const __nv_bfloat16* residual;
  const float2* rope; __nv_bfloat16* kc; __nv_bfloat16* vc; const int* pos_dev;
  long long ldx, ldw, ldy, ldr;
  int N, K, mode /*0 plain, 1 rmsnorm, 2 swiglu*/, out_fp32, n_heads, n_kv_heads, pos, max_seq;
  int gate_F;  // > 0: W = w1 | w3 ([2F][K]); a warp owns ONE feature n (rows n and F + n) and writes
               // y[n] = bf16(silu(g)) * u -- the gated SiLU as the epilogue (llama_ens5.py:213-214); N = F tasks
  float eps;
};
constexpr int DL_PF = 8;  // prefetched 16-byte chunks per lane per weight row (issued before the PDL wait)


// ---- activation prologue shared by the decode projections (once per CTA, result in shared memory)
template <int M>
__device__ __forceinline__ void stage_activations(const DecLinParams& p, __nv_bfloat16* xs, float (*s_red)[M]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, nchunks = K >> 3;
  if (p.mode == 1) {
    // one global pass: raw rows -> smem while accumulating sum(x^2); the transform re-reads smem
    float ss[M];
#pragma unroll
    for (int m = 0; m < M; ++m) ss[m] = 0.f;
    for (int ch = threadIdx.x; ch < nchunks; ch += LIN_WARPS * 32) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint4 q = *reinterpret_cast<const uint4*>(p.X + m * p.ldx + ch * 8);
        *reinterpret_cast<uint4*>(xs + m * K + ch * 8) = q;
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w[t]); ss[m] += f.x * f.x + f.y * f.y; }
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss[m] += __shfl_xor_sync(0xffffffffu, ss[m], o);
      if (lane == 0) s_red[warp][m] = ss[m];
    }
    __syncthreads();
    float rstd[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LIN_WARPS; ++w) t += s_red[w][m];
      rstd[m] = rsqrtf(t / (float)K + p.eps);
    }
    for (int ch = threadIdx.x; ch < nchunks; ch += LIN_WARPS * 32) {  // same chunks this thread staged
      const uint4 nq = __ldg(reinterpret_cast<const uint4*>(p.norm_w + ch * 8));
      const uint32_t nw[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint4 q = *reinterpret_cast<const uint4*>(xs + m * K + ch * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16x2(w[t]), g = unpack_bf16x2(nw[t]);
          o[t] = pack_bf16x2(bf16_round(f.x * rstd[m]) * g.x, bf16_round(f.y * rstd[m]) * g.y);
        }
        *reinterpret_cast<uint4*>(xs + m * K + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  } else if (p.mode == 2) {
    for (int ch = threadIdx.x; ch < nchunks; ch += LIN_WARPS * 32) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint4 gq = *reinterpret_cast<const uint4*>(p.X + m * p.ldx + ch * 8);
        const uint4 uq = *reinterpret_cast<const uint4*>(p.X + m * p.ldx + K + ch * 8);
        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w}, uw[4] = {uq.x, uq.y, uq.z, uq.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 g = unpack_bf16x2(gw[t]), u = unpack_bf16x2(uw[t]);
          o[t] = pack_bf16x2(bf16_round(g.x / (1.f + __expf(-g.x))) * u.x,
                             bf16_round(g.y / (1.f + __expf(-g.y))) * u.y);
        }
        *reinterpret_cast<uint4*>(xs + m * K + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  } else {
    for (int ch = threadIdx.x; ch < nchunks; ch += LIN_WARPS * 32)
#pragma unroll
      for (int m = 0; m < M; ++m)
        *reinterpret_cast<uint4*>(xs + m * K + ch * 8) = *reinterpret_cast<const uint4*>(p.X + m * p.ldx + ch * 8);
  }
}

// ---- epilogue of one output pair (lane 0 of the owning warp): gate | RoPE + cache append | residual / logits
template <int M>
__device__ __forceinline__ void store_outputs(const DecLinParams& p, int n, const float (&acc)[LIN_COLS_PER_WARP][M]) {
    if (p.gate_F > 0) {  // both Linear outputs round to bf16, silu rounds, the product rounds
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float g = bf16_round(acc[0][m]), u = bf16_round(acc[1][m]);
        reinterpret_cast<__nv_bfloat16*>(p.Y)[m * p.ldy + n] = __float2bfloat16_rn(bf16_round(g / (1.f + __expf(-g))) * u);
      }
      return;
    }
    if (p.rope) {
      // QKV projection (columns q [0, D) | k [D, D + Dkv) | v): n is even, so this warp owns one RoPE pair (or a v pair)
      const int D = p.n_heads * 128, Dkv = p.n_kv_heads * 128;
      const int pos = p.pos_dev ? *p.pos_dev : p.pos;
      const int sect = n < D ? 0 : (n < D + Dkv ? 1 : 2);  // 0 q, 1 k, 2 v
      const int within = sect == 0 ? n : n - D - (sect - 1) * Dkv;
      const float2 cs = p.rope[(long long)pos * 64 + ((within & 127) >> 1)];
#pragma unroll
      for (int m = 0; m < M; ++m) {
        float a = bf16_round(acc[0][m]), b = bf16_round(acc[1][m]);
        if (sect < 2) { const float ra = a * cs.x - b * cs.y, rb = a * cs.y + b * cs.x; a = ra; b = rb; }
        const uint32_t r = pack_bf16x2(a, b);
        *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.Y) + m * p.ldy + n) = r;
        if (sect >= 1) {
          __nv_bfloat16* cache = sect == 1 ? p.kc : p.vc;
          *reinterpret_cast<uint32_t*>(cache + ((long long)m * p.max_seq + pos) * Dkv + within) = r;
        }
      }
      return;
    }
#pragma unroll
    for (int c = 0; c < LIN_COLS_PER_WARP; ++c) {
      const int nn = n + c;
      if (nn >= p.N) break;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        float v = acc[c][m];
        if (p.out_fp32) {
          v = bf16_round(v);  // logits: .float() of a bf16 GEMM result (llama_ens5.py:530-531)
          if (p.residual) v += __bfloat162float(p.residual[m * p.ldr + nn]);
          reinterpret_cast<float*>(p.Y)[m * p.ldy + nn] = v;
        } else {
          if (p.residual) v = bf16_round(v) + __bfloat162float(p.residual[m * p.ldr + nn]);
          reinterpret_cast<__nv_bfloat16*>(p.Y)[m * p.ldy + nn] = __float2bfloat16_rn(v);
        }
      }
    }
}

template <int M>
__global__ void __launch_bounds__(LIN_WARPS * 32)
decode_linear_kernel(const DecLinParams p) {
  extern __shared__ __align__(16) uint8_t dl_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(dl_smem);  // [M][K]
  __shared__ float s_red[LIN_WARPS][M];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = p.gate_F > 0 ? 1 : LIN_COLS_PER_WARP;  // output indices a warp advances by
  const int n0 = (blockIdx.x * LIN_WARPS + warp) * per_warp;
  const bool active = n0 < p.N;
  const int K = p.K, nchunks = K >> 3;
  const __nv_bfloat16* w0 = p.W + (long long)(active ? n0 : 0) * p.ldw;
  const __nv_bfloat16* w1 = p.gate_F > 0 ? w0 + (long long)p.gate_F * p.ldw : ((n0 + 1 < p.N) ? w0 + p.ldw : w0);
  uint4 pa[DL_PF], pb[DL_PF];
#pragma unroll
  for (int i = 0; i < DL_PF; ++i) {
    const int ch = lane + i * 32;
    if (ch < nchunks) { pa[i] = ldg_stream(w0 + ch * 8); pb[i] = ldg_stream(w1 + ch * 8); }
  }
  griddep_launch();
  griddep_wait();

  stage_activations<M>(p, xs, s_red);
  __syncthreads();

  // ---- persistent loop over column pairs: the activations are staged once per CTA, every warp walks
  // pairs n0, n0 + stride, ... ; the first pair's leading chunks were prefetched before the PDL wait
  const int stride = gridDim.x * LIN_WARPS * per_warp;
  bool first_pair = true;
  for (int n = n0; n < p.N; n += stride, first_pair = false) {
    const __nv_bfloat16* wa = p.W + (long long)n * p.ldw;
    const __nv_bfloat16* wb = p.gate_F > 0 ? wa + (long long)p.gate_F * p.ldw : ((n + 1 < p.N) ? wa + p.ldw : wa);
    float acc[LIN_COLS_PER_WARP][M];
#pragma unroll
    for (int c = 0; c < LIN_COLS_PER_WARP; ++c)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[c][m] = 0.f;
    if (!first_pair) {
#pragma unroll
      for (int i = 0; i < DL_PF; ++i) {
        const int ch = lane + i * 32;
        if (ch < nchunks) { pa[i] = ldg_stream(wa + ch * 8); pb[i] = ldg_stream(wb + ch * 8); }
      }
    }
#pragma unroll
    for (int i = 0; i < DL_PF; ++i) {
      const int ch = lane + i * 32;
      if (ch < nchunks) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const uint4 x = *reinterpret_cast<const uint4*>(xs + m * K + ch * 8);
          acc[0][m] += dot8(pa[i], x);
          acc[1][m] += dot8(pb[i], x);
        }
      }
    }
#pragma unroll 4
    for (int ch = lane + DL_PF * 32; ch < nchunks; ch += 32) {
      const uint4 a = ldg_stream(wa + ch * 8);
      const uint4 b = ldg_stream(wb + ch * 8);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint4 x = *reinterpret_cast<const uint4*>(xs + m * K + ch * 8);
        acc[0][m] += dot8(a, x);
        acc[1][m] += dot8(b, x);
      }
    }
#pragma unroll
    for (int c = 0; c < LIN_COLS_PER_WARP; ++c)
#pragma unroll
      for (int m = 0; m < M; ++m)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c][m] += __shfl_xor_sync(0xffffffffu, acc[c][m], o);
    if (lane == 0) store_outputs<M>(p, n, acc);
  }
}

// ------------------------------------------------------------------ single-query attention
// grid (B*H, nsplit).  Each CTA (4 warps) scans its slice of the context: a warp takes one key at
// a time, lanes hold 4 consecutive dims (head_dim 128).  Partial (m, l, acc[128]) per CTA goes to
// scratch; the combine kernel merges the splits.
constexpr int DEC_WARPS = 4;
constexpr int DEC_MAX_SPLIT = 32;

__global__ void __launch_bounds__(DEC_WARPS * 32)
attn_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                           const __nv_bfloat16* __restrict__ vc, float* __restrict__ scratch, int H, int Hkv,
                           int ctx, int max_seq, long long ldq, float scale_log2,
                           const int* __restrict__ pos_dev, int* __restrict__ counters,
                           __nv_bfloat16* __restrict__ o, long long ldo) {
  constexpr int HD = 128;
  __shared__ bool s_last;
  if (pos_dev) ctx = *pos_dev + 1;  // CUDA-graph replay: the context length lives in device memory
  __shared__ float s_acc[DEC_WARPS][HD];
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int nsplit = gridDim.y, split = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (ctx + nsplit - 1) / nsplit;
  const int k_begin = split * per, k_end = min(ctx, k_begin + per);
  const uint2 qraw = *reinterpret_cast<const uint2*>(q + b * ldq + h * HD + lane * 4);
  const float2 q01 = unpack_bf16x2(qraw.x), q23 = unpack_bf16x2(qraw.y);
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3v = 0.f;
  const long long base = ((long long)b * max_seq) * Hkv * HD + (h / (H / Hkv)) * HD + lane * 4;  // group head of h
  // 8 keys per warp in flight: all K / V loads of a group are issued before the dependent
  // online-softmax chain (one L2 round trip per group instead of per key)
  constexpr int KU = 8;
  for (int t0 = k_begin + warp; t0 < k_end; t0 += DEC_WARPS * KU) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL successor may start prefetching weights
    uint2 kr[KU], vr[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int t = t0 + u * DEC_WARPS;
      if (t < k_end) {
        const long long off = base + (long long)t * Hkv * HD;
        kr[u] = *reinterpret_cast<const uint2*>(kc + off);
        vr[u] = *reinterpret_cast<const uint2*>(vc + off);
      }
    }
    float sc[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const float2 k01 = unpack_bf16x2(kr[u].x), k23 = unpack_bf16x2(kr[u].y);
      sc[u] = q01.x * k01.x + q01.y * k01.y + q23.x * k23.x + q23.y * k23.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < KU; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      if (t0 + u * DEC_WARPS < k_end) {
        const float s = sc[u] * scale_log2;
        const float mn = fmaxf(m, s);
        const float corr = exp2f(m - mn), p = exp2f(s - mn);
        const float2 v01 = unpack_bf16x2(vr[u].x), v23 = unpack_bf16x2(vr[u].y);
        l = l * corr + p;
        a0 = a0 * corr + p * v01.x;
        a1 = a1 * corr + p * v01.y;
        a2 = a2 * corr + p * v23.x;
        a3v = a3v * corr + p * v23.y;
        m = mn;
      }
    }
  }
  s_acc[warp][lane * 4 + 0] = a0;
  s_acc[warp][lane * 4 + 1] = a1;
  s_acc[warp][lane * 4 + 2] = a2;
  s_acc[warp][lane * 4 + 3] = a3v;
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  __syncthreads();
  // merge the warps: thread d (< 128) owns one output dim
  const int d = threadIdx.x;
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < DEC_WARPS; ++w) M = fmaxf(M, s_m[w]);
  float L = 0.f, A = 0.f;
#pragma unroll
  for (int w = 0; w < DEC_WARPS; ++w) {
    const float c = s_m[w] == -INFINITY ? 0.f : exp2f(s_m[w] - M);
    L += s_l[w] * c;
    A += s_acc[w][d] * c;
  }
  float* out = scratch + ((long long)bh * nsplit + split) * (HD + 2);
  out[d] = A;
  if (d == 0) { out[HD] = M; out[HD + 1] = L; }
  // The last split of this (b, h) to finish merges all of them (no second launch).  The counter
  // resets itself for the next step.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(&counters[bh], 1);
    s_last = done == nsplit - 1;
    if (s_last) counters[bh] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile float* in = scratch + (long long)bh * nsplit * (HD + 2);
  float Mx = -INFINITY;
  for (int s = 0; s < nsplit; ++s) Mx = fmaxf(Mx, in[s * (HD + 2) + HD]);
  float Lx = 0.f, Ax = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = in[s * (HD + 2) + HD];
    const float c = ms == -INFINITY ? 0.f : exp2f(ms - Mx);
    Lx += in[s * (HD + 2) + HD + 1] * c;
    Ax += in[s * (HD + 2) + d] * c;
  }
  o[b * ldo + h * HD + d] = __float2bfloat16_rn(Ax / Lx);
}

__global__ void __launch_bounds__(128)
attn_decode_combine_kernel(const float* __restrict__ scratch, __nv_bfloat16* __restrict__ o, int H,
                           int nsplit, long long ldo) {
  constexpr int HD = 128;
  const int bh = blockIdx.x, b = bh / H, h = bh % H, d = threadIdx.x;
  const float* in = scratch + (long long)bh * nsplit * (HD + 2);
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, in[s * (HD + 2) + HD]);
  float L = 0.f, A = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = in[s * (HD + 2) + HD];
    const float c = ms == -INFINITY ? 0.f : exp2f(ms - M);
    L += in[s * (HD + 2) + HD + 1] * c;
    A += in[s * (HD + 2) + d] * c;
  }
  o[b * ldo + h * HD + d] = __float2bfloat16_rn(A / L);
}

// Fused RoPE (q, k) + KV-cache append for one new token per sequence (LLM/llama_ens5.py:118-127 with
// seqlen 1).  grid = B*H CTAs of 64 threads; thread j rotates pair j of q and k of head h in place,
// writes the rotated k and the v pair into cache[b, pos, h].  pos comes from device memory when
// pos_dev != null so the whole decode step can be replayed as one CUDA graph.
__global__ void __launch_bounds__(64)
decode_rope_kv_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, __nv_bfloat16* __restrict__ kc,
                      __nv_bfloat16* __restrict__ vc, int H, int Hkv, const float2* __restrict__ table,
                      const int* __restrict__ pos_dev, int pos_host, int max_seq) {
  constexpr int HD = 128;
  const int pos = pos_dev ? *pos_dev : pos_host;
  const int b = blockIdx.x / H, h = blockIdx.x % H, j = threadIdx.x;
  const long long D = (long long)H * HD, Dkv = (long long)Hkv * HD;
  __nv_bfloat16* row = qkv + b * ld + h * HD + 2 * j;
  const float2 cs = table[(long long)pos * (HD / 2) + j];
  {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL successor may start prefetching weights
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row));
    *reinterpret_cast<uint32_t*>(row) = pack_bf16x2(x.x * cs.x - x.y * cs.y, x.x * cs.y + x.y * cs.x);
  }
  if (h >= Hkv) return;  // the blocks of the first Hkv heads also rotate / append the k and v head of the same index
  const long long ci = (((long long)b * max_seq + pos) * Hkv + h) * HD + 2 * j;
  {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + D));
    const uint32_t r = pack_bf16x2(x.x * cs.x - x.y * cs.y, x.x * cs.y + x.y * cs.x);
    *reinterpret_cast<uint32_t*>(row + D) = r;
    *reinterpret_cast<uint32_t*>(kc + ci) = r;
  }
  *reinterpret_cast<uint32_t*>(vc + ci) = *reinterpret_cast<const uint32_t*>(row + D + Dkv);
}

}  // namespace a3

using namespace a3;
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

template <int M>
static int launch_small_m(const void* X, const void* W, void* Y, int N, int K, long long ldx,
                          long long ldw, long long ldy, const void* residual, long long ldr,
                          int out_fp32, cudaStream_t s) {
  const int cols_per_block = LIN_WARPS * LIN_COLS_PER_WARP;
  const int grid = (N + cols_per_block - 1) / cols_per_block;
  linear_small_m_kernel<M><<<grid, LIN_WARPS * 32, 0, s>>>(BF(X), BF(W), Y, N, K, ldx, ldw, ldy,
                                                           BF(residual), ldr, out_fp32);
  return check_launch("a3_linear_small_m");
}

extern "C" int a3_linear_small_m(const void* X, const void* W, void* Y, int M, int N, int K,
                                 long long ldx, long long ldw, long long ldy, const void* residual,
                                 long long ldr, int out_fp32, void* stream) {
  if (M <= 0 || N <= 0) return A3_OK;
  if (M > SM_MAXM) return set_error(A3_ERR_ARG, "a3_linear_small_m: M <= 16 (use a3_gemm_bf16)");
  if ((K & 7) || (ldx & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(X) & 15) ||
      (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(A3_ERR_ARG, "a3_linear_small_m: K, ldx, ldw % 8 == 0 and 16-byte alignment");
  cudaStream_t s = ST(stream);
#define CASE(m) case m: return launch_small_m<m>(X, W, Y, N, K, ldx, ldw, ldy, residual, ldr, out_fp32, s);
  switch (M) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
    CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
  }
#undef CASE
  return A3_ERR_ARG;
}


// Two alternatives to the register-streamed main loop were built and measured at batch 1 (same box, 7B, ctx ~900;
// profiles/r02_decode_stream_experiments.jsonl), both bit-identical in their results and both slower:
//   * per-warp shared-memory rings filled by 1-D bulk copies (cp.async.bulk + mbarrier), 2-4 KB per request:
//     4.0-5.9 ms/token against 2.89 -- small bulk requests sustain only 2.5-3.3 TB/s across the chip;
//   * per-lane rings filled by cp.async (LDGSTS.128), depth 6 / 10 / 12: 6.1 / 4.5 / 4.3 ms/token.
// The idea behind both (bytes in flight in shared memory instead of registers, so one CTA per SM suffices and the
// PDL successor can fill its ring beside it) stays open; plain 16-byte loads into registers remain the fastest
// way this chip streams a weight matrix from a CUDA-core kernel.
template <int M>
static int launch_decode_linear(const DecLinParams& p, cudaStream_t s) {
  const size_t smem = (size_t)M * p.K * 2;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&decode_linear_kernel<M>), smem)) return rc;
  const int cols_per_block = LIN_WARPS * (p.gate_F > 0 ? 1 : LIN_COLS_PER_WARP);
  int grid = (p.N + cols_per_block - 1) / cols_per_block;
  const int slots = num_sms() * 2;  // persistent: one resident wave (2 CTAs per SM), X staged once per CTA
  if (grid > slots) grid = slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LIN_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_linear_kernel<M>, p);
  if (e != cudaSuccess) return set_error(A3_ERR_CUDA, cudaGetErrorString(e));
  return check_launch("a3_decode_linear");
}

extern "C" int a3_decode_linear(const void* X, const void* W, void* Y, int M, int N, int K,
                                long long ldx, long long ldw, long long ldy, const void* norm_w,
                                float eps, int swiglu, const void* residual, long long ldr,
                                int out_fp32, const float* rope_table, int n_heads, int n_kv_heads, void* kcache,
                                void* vcache, const int* pos_dev, int pos, int max_seq,
                                void* stream) {
  if (M <= 0 || N <= 0) return A3_OK;
  if (M > 8) return set_error(A3_ERR_ARG, "a3_decode_linear: M <= 8 (use a3_linear_small_m / a3_gemm_bf16)");
  if ((K & 7) || (ldx & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(X) & 15) ||
      (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(A3_ERR_ARG, "a3_decode_linear: K, ldx, ldw % 8 == 0 and 16-byte alignment");
  if ((size_t)M * K * 2 > 200 * 1024) return set_error(A3_ERR_ARG, "a3_decode_linear: M*K too large for smem staging");
  if (norm_w && swiglu) return set_error(A3_ERR_ARG, "a3_decode_linear: norm and swiglu prologues are exclusive");
  if (rope_table && (n_kv_heads <= 0 || n_heads % n_kv_heads || N != (n_heads + 2 * n_kv_heads) * 128 || out_fp32 ||
                     residual || !kcache || !vcache || (ldy & 1)))
    return set_error(A3_ERR_ARG, "a3_decode_linear: rope epilogue needs N == (H + 2*Hkv)*128, H % Hkv == 0, bf16 out, caches");
  DecLinParams p;
  p.X = BF(X); p.W = BF(W); p.Y = Y; p.norm_w = BF(norm_w); p.residual = BF(residual);
  p.rope = reinterpret_cast<const float2*>(rope_table); p.kc = BFM(kcache); p.vc = BFM(vcache);
  p.pos_dev = pos_dev; p.ldx = ldx; p.ldw = ldw; p.ldy = ldy; p.ldr = ldr;
  p.N = N; p.K = K; p.mode = norm_w ? 1 : (swiglu ? 2 : 0); p.out_fp32 = out_fp32;
  p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.pos = pos; p.max_seq = max_seq; p.eps = eps; p.gate_F = 0;
  cudaStream_t s = ST(stream);
#define CASE(m) case m: return launch_decode_linear<m>(p, s);
  switch (M) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) }
#undef CASE
  return A3_ERR_ARG;
}

extern "C" int a3_decode_linear_gated(const void* X, const void* W13, void* Y, int M, int F, int K, long long ldx,
                                      long long ldw, long long ldy, const void* norm_w, float eps, void* stream) {
  if (M <= 0 || F <= 0) return A3_OK;
  if (M > 8) return set_error(A3_ERR_ARG, "a3_decode_linear_gated: M <= 8");
  if ((K & 7) || (ldx & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(X) & 15) ||
      (reinterpret_cast<uintptr_t>(W13) & 15))
    return set_error(A3_ERR_ARG, "a3_decode_linear_gated: K, ldx, ldw % 8 == 0 and 16-byte alignment");
  if ((size_t)M * K * 2 > 200 * 1024) return set_error(A3_ERR_ARG, "a3_decode_linear_gated: M*K too large for smem staging");
  DecLinParams p;
  p.X = BF(X); p.W = BF(W13); p.Y = Y; p.norm_w = BF(norm_w); p.residual = nullptr;
  p.rope = nullptr; p.kc = nullptr; p.vc = nullptr; p.pos_dev = nullptr;
  p.ldx = ldx; p.ldw = ldw; p.ldy = ldy; p.ldr = 0;
  p.N = F; p.K = K; p.mode = norm_w ? 1 : 0; p.out_fp32 = 0; p.n_heads = 0; p.n_kv_heads = 0; p.pos = 0; p.max_seq = 0; p.eps = eps;
  p.gate_F = F;
  cudaStream_t s = ST(stream);
#define CASE(m) case m: return launch_decode_linear<m>(p, s);
  switch (M) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) }
#undef CASE
  return A3_ERR_ARG;
}

extern "C" int a3_decode_rope_kv(void* qkv, long long ld, void* kcache, void* vcache, int B, int H, int Hkv,
                                 int head_dim, const float* table, int pos, const int* pos_dev,
                                 int max_seq, void* stream) {
  if (head_dim != 128) return set_error(A3_ERR_UNSUPPORTED, "a3_decode_rope_kv: head_dim must be 128");
  if ((ld & 1) || (!pos_dev && (pos < 0 || pos >= max_seq)) || Hkv <= 0 || H % Hkv)
    return set_error(A3_ERR_ARG, "a3_decode_rope_kv: 0 <= pos < max_seq, H % Hkv == 0");
  if (B * H <= 0) return A3_OK;
  decode_rope_kv_kernel<<<B * H, 64, 0, ST(stream)>>>(BFM(qkv), ld, BFM(kcache), BFM(vcache), H, Hkv,
                                                     reinterpret_cast<const float2*>(table), pos_dev, pos,
                                                     max_seq);
  return check_launch("a3_decode_rope_kv");
}

extern "C" int a3_attn_decode(const void* q, const void* kcache, const void* vcache, void* o,
                              float* scratch, int B, int H, int Hkv, int head_dim, int ctx, int max_seq,
                              long long ldq, long long ldo, float scale, const int* pos_dev,
                              void* stream) {
  if (head_dim != 128) return set_error(A3_ERR_UNSUPPORTED, "a3_attn_decode: head_dim must be 128");
  if (ctx <= 0 || ctx > max_seq) return set_error(A3_ERR_ARG, "a3_attn_decode: 0 < ctx <= max_seq");
  if (B * H <= 0) return A3_OK;
  if (Hkv <= 0 || H % Hkv) return set_error(A3_ERR_ARG, "a3_attn_decode: n_heads % n_kv_heads == 0");
  int nsplit = (2 * num_sms() + B * H - 1) / (B * H);
  const int max_useful = (ctx + 4 * DEC_WARPS - 1) / (4 * DEC_WARPS);
  if (!pos_dev && nsplit > max_useful) nsplit = max_useful;  // graph mode: split count must not depend on ctx
  if (nsplit > DEC_MAX_SPLIT) nsplit = DEC_MAX_SPLIT;
  if (nsplit < 1) nsplit = 1;
  attn_decode_partial_kernel<<<dim3(B * H, nsplit), DEC_WARPS * 32, 0, ST(stream)>>>(
      BF(q), BF(kcache), BF(vcache), scratch, H, Hkv, ctx, max_seq, ldq, scale * 1.4426950408889634f, pos_dev,
      reinterpret_cast<int*>(scratch + (long long)B * H * DEC_MAX_SPLIT * (128 + 2)), BFM(o), ldo);
  return check_launch("a3_attn_decode");
}

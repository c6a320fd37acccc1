// Decode-step kernels (q_len = 1): weight-streaming small-M linear and split-KV single-query
// attention over the KV cache.  Both are HBM-bound (13.2 GB of weights per token for
// 7B): 16-byte loads, >= 8 loads in flight per lane, grids several CTAs per SM.
#include "common.cuh"
#include "ptx.cuh"

namespace a3 {

constexpr int SM_MAXM = 16;
constexpr int LIN_WARPS = 8;
constexpr int LIN_COLS_PER_WARP = 2;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
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

// ------------------------------------------------------------------ single-query attention
// grid (B*H, nsplit).  Each CTA (4 warps) scans its slice of the context: a warp takes one key at
// a time, lanes hold 4 consecutive dims (head_dim 128).  Partial (m, l, acc[128]) per CTA goes to
// scratch; the combine kernel merges the splits.
constexpr int DEC_WARPS = 4;
constexpr int DEC_MAX_SPLIT = 32;

__global__ void __launch_bounds__(DEC_WARPS * 32)
attn_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                           const __nv_bfloat16* __restrict__ vc, float* __restrict__ scratch, int H,
                           int ctx, int max_seq, long long ldq, float scale_log2,
                           const int* __restrict__ pos_dev) {
  constexpr int HD = 128;
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
  const long long base = ((long long)b * max_seq) * H * HD + h * HD + lane * 4;
  for (int t = k_begin + warp; t < k_end; t += DEC_WARPS) {
    const long long off = base + (long long)t * H * HD;
    const uint2 kr = *reinterpret_cast<const uint2*>(kc + off);
    const uint2 vr = *reinterpret_cast<const uint2*>(vc + off);
    const float2 k01 = unpack_bf16x2(kr.x), k23 = unpack_bf16x2(kr.y);
    float s = q01.x * k01.x + q01.y * k01.y + q23.x * k23.x + q23.y * k23.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s *= scale_log2;
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    const float2 v01 = unpack_bf16x2(vr.x), v23 = unpack_bf16x2(vr.y);
    l = l * corr + p;
    a0 = a0 * corr + p * v01.x;
    a1 = a1 * corr + p * v01.y;
    a2 = a2 * corr + p * v23.x;
    a3v = a3v * corr + p * v23.y;
    m = mn;
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
                      __nv_bfloat16* __restrict__ vc, int H, const float2* __restrict__ table,
                      const int* __restrict__ pos_dev, int pos_host, int max_seq) {
  constexpr int HD = 128;
  const int pos = pos_dev ? *pos_dev : pos_host;
  const int b = blockIdx.x / H, h = blockIdx.x % H, j = threadIdx.x;
  const long long D = (long long)H * HD;
  __nv_bfloat16* row = qkv + b * ld + h * HD + 2 * j;
  const float2 cs = table[(long long)pos * (HD / 2) + j];
  const long long ci = (((long long)b * max_seq + pos) * H + h) * HD + 2 * j;
  {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row));
    *reinterpret_cast<uint32_t*>(row) = pack_bf16x2(x.x * cs.x - x.y * cs.y, x.x * cs.y + x.y * cs.x);
  }
  {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + D));
    const uint32_t r = pack_bf16x2(x.x * cs.x - x.y * cs.y, x.x * cs.y + x.y * cs.x);
    *reinterpret_cast<uint32_t*>(row + D) = r;
    *reinterpret_cast<uint32_t*>(kc + ci) = r;
  }
  *reinterpret_cast<uint32_t*>(vc + ci) = *reinterpret_cast<const uint32_t*>(row + 2 * D);
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

extern "C" int a3_decode_rope_kv(void* qkv, long long ld, void* kcache, void* vcache, int B, int H,
                                 int head_dim, const float* table, int pos, const int* pos_dev,
                                 int max_seq, void* stream) {
  if (head_dim != 128) return set_error(A3_ERR_UNSUPPORTED, "a3_decode_rope_kv: head_dim must be 128");
  if ((ld & 1) || (!pos_dev && (pos < 0 || pos >= max_seq)))
    return set_error(A3_ERR_ARG, "a3_decode_rope_kv: 0 <= pos < max_seq");
  if (B * H <= 0) return A3_OK;
  decode_rope_kv_kernel<<<B * H, 64, 0, ST(stream)>>>(BFM(qkv), ld, BFM(kcache), BFM(vcache), H,
                                                     reinterpret_cast<const float2*>(table), pos_dev, pos,
                                                     max_seq);
  return check_launch("a3_decode_rope_kv");
}

extern "C" int a3_attn_decode(const void* q, const void* kcache, const void* vcache, void* o,
                              float* scratch, int B, int H, int head_dim, int ctx, int max_seq,
                              long long ldq, long long ldo, float scale, const int* pos_dev,
                              void* stream) {
  if (head_dim != 128) return set_error(A3_ERR_UNSUPPORTED, "a3_attn_decode: head_dim must be 128");
  if (ctx <= 0 || ctx > max_seq) return set_error(A3_ERR_ARG, "a3_attn_decode: 0 < ctx <= max_seq");
  if (B * H <= 0) return A3_OK;
  int nsplit = (2 * num_sms() + B * H - 1) / (B * H);
  const int max_useful = (ctx + 4 * DEC_WARPS - 1) / (4 * DEC_WARPS);
  if (!pos_dev && nsplit > max_useful) nsplit = max_useful;  // graph mode: split count must not depend on ctx
  if (nsplit > DEC_MAX_SPLIT) nsplit = DEC_MAX_SPLIT;
  if (nsplit < 1) nsplit = 1;
  attn_decode_partial_kernel<<<dim3(B * H, nsplit), DEC_WARPS * 32, 0, ST(stream)>>>(
      BF(q), BF(kcache), BF(vcache), scratch, H, ctx, max_seq, ldq, scale * 1.4426950408889634f, pos_dev);
  int rc = check_launch("a3_attn_decode(partial)");
  if (rc != A3_OK) return rc;
  attn_decode_combine_kernel<<<B * H, 128, 0, ST(stream)>>>(scratch, BFM(o), H, nsplit, ldo);
  return check_launch("a3_attn_decode(combine)");
}

// Gather / scatter, ViT front end, loss, sampling and optimiser kernels of the hot path.  All are
// HBM-bound byte movers or row reductions: 16-byte vector accesses, one CTA per row or a
// grid-stride over flat ranges sized in multiples of the SM count.
#include <cstdio>
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace a3 {

__device__ __forceinline__ float warp_sum_m(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_m(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float block_sum_m(float v, float* red) {
  v = warp_sum_m(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
  return warp_sum_m(t);
}
__device__ __forceinline__ float block_max_m(float v, float* red) {
  v = warp_max_m(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (blockDim.x >> 5)) ? red[l] : -INFINITY;
  return warp_max_m(t);
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = unpack_bf16x2(w[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// ------------------------------------------------------------------ embedding gather + splice
// One CTA per output row of h [B][L][D], L = T + W, W = n_views * (Nimg + 2) (0 without image).
// Layout (llama_ens5.py:469-479): BOS | per view: start_img, Nimg tokens, end_img | text; the image
// tokens are view-major: img[(v * B + b) * Nimg + i]  (torch.chunk(image_feats, 5), llama_ens5.py:457).
__global__ void embed_splice_kernel(const long long* __restrict__ tokens,
                                    const __nv_bfloat16* __restrict__ E,
                                    const __nv_bfloat16* __restrict__ img,
                                    const __nv_bfloat16* __restrict__ start_img,
                                    const __nv_bfloat16* __restrict__ end_img,
                                    __nv_bfloat16* __restrict__ h, int B, int T, int Nimg, int D,
                                    long long ld_img, int n_views,
                                    const __nv_bfloat16* __restrict__ start2 = nullptr,
                                    const __nv_bfloat16* __restrict__ end2 = nullptr, int views_per_stream = 0) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL successor may start prefetching weights
  const int per = Nimg + 2;
  const int W = img ? n_views * per : 0;
  const int L = T + W;
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  const __nv_bfloat16* src;
  if (l == 0) src = E + tokens[(long long)b * T] * D;
  else if (l > W) src = E + tokens[(long long)b * T + (l - W)] * D;
  else {
    const int v = (l - 1) / per, r = (l - 1) % per;
    const bool second = start2 != nullptr && v >= views_per_stream;  // RGB-D: the depth stream's own tags
    if (r == 0) src = second ? start2 : start_img;
    else if (r == per - 1) src = second ? end2 : end_img;
    else src = img + (((long long)v * B + b) * Nimg + (r - 1)) * ld_img;
  }
  // note: for W == 0, l >= 1 > W takes the text branch.
  __nv_bfloat16* dst = h + (long long)blockIdx.x * D;
  for (int c = threadIdx.x; c < D / 8; c += blockDim.x)
    reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
}

// Backward.  Text rows scatter-add into dE with bf16x2 atomics (dE pre-zeroed by the caller);
// image rows are copied to d_img; start / end tags are summed over the batch in fp32.
__global__ void embed_splice_bwd_kernel(const long long* __restrict__ tokens,
                                        const __nv_bfloat16* __restrict__ dh,
                                        __nv_bfloat16* __restrict__ dE,
                                        __nv_bfloat16* __restrict__ d_img, int B, int T, int Nimg, int D,
                                        int has_img, int n_views) {
  const int per = Nimg + 2;
  const int W = has_img ? n_views * per : 0;
  const int L = T + W;
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  const __nv_bfloat16* src = dh + (long long)blockIdx.x * D;
  if (l == 0 || l > W) {
    const long long tok = tokens[(long long)b * T + (l == 0 ? 0 : l - W)];
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dE + tok * D);
    for (int c = threadIdx.x; c < D / 2; c += blockDim.x)
      atomicAdd(dst + c, reinterpret_cast<const __nv_bfloat162*>(src)[c]);
  } else {
    const int v = (l - 1) / per, r = (l - 1) % per;
    if (r > 0 && r < per - 1 && d_img) {
      __nv_bfloat16* dst = d_img + (((long long)v * B + b) * Nimg + (r - 1)) * D;
      for (int c = threadIdx.x; c < D / 8; c += blockDim.x)
        reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
    }
  }
}
// start_img / end_img are shared by every view and sample: sum their rows in fp32.
__global__ void tag_grad_kernel(const __nv_bfloat16* __restrict__ dh, float* __restrict__ d_start,
                                float* __restrict__ d_end, int B, int L, int per, int n_views, int D) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float s = 0.f, e = 0.f;
  for (int b = 0; b < B; ++b)
    for (int v = 0; v < n_views; ++v) {
      s += __bfloat162float(dh[((long long)b * L + 1 + v * per) * D + c]);
      e += __bfloat162float(dh[((long long)b * L + (v + 1) * per) * D + c]);
    }
  d_start[c] = s;
  d_end[c] = e;
}

// Reference tiling of a (2V x 2V) image into five V x V views (llama_ens5.py:383-385): view 0 = bicubic
// down-sample computed on the fp16-rounded image (F.interpolate(image.half(), "bicubic"): align_corners
// False, A = -0.75, fp32 accumulation, fp16 result), views 1..4 = the four quadrants.  Output fp32
// [5][B][3][V][V] (view-major, as torch.cat([image_224] + image_parts, dim=0)).
template <typename T>
__global__ void build_views_kernel(const T* __restrict__ image, float* __restrict__ out, int B, int V) {
  const int S = 2 * V;
  const long long total = 5LL * B * 3 * V * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % V), y = (int)((i / V) % V);
    const long long plane = (i / ((long long)V * V)) % (3LL * B);  // b * 3 + c
    const int v = (int)(i / ((long long)V * V * 3 * B));
    const T* src = image + plane * S * S;
    float r;
    if (v == 0) {
      // source coordinate 2*d + 0.5 -> taps 2d-1 .. 2d+2 (clamped), t = 0.5 in both directions
      const float w[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = min(max(2 * y - 1 + ky, 0), S - 1);
        float row = 0.f;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = min(max(2 * x - 1 + kx, 0), S - 1);
          row += w[kx] * __half2float(__float2half_rn((float)src[(long long)yy * S + xx]));
        }
        acc += w[ky] * row;
      }
      r = __half2float(__float2half_rn(acc));
    } else {
      const int oy = (v - 1) / 2 * V, ox = (v - 1) % 2 * V;
      r = (float)src[(long long)(oy + y) * S + ox + x];
    }
    out[i] = r;
  }
}

// ------------------------------------------------------------------ ViT front end
// A[(b, py, px)][k], k = c*P*P + ky*P + kx (conv1.weight.reshape(1024,-1) order), zero padded.
template <typename T>
__global__ void vit_patchify_kernel(const T* __restrict__ image, __nv_bfloat16* __restrict__ A,
                                    int S, int P, int Kpad) {
  const int g = S / P;
  const int b = blockIdx.x / (g * g);
  const int pp = blockIdx.x % (g * g);
  const int py = pp / g, px = pp % g;
  const int K = 3 * P * P;
  __nv_bfloat16* dst = A + (long long)blockIdx.x * Kpad;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < K) {
      const int c = k / (P * P), r = k % (P * P);
      const int ky = r / P, kx = r % P;
      v = (float)image[(((long long)b * 3 + c) * S + (py * P + ky)) * S + px * P + kx];
    }
    dst[k] = __float2bfloat16_rn(v);
  }
}

// t = bf16(patch_out + pos) (cls row: bf16(cls + pos[0])); y = bf16(LayerNorm(t)).  D <= 2048.
__global__ void __launch_bounds__(256)
vit_assemble_lnpre_kernel(const __nv_bfloat16* __restrict__ patch_out,
                          const __nv_bfloat16* __restrict__ cls,
                          const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ lw,
                          const __nv_bfloat16* __restrict__ lb, __nv_bfloat16* __restrict__ x, int Np,
                          int D, float eps) {
  __shared__ float red[32];
  const int N = Np + 1;
  const int b = blockIdx.x / N, t = blockIdx.x % N;
  const __nv_bfloat16* src = t == 0 ? cls : patch_out + ((long long)b * Np + (t - 1)) * D;
  const int c = threadIdx.x;
  const bool on = c < D / 8;
  float v[8];
  float s = 0.f;
  if (on) {
    float pv[8];
    ld8(src + c * 8, v);
    ld8(pos + (long long)t * D + c * 8, pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = bf16_round(v[j] + pv[j]);
      s += v[j];
    }
  }
  const float mean = block_sum_m(s, red) / (float)D;
  float ss = 0.f;
  if (on) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += (v[j] - mean) * (v[j] - mean);
  }
  const float rstd = rsqrtf(block_sum_m(ss, red) / (float)D + eps);
  if (on) {
    float wv[8], bv[8], o[8];
    ld8(lw + c * 8, wv);
    ld8(lb + c * 8, bv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * wv[j] + bv[j];
    st8(x + (long long)blockIdx.x * D + c * 8, o);
  }
}

// ------------------------------------------------------------------ loss
__global__ void count_nonzero_kernel(const long long* __restrict__ labels, int n,
                                     int* __restrict__ count) {
  __shared__ int red[32];
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += labels[i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *count = t;
  }
}

__global__ void shift_labels_kernel(const long long* __restrict__ labels, long long ld,
                                    long long* __restrict__ shifted, int B, int T,
                                    int* __restrict__ count) {
  __shared__ int red[32];
  int c = 0;
  for (int i = threadIdx.x; i < B * T; i += blockDim.x) {
    const int b = i / T, t = i - b * T;
    const long long v = t + 1 < T ? labels[b * ld + t + 1] : 0;
    shifted[i] = v;
    c += v != 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *count = t;
  }
}

// One CTA per row.  CrossEntropyLoss(ignore_index=0) in fp32 on bf16 logits; optionally
// overwrites the row with dlogits = (softmax - onehot) * grad_scale / count.
__global__ void __launch_bounds__(256)
cross_entropy_kernel(__nv_bfloat16* __restrict__ logits, long long ld,
                     const long long* __restrict__ labels, float* __restrict__ row_loss,
                     const int* __restrict__ count, int V, float grad_scale, int want_grad) {
  __shared__ float red[32];
  __shared__ float s_label_logit;
  const int row = blockIdx.x;
  __nv_bfloat16* lr = logits + row * ld;
  const long long label = labels[row];
  const int nchunks = V >> 3;
  if (label == 0) {  // ignored row: zero loss, zero gradient
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    if (want_grad) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int c = threadIdx.x; c < nchunks; c += blockDim.x) reinterpret_cast<uint4*>(lr)[c] = z;
    }
    return;
  }
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
    float v[8];
    ld8(lr + c * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, v[j]);
  }
  mx = block_max_m(mx, red);
  float sum = 0.f;
  for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
    float v[8];
    ld8(lr + c * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += __expf(v[j] - mx);
    // the label's logit is captured HERE, by the thread that owns its chunk: the gradient pass below overwrites the
    // row in place, so reading lr[label] after the reduction races with that store (found by compute-sanitizer:
    // the loss came out 1 % off when the instrumented timing let the store win)
    if (c == static_cast<int>(label >> 3)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j == static_cast<int>(label & 7)) s_label_logit = v[j];
    }
  }
  sum = block_sum_m(sum, red);  // (block barriers inside: s_label_logit is visible afterwards)
  const float lse = mx + logf(sum);
  if (threadIdx.x == 0) row_loss[row] = lse - s_label_logit;
  if (want_grad) {
    const float scale = grad_scale / (float)max(*count, 1);
    const float inv = 1.f / sum;
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
      float v[8], o[8];
      ld8(lr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pr = __expf(v[j] - mx) * inv;
        o[j] = (pr - ((long long)(c * 8 + j) == label ? 1.f : 0.f)) * scale;
      }
      st8(lr + c * 8, o);
    }
  }
}

__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int rows,
                                   const int* __restrict__ count, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) s += row_loss[i];
  s = block_sum_m(s, red);
  if (threadIdx.x == 0) out[0] = s / (float)max(*count, 1);
}

// argmax over fp32 rows, lowest index wins ties (torch.argmax on CUDA returns the first maximum).
// 1024 threads, four independent loads in flight per thread: the row (128 KB at V = 32000) is one
// latency-bound sweep (a 256-thread, one-load-at-a-time version took 26.7 us per decoded token).
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, long long ld, int V, long long* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* lr = logits + blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i0 = threadIdx.x; i0 < V; i0 += 4 * blockDim.x) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k * blockDim.x;
      v[k] = i < V ? __ldg(lr + i) : -INFINITY;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k * blockDim.x;
      if (i < V && (v[k] > best || (v[k] == best && i < bi))) { best = v[k]; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

// ------------------------------------------------------------------ optimiser
constexpr int SUMSQ_BLOCKS = 1184;  // 148 SMs x 8

// out[0] += sum(g^2).  Fixed grid, partials combined by the last CTA in index order -> the result
// is deterministic (no float atomics).  The partials and the arrival counter live in CALLER-OWNED
// scratch (A3_SUMSQ_SCRATCH_BYTES, zeroed once; the counter resets itself), so launches in flight on
// different streams never share state as long as each stream has its own scratch.
__global__ void __launch_bounds__(256)
sumsq_kernel(const __nv_bfloat16* __restrict__ g, long long n, float* __restrict__ out,
             float* __restrict__ g_sumsq_partial, unsigned int* __restrict__ done_ctr) {
  __shared__ float red[32];
  __shared__ bool last;
  float s = 0.f;
  const long long nchunks = n >> 3;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < nchunks;
       c += (long long)gridDim.x * blockDim.x) {
    float v[8];
    ld8(g + c * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j] * v[j];
  }
  if (blockIdx.x == 0)
    for (long long i = (nchunks << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = __bfloat162float(g[i]);
      s += v * v;
    }
  s = block_sum_m(s, red);
  if (threadIdx.x == 0) {
    g_sumsq_partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(done_ctr, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
      t += reinterpret_cast<volatile float*>(g_sumsq_partial)[i];
    t = block_sum_m(t, red);
    if (threadIdx.x == 0) {
      out[0] += t;
      *done_ctr = 0;
    }
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction) on fp32 masters with a
// bf16 shadow copy written in the same pass.  Gradient = bf16 grad * clip coefficient, where
// coef = min(1, max_norm / (sqrt(norm_sq) + 1e-6))  (util/clip_grad.py:187-193).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ master, __nv_bfloat16* __restrict__ wbf,
             const __nv_bfloat16* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1,
             float bc2_sqrt, const float* __restrict__ norm_sq, float max_norm) {
  float coef = 1.f;
  if (norm_sq && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*norm_sq) + 1e-6f));
  const float step_size = lr / bc1;
  const long long nchunks = n >> 3;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < nchunks;
       c += (long long)gridDim.x * blockDim.x) {
    float g[8], o[8];
    ld8(grad + c * 8, g);
    float4 p0 = reinterpret_cast<float4*>(master)[2 * c], p1 = reinterpret_cast<float4*>(master)[2 * c + 1];
    float4 m0 = reinterpret_cast<float4*>(m)[2 * c], m1 = reinterpret_cast<float4*>(m)[2 * c + 1];
    float4 v0 = reinterpret_cast<float4*>(v)[2 * c], v1 = reinterpret_cast<float4*>(v)[2 * c + 1];
    float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = g[j] * coef;
      pp[j] *= (1.f - lr * wd);
      mm[j] = beta1 * mm[j] + (1.f - beta1) * gj;
      vv[j] = beta2 * vv[j] + (1.f - beta2) * gj * gj;
      const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
      pp[j] -= step_size * (mm[j] / denom);
      o[j] = pp[j];
    }
    reinterpret_cast<float4*>(master)[2 * c] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(master)[2 * c + 1] = make_float4(pp[4], pp[5], pp[6], pp[7]);
    reinterpret_cast<float4*>(m)[2 * c] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(m)[2 * c + 1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
    reinterpret_cast<float4*>(v)[2 * c] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(v)[2 * c + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
    st8(wbf + c * 8, o);
  }
  // scalar tail (n % 8)
  if (blockIdx.x == 0) {
    for (long long i = (nchunks << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float gj = __bfloat162float(grad[i]) * coef;
      float p = master[i] * (1.f - lr * wd);
      const float mj = beta1 * m[i] + (1.f - beta1) * gj;
      const float vj = beta2 * v[i] + (1.f - beta2) * gj * gj;
      p -= step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
      master[i] = p; m[i] = mj; v[i] = vj;
      wbf[i] = __float2bfloat16_rn(p);
    }
  }
}

// ------------------------------------------------------------------ KV cache append
__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                 long long ld, __nv_bfloat16* __restrict__ kc,
                                 __nv_bfloat16* __restrict__ vc, int T, int HD, int start,
                                 int max_seq) {
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const long long src = (long long)blockIdx.x * ld;
  const long long dst = ((long long)b * max_seq + start + t) * HD;
  for (int c = threadIdx.x; c < HD / 8; c += blockDim.x) {
    reinterpret_cast<uint4*>(kc + dst)[c] = reinterpret_cast<const uint4*>(k + src)[c];
    reinterpret_cast<uint4*>(vc + dst)[c] = reinterpret_cast<const uint4*>(v + src)[c];
  }
}


// ------------------------------------------------------------------ grouped-query attention glue
// The reference materialises the repeated K / V heads for its SDPA branch (repeat_kv, absent LLM/llama.py; call
// llama_ens5.py:131-133): expanded head h = group head h / n_rep.  expand: compact [rows][D + 2 Dkv] (q | k | v with Hkv
// heads) -> expanded [rows][3 D]; reduce: its backward -- dq copied, dk / dv summed over the n_rep copies of a group
// (fp32, one rounding).  16-byte vectors; one thread per 8 output elements.
__global__ void __launch_bounds__(256)
repeat_kv_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long ld_out,
                 long long rows, int H, int Hkv, int hd) {
  const int D = H * hd, Dkv = Hkv * hd, n_rep = H / Hkv, vec_per_row = 3 * D / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int col = (int)(i - r * vec_per_row) * 8;
    const int sect = col / D, within = col - sect * D;
    const int src = sect == 0 ? col : D + (sect - 1) * Dkv + (within / hd / n_rep) * hd + within % hd;
    *reinterpret_cast<uint4*>(out + r * ld_out + col) = *reinterpret_cast<const uint4*>(in + r * ld_in + src);
  }
}

__global__ void __launch_bounds__(256)
reduce_kv_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long ld_out,
                 long long rows, int H, int Hkv, int hd) {
  const int D = H * hd, Dkv = Hkv * hd, n_rep = H / Hkv, vec_per_row = (D + 2 * Dkv) / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int col = (int)(i - r * vec_per_row) * 8;
    if (col < D) {
      *reinterpret_cast<uint4*>(out + r * ld_out + col) = *reinterpret_cast<const uint4*>(in + r * ld_in + col);
      continue;
    }
    const int sect = (col - D) / Dkv, within = (col - D) - sect * Dkv;  // 0 = k, 1 = v
    const int g = within / hd, e = within % hd;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int rep = 0; rep < n_rep; ++rep) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + r * ld_in + (sect + 1) * D + (g * n_rep + rep) * hd + e);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack_bf16x2(w[t]);
        acc[2 * t] += f.x;
        acc[2 * t + 1] += f.y;
      }
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + col) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

}  // namespace a3

using namespace a3;
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int a3_embed_splice(const long long* tokens, const void* E, const void* img,
                               const void* start_img, const void* end_img, void* h, int B, int T,
                               int Nimg, int D, long long ld_img, int n_views, void* stream) {
  if (D & 7) return set_error(A3_ERR_ARG, "a3_embed_splice: D % 8 != 0");
  if (img && (!start_img || !end_img)) return set_error(A3_ERR_ARG, "a3_embed_splice: tags missing");
  if (n_views < 1) n_views = 1;
  const int L = T + (img ? n_views * (Nimg + 2) : 0);
  if (B * L <= 0) return A3_OK;
  embed_splice_kernel<<<B * L, 128, 0, ST(stream)>>>(tokens, BF(E), BF(img), BF(start_img),
                                                     BF(end_img), BFM(h), B, T, Nimg, D, ld_img, n_views);
  return check_launch("a3_embed_splice");
}

extern "C" int a3_embed_splice_rgbd(const long long* tokens, const void* E, const void* img,
                                    const void* start_img, const void* end_img, const void* start_depth_img,
                                    const void* end_depth_img, void* h, int B, int T, int Nimg, int D,
                                    long long ld_img, int n_views, void* stream) {
  if (D & 7) return set_error(A3_ERR_ARG, "a3_embed_splice_rgbd: D % 8 != 0");
  if (!img || !start_img || !end_img || !start_depth_img || !end_depth_img)
    return set_error(A3_ERR_ARG, "a3_embed_splice_rgbd: image tokens and all four tags are required");
  if (n_views < 1) n_views = 1;
  const int L = T + 2 * n_views * (Nimg + 2);
  if (B * L <= 0) return A3_OK;
  embed_splice_kernel<<<B * L, 128, 0, ST(stream)>>>(tokens, BF(E), BF(img), BF(start_img), BF(end_img),
                                                     BFM(h), B, T, Nimg, D, ld_img, 2 * n_views,
                                                     BF(start_depth_img), BF(end_depth_img), n_views);
  return check_launch("a3_embed_splice_rgbd");
}

extern "C" int a3_embed_splice_bwd(const long long* tokens, const void* dh, void* dE, void* d_img,
                                   float* d_start_f32, float* d_end_f32, int B, int T, int Nimg,
                                   int D, int n_views, void* stream) {
  if (D & 7) return set_error(A3_ERR_ARG, "a3_embed_splice_bwd: D % 8 != 0");
  const int has_img = d_start_f32 != nullptr;
  if (n_views < 1) n_views = 1;
  const int W = has_img ? n_views * (Nimg + 2) : 0;
  const int L = T + W;
  if (B * L <= 0) return A3_OK;
  embed_splice_bwd_kernel<<<B * L, 128, 0, ST(stream)>>>(tokens, BF(dh), BFM(dE), BFM(d_img), B, T, Nimg,
                                                         D, has_img, n_views);
  int rc = check_launch("a3_embed_splice_bwd");
  if (rc != A3_OK || !has_img) return rc;
  tag_grad_kernel<<<(D + 255) / 256, 256, 0, ST(stream)>>>(BF(dh), d_start_f32, d_end_f32, B, L, Nimg + 2,
                                                           n_views, D);
  return check_launch("a3_embed_splice_bwd(tags)");
}

extern "C" int a3_build_views(const void* image, int image_is_fp32, float* views, int B, int V,
                              void* stream) {
  if (B <= 0 || V <= 0) return A3_OK;
  const long long total = 5LL * B * 3 * V * V;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  if (image_is_fp32)
    build_views_kernel<float><<<(int)blocks, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(image), views, B, V);
  else
    build_views_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, ST(stream)>>>(BF(image), views, B, V);
  return check_launch("a3_build_views");
}

extern "C" int a3_vit_patchify(const void* image, int image_is_fp32, void* A, int B, int S,
                               int patch, int Kpad, void* stream) {
  if (patch <= 0 || S % patch || Kpad < 3 * patch * patch || (Kpad & 7))
    return set_error(A3_ERR_ARG, "a3_vit_patchify: S % patch == 0, Kpad >= 3*patch^2, Kpad % 8 == 0");
  const int g = S / patch;
  if (B * g * g <= 0) return A3_OK;
  if (image_is_fp32)
    vit_patchify_kernel<float><<<B * g * g, 128, 0, ST(stream)>>>(
        reinterpret_cast<const float*>(image), BFM(A), S, patch, Kpad);
  else
    vit_patchify_kernel<__nv_bfloat16><<<B * g * g, 128, 0, ST(stream)>>>(BF(image), BFM(A), S, patch, Kpad);
  return check_launch("a3_vit_patchify");
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long n8 = n >> 3;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n8; c += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * c], b = reinterpret_cast<const float4*>(in)[2 * c + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st8(out + c * 8, v);
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}

// NCHW image (fp32 or bf16) -> bf16 rows in which every `patch`-pixel group starts on a 16-element (32-byte)
// boundary: out[(b, c)][y][px][16], kx >= patch zero.  This is the autocast cast of the conv input; the padding is
// what lets a TMA box start at an arbitrary patch column (TMA needs a 16-byte aligned inner coordinate -- a raw
// row of 14-pixel patches is 28-byte periodic: measured, cp.async.bulk.tensor traps on it).
template <typename T>
__global__ void vit_pad_cast_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int S, int patch,
                                    long long rows) {
  const int g = S / patch;
  const long long total = rows * g;  // one thread per (row, px): 16 outputs
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / g;
    const int px = (int)(i - r * g);
    const T* src = in + r * S + px * patch;
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      lo[k] = k < patch ? (float)src[k] : 0.f;
      hi[k] = k + 8 < patch ? (float)src[k + 8] : 0.f;
    }
    __nv_bfloat16* dst = out + i * 16;
    st8(dst, lo);
    st8(dst + 8, hi);
  }
}

extern "C" int a3_vit_pad_cast(const void* image, int image_is_fp32, void* out, int B, int S, int patch, void* stream) {
  if (patch <= 0 || patch > 16 || S % patch) return set_error(A3_ERR_ARG, "a3_vit_pad_cast: patch <= 16, S % patch == 0");
  if (B <= 0) return A3_OK;
  const long long rows = (long long)B * 3 * S, total = rows * (S / patch);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  if (image_is_fp32)
    vit_pad_cast_kernel<float><<<(int)blocks, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(image), BFM(out), S, patch, rows);
  else
    vit_pad_cast_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, ST(stream)>>>(BF(image), BFM(out), S, patch, rows);
  return check_launch("a3_vit_pad_cast");
}

extern "C" int a3_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  if (n <= 0) return A3_OK;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(A3_ERR_ARG, "a3_cast_f32_bf16: 16-byte aligned pointers");
  long long blocks = ((n >> 3) + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(in, BFM(out), n);
  return check_launch("a3_cast_f32_bf16");
}


static int gqa_args_ok(const char* who, const void* a, const void* b, long long lda, long long ldb, long long rows, int H,
                       int Hkv, int hd) {
  if (rows < 0 || H <= 0 || Hkv <= 0 || H % Hkv || hd <= 0 || (hd & 7) || (lda & 7) || (ldb & 7) ||
      ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15)) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "%s: n_heads %% n_kv_heads == 0, head_dim %% 8 == 0, strides %% 8 == 0, 16-byte aligned pointers", who);
    return set_error(A3_ERR_ARG, buf);
  }
  return A3_OK;
}

extern "C" int a3_repeat_kv(const void* qkv_compact, long long ld_compact, void* qkv, long long ld, long long rows,
                            int n_heads, int n_kv_heads, int head_dim, void* stream) {
  if (int rc = gqa_args_ok("a3_repeat_kv", qkv_compact, qkv, ld_compact, ld, rows, n_heads, n_kv_heads, head_dim)) return rc;
  if (rows == 0) return A3_OK;
  const long long vec = rows * (3LL * n_heads * head_dim / 8);
  long long blocks = (vec + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  repeat_kv_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(BF(qkv_compact), ld_compact, BFM(qkv), ld, rows, n_heads, n_kv_heads, head_dim);
  return check_launch("a3_repeat_kv");
}

extern "C" int a3_reduce_kv(const void* dqkv, long long ld, void* dqkv_compact, long long ld_compact, long long rows,
                            int n_heads, int n_kv_heads, int head_dim, void* stream) {
  if (int rc = gqa_args_ok("a3_reduce_kv", dqkv, dqkv_compact, ld, ld_compact, rows, n_heads, n_kv_heads, head_dim)) return rc;
  if (rows == 0) return A3_OK;
  const long long vec = rows * ((long long)(n_heads + 2 * n_kv_heads) * head_dim / 8);
  long long blocks = (vec + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  reduce_kv_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(BF(dqkv), ld, BFM(dqkv_compact), ld_compact, rows, n_heads, n_kv_heads, head_dim);
  return check_launch("a3_reduce_kv");
}

extern "C" int a3_vit_assemble_lnpre(const void* patch_out, const void* cls, const void* pos,
                                     const void* ln_w, const void* ln_b, void* x, int B, int Np, int D,
                                     float eps, void* stream) {
  if ((D & 7) || D > 2048) return set_error(A3_ERR_ARG, "a3_vit_assemble_lnpre: D % 8 == 0, D <= 2048");
  if (B * (Np + 1) <= 0) return A3_OK;
  vit_assemble_lnpre_kernel<<<B * (Np + 1), 256, 0, ST(stream)>>>(BF(patch_out), BF(cls), BF(pos),
                                                                  BF(ln_w), BF(ln_b), BFM(x), Np, D, eps);
  return check_launch("a3_vit_assemble_lnpre");
}

extern "C" int a3_count_nonzero(const long long* labels, int n, int* count, void* stream) {
  count_nonzero_kernel<<<1, 1024, 0, ST(stream)>>>(labels, n, count);
  return check_launch("a3_count_nonzero");
}

extern "C" int a3_shift_labels(const long long* labels, long long ld_labels, long long* shifted, int B,
                               int T, int* count, void* stream) {
  if (B <= 0 || T <= 0) return set_error(A3_ERR_ARG, "a3_shift_labels: empty batch");
  shift_labels_kernel<<<1, 1024, 0, ST(stream)>>>(labels, ld_labels, shifted, B, T, count);
  return check_launch("a3_shift_labels");
}

extern "C" int a3_cross_entropy(void* logits, long long ld, const long long* labels, float* row_loss,
                                const int* count, int rows, int V, float grad_scale, int want_grad,
                                void* stream) {
  if ((V & 7) || (ld & 7)) return set_error(A3_ERR_ARG, "a3_cross_entropy: V % 8 == 0 and ld % 8 == 0");
  if (rows <= 0) return A3_OK;
  cross_entropy_kernel<<<rows, 256, 0, ST(stream)>>>(BFM(logits), ld, labels, row_loss, count, V,
                                                     grad_scale, want_grad);
  return check_launch("a3_cross_entropy");
}

extern "C" int a3_loss_reduce(const float* row_loss, int rows, const int* count, float* out,
                              void* stream) {
  loss_reduce_kernel<<<1, 1024, 0, ST(stream)>>>(row_loss, rows, count, out);
  return check_launch("a3_loss_reduce");
}

extern "C" int a3_argmax(const float* logits, long long ld, int B, int V, long long* out, void* stream) {
  if (B <= 0) return A3_OK;
  argmax_kernel<<<B, 1024, 0, ST(stream)>>>(logits, ld, V, out);
  return check_launch("a3_argmax");
}

extern "C" int a3_sumsq_bf16(const void* g, long long n, float* out, void* scratch, void* stream) {
  static_assert((SUMSQ_BLOCKS + 1) * 4 <= A3_SUMSQ_SCRATCH_BYTES, "scratch too small");
  if (n <= 0) return A3_OK;
  if (reinterpret_cast<uintptr_t>(g) & 15) return set_error(A3_ERR_ARG, "a3_sumsq_bf16: 16-byte alignment");
  if (!scratch || (reinterpret_cast<uintptr_t>(scratch) & 3))
    return set_error(A3_ERR_ARG, "a3_sumsq_bf16: scratch (A3_SUMSQ_SCRATCH_BYTES, zero-initialised) required");
  long long blocks = ((n >> 3) + 255) / 256;
  if (blocks > SUMSQ_BLOCKS) blocks = SUMSQ_BLOCKS;
  if (blocks < 1) blocks = 1;
  float* part = reinterpret_cast<float*>(scratch);
  sumsq_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(BF(g), n, out, part,
                                                     reinterpret_cast<unsigned int*>(part + SUMSQ_BLOCKS));
  return check_launch("a3_sumsq_bf16");
}

// Resident AdamW CTAs per SM.  16 = the stand-alone optimum (6.1 TB/s).  The pipelined optimiser step
// (Engine.overlap_optimizer) runs AdamW on a side stream underneath the next step's forward: there a
// small grid (1-2 CTAs of 64 registers x 256 threads per SM) fits NEXT TO a resident 384-thread GEMM CTA
// instead of taking the SM's register file first.
static int g_adamw_blocks_per_sm = 16;
extern "C" int a3_set_adamw_blocks_per_sm(int n) {
  if (n < 1 || n > 32) return set_error(A3_ERR_ARG, "a3_set_adamw_blocks_per_sm: 1..32");
  g_adamw_blocks_per_sm = n;
  return A3_OK;
}

extern "C" int a3_adamw(float* master, void* weight_bf16, const void* grad_bf16, float* m, float* v,
                        long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int step, const float* norm_sq, float max_norm, void* stream) {
  if (n <= 0) return A3_OK;
  if (step < 1) return set_error(A3_ERR_ARG, "a3_adamw: step starts at 1");
  if ((reinterpret_cast<uintptr_t>(master) & 31) || (reinterpret_cast<uintptr_t>(m) & 31) ||
      (reinterpret_cast<uintptr_t>(v) & 31) || (reinterpret_cast<uintptr_t>(weight_bf16) & 15) ||
      (reinterpret_cast<uintptr_t>(grad_bf16) & 15))
    return set_error(A3_ERR_ARG, "a3_adamw: buffers must be 32-byte (fp32) / 16-byte (bf16) aligned");
  const double bc1 = 1.0 - pow((double)beta1, step);
  const double bc2 = 1.0 - pow((double)beta2, step);
  long long blocks = ((n >> 3) + 255) / 256;
  const long long cap = (long long)num_sms() * g_adamw_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(master, BFM(weight_bf16), BF(grad_bf16), m, v, n, lr,
                                                    beta1, beta2, eps, weight_decay, (float)bc1,
                                                    (float)sqrt(bc2), norm_sq, max_norm);
  return check_launch("a3_adamw");
}

extern "C" int a3_kv_append(const void* k, const void* v, long long ld, void* kcache, void* vcache,
                            int B, int T, int H, int head_dim, int start, int max_seq, void* stream) {
  if (((H * head_dim) & 7) || (ld & 7)) return set_error(A3_ERR_ARG, "a3_kv_append: alignment");
  if (start < 0 || start + T > max_seq) return set_error(A3_ERR_ARG, "a3_kv_append: cache overflow");
  if (B * T <= 0) return A3_OK;
  kv_append_kernel<<<B * T, 128, 0, ST(stream)>>>(BF(k), BF(v), ld, BFM(kcache), BFM(vcache), T,
                                                  H * head_dim, start, max_seq);
  return check_launch("a3_kv_append");
}

// Row-wise HBM-bound kernels of the hot path: RMSNorm / LayerNorm (fwd + bwd), gated-SiLU glue,
// RoPE on a strided buffer, column sums.  One CTA per row (or a grid-stride over rows for the
// backward passes that also reduce over rows), 16-byte vector loads, warp-shuffle + shared-memory
// reductions, fp32 statistics.  Rounding points follow SURVEY A.9.
#include "common.cuh"
#include "ptx.cuh"

namespace a3 {

constexpr int ROW_THREADS = 256;
constexpr int MAX_CHUNKS = 4;  // per-thread 8-element chunks cached in registers: D <= 8192

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for ROW_THREADS threads; every thread receives the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = unpack_bf16x2(w[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// ------------------------------------------------------------------------------ RMSNorm forward
__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out, int D,
                   long long ldx, long long ldy, float eps, int group_rows, long long group_stride) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL successor may start prefetching weights
  __shared__ float red[32];
  const int row = blockIdx.x;
  // grouped input rows (the text rows of every sample of a [B][L][D] stream): row r of the output is row
  // r % group_rows of group r / group_rows
  const __nv_bfloat16* xr = group_rows > 0
                                ? x + (row / group_rows) * group_stride + (row % group_rows) * ldx
                                : x + row * ldx;
  const int nchunks = D >> 3;
  float xv[MAX_CHUNKS][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
      load8(xr + c * 8, xv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += xv[i][j] * xv[i][j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)D + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
      float wv[8], o[8];
      load8(w + c * 8, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = bf16_round(xv[i][j] * rstd) * wv[j];
      store8(yr + c * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------ RMSNorm backward
// y = n * w, n = bf16(x * rstd).  dn = dy * w; dx = rstd * (dn - xhat * mean(dn * xhat));
// dw += dy * n.  Each CTA walks rows blockIdx.x, +gridDim.x, ... and keeps its dw slice in
// registers; the [nparts][D] partials are summed by colsum_partials.
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = unpack_bf16x2(w[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}

// Software-pipelined over rows: while row r is reduced (two block barriers) the 16-byte loads of row
// r + gridDim.x (x, dy) and of row r's previous dx are already in flight as packed bf16, so each CTA
// keeps ~1.5 rows of traffic outstanding instead of idling through the reduction (3.4 -> ~5 TB/s).
template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS, NCH <= 2 ? 2 : 1)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd_in,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_partial, int rows, int D,
                   int add_to_dx, int group_rows, long long group_stride) {
  __shared__ float red[32];
  const int nchunks = D >> 3;
  float dw[NCH][8];
  uint4 xp[NCH], gp[NCH];
  // x / dx rows may be grouped (see rmsnorm_fwd_kernel); dy and rstd are dense [rows][D]
  auto xoff = [&](int r) -> long long {
    return group_rows > 0 ? (long long)(r / group_rows) * group_stride + (long long)(r % group_rows) * D
                          : (long long)r * D;
  };
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[i][j] = 0.f;
    if (c < nchunks && (int)blockIdx.x < rows) {
      xp[i] = *reinterpret_cast<const uint4*>(x + xoff(blockIdx.x) + c * 8);
      gp[i] = *reinterpret_cast<const uint4*>(dy + (long long)blockIdx.x * D + c * 8);
    }
  }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int nrow = row + gridDim.x;
    __nv_bfloat16* dr = dx + xoff(row);
    const long long nxo = nrow < rows ? xoff(nrow) : 0;
    uint4 xn[NCH], gn[NCH], pv[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = threadIdx.x + i * ROW_THREADS;
      if (c < nchunks) {
        if (nrow < rows) {
          xn[i] = *reinterpret_cast<const uint4*>(x + nxo + c * 8);
          gn[i] = *reinterpret_cast<const uint4*>(dy + (long long)nrow * D + c * 8);
        }
        if (add_to_dx) pv[i] = *reinterpret_cast<const uint4*>(dr + c * 8);
      }
    }
    const float rstd = rstd_in[row];
    float xh[NCH][8], dn[NCH][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = threadIdx.x + i * ROW_THREADS;
      if (c < nchunks) {
        float g[8], wv[8];
        unpack8(xp[i], xh[i]);
        unpack8(gp[i], g);
        load8(w + c * 8, wv);  // 8 KB, L1-resident after the first row
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] *= rstd;
          dw[i][j] += g[j] * bf16_round(xh[i][j]);
          dn[i][j] = g[j] * wv[j];
          dot += dn[i][j] * xh[i][j];
        }
      }
    }
    dot = block_sum(dot, red) / (float)D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = threadIdx.x + i * ROW_THREADS;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dn[i][j] - xh[i][j] * dot);
        if (add_to_dx) {
          float prev[8];
          unpack8(pv[i], prev);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = bf16_round(o[j]) + prev[j];
        }
        store8(dr + c * 8, o);
        xp[i] = xn[i];
        gp[i] = gn[i];
      }
    }
  }
  float* out = dw_partial + (long long)blockIdx.x * D;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
      *reinterpret_cast<float4*>(out + c * 8) = make_float4(dw[i][0], dw[i][1], dw[i][2], dw[i][3]);
      *reinterpret_cast<float4*>(out + c * 8 + 4) =
          make_float4(dw[i][4], dw[i][5], dw[i][6], dw[i][7]);
    }
  }
}

// ------------------------------------------------------------------------------ LayerNorm
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int D,
                     long long ldx, long long ldy, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  const int nchunks = D >> 3;
  float xv[MAX_CHUNKS][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
      load8(xr + c * 8, xv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += xv[i][j];
    }
  }
  const float mean = block_sum(s, red) / (float)D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = xv[i][j] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)D + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
      float wv[8], bv[8], o[8];
      load8(w + c * 8, wv);
      load8(b + c * 8, bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xv[i][j] - mean) * rstd * wv[j] + bv[j];
      store8(yr + c * 8, o);
    }
  }
}

// dx = rstd * (dn - mean(dn) - xhat * mean(dn * xhat)), dn = dy * w; dw += dy * xhat; db += dy.
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const __nv_bfloat16* __restrict__ w, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                     float* __restrict__ dw_partial, float* __restrict__ db_partial, int rows,
                     int D) {
  __shared__ float red[32];
  const int nchunks = D >> 3;
  float dw[MAX_CHUNKS][8], db[MAX_CHUNKS][8], wv[MAX_CHUNKS][8];
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[i][j] = db[i][j] = 0.f;
    if (c < nchunks) load8(w + c * 8, wv[i]);
  }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const __nv_bfloat16* xr = x + (long long)row * D;
    const __nv_bfloat16* gr = dy + (long long)row * D;
    float xh[MAX_CHUNKS][8], dn[MAX_CHUNKS][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS; ++i) {
      const int c = threadIdx.x + i * ROW_THREADS;
      if (c < nchunks) {
        float g[8];
        load8(xr + c * 8, xh[i]);
        load8(gr + c * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mean) * rstd;
          dw[i][j] += g[j] * xh[i][j];
          db[i][j] += g[j];
          dn[i][j] = g[j] * wv[i][j];
          s1 += dn[i][j];
          s2 += dn[i][j] * xh[i][j];
        }
      }
    }
    s1 = block_sum(s1, red) / (float)D;
    s2 = block_sum(s2, red) / (float)D;
    __nv_bfloat16* dr = dx + (long long)row * D;
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS; ++i) {
      const int c = threadIdx.x + i * ROW_THREADS;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dn[i][j] - s1 - xh[i][j] * s2);
        store8(dr + c * 8, o);
      }
    }
  }
  float* ow = dw_partial + (long long)blockIdx.x * D;
  float* ob = db_partial + (long long)blockIdx.x * D;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ow[c * 8 + j] = dw[i][j];
        ob[c * 8 + j] = db[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------ column sums
// out[c] (+)= sum_p partial[p][c].  CTA = 32 x 8 threads over a 32-column strip: the partial rows are
// strided over threadIdx.y and combined through shared memory in a fixed order (deterministic).  D / 32 CTAs
// (128 at D = 4096) instead of D / 256 column-per-thread CTAs that walked all ~300 rows serially (25 -> ~7 us).
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ partial, int nparts, int D, void* __restrict__ out, int out_fp32,
                       int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < D)
    for (int p = threadIdx.y; p < nparts; p += 8) s += partial[(long long)p * D + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || c >= D) return;
  s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
  if (out_fp32) {
    float* o = reinterpret_cast<float*>(out);
    o[c] = accumulate ? o[c] + s : s;
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (accumulate) s += __bfloat162float(o[c]);
    o[c] = __float2bfloat16_rn(s);
  }
}

// out[n] = sum_m X[m][n].  CTA = 32 x 8 threads over a 32-column strip; rows strided by 8 and
// then reduced through shared memory.  Deterministic.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, int M, int N,
                                   long long ldx, __nv_bfloat16* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (col < N)
    for (int m = threadIdx.y; m < M; m += 8) s += __bfloat162float(X[(long long)m * ldx + col]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (accumulate) t += __bfloat162float(out[col]);
    out[col] = __float2bfloat16_rn(t);
  }
}

// ------------------------------------------------------------------------------ SwiGLU glue
// gu = [rows][2F] (g | u);  y = bf16(bf16(silu(g)) * u)   (llama_ens5.py:213-214: F.silu rounds,
// the product rounds).
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ y,
                                  long long total_chunks, int Fc) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL successor may start prefetching weights
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Fc;
    const int c = (int)(i - row * Fc);
    const __nv_bfloat16* g = gu + row * (2LL * Fc * 8) + c * 8;
    float gv[8], uv[8], o[8];
    load8(g, gv);
    load8(g + (long long)Fc * 8, uv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(gv[j] / (1.f + __expf(-gv[j]))) * uv[j];
    store8(y + row * ((long long)Fc * 8) + c * 8, o);
  }
}
// dg = dy * u * sig * (1 + g * (1 - sig));  du = dy * bf16(silu(g))
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu,
                                  const __nv_bfloat16* __restrict__ dy,
                                  __nv_bfloat16* __restrict__ dgu, long long total_chunks, int Fc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Fc;
    const int c = (int)(i - row * Fc);
    const long long off = row * (2LL * Fc * 8) + c * 8;
    float gv[8], uv[8], d[8], dg[8], du[8];
    load8(gu + off, gv);
    load8(gu + off + (long long)Fc * 8, uv);
    load8(dy + row * ((long long)Fc * 8) + c * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sig = 1.f / (1.f + __expf(-gv[j]));
      const float silu = bf16_round(gv[j] * sig);
      // autograd: d(prod)/d(silu) = dy*u rounded to bf16, then silu' in fp32
      dg[j] = bf16_round(d[j] * uv[j]) * sig * (1.f + gv[j] * (1.f - sig));
      du[j] = d[j] * silu;
    }
    store8(dgu + off, dg);
    store8(dgu + off + (long long)Fc * 8, du);
  }
}

// ------------------------------------------------------------------------------ RoPE (strided)
// One thread per (row, head, 4 pairs).  table float2 [pos][head_dim/2].
__global__ void rope_apply_kernel(__nv_bfloat16* __restrict__ x, long long ld, long long total,
                                  int heads, int head_dim, const float2* __restrict__ table,
                                  int period, int pos0, int inverse) {
  const int cpr = heads * head_dim / 8;  // 8-element chunks per row
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cpr;
    const int c = (int)(i - row * cpr);
    const int col = c * 8;
    const int pos = pos0 + (int)(row % period);
    const float2* tp = table + (long long)pos * (head_dim / 2) + ((col % head_dim) >> 1);
    __nv_bfloat16* p = x + row * ld + col;
    float v[8], o[8];
    load8(p, v);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float2 cs = tp[j >> 1];
      if (inverse) cs.y = -cs.y;
      o[j] = v[j] * cs.x - v[j + 1] * cs.y;
      o[j + 1] = v[j] * cs.y + v[j + 1] * cs.x;
    }
    store8(p, o);
  }
}

static inline int grid_for(long long work, int threads, int cap_mult = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)num_sms() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace a3

using namespace a3;
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int check_rowshape(const char* what, int rows, int D) {
  if (rows < 0 || D <= 0) return set_error(A3_ERR_ARG, what);
  if ((D & 7) || D > 8 * ROW_THREADS * MAX_CHUNKS) return set_error(A3_ERR_ARG, what);
  return A3_OK;
}

extern "C" int a3_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int D,
                              long long ldx, long long ldy, float eps, int group_rows,
                              long long group_stride, void* stream) {
  if (check_rowshape("a3_rmsnorm_fwd: D must be a multiple of 8 and <= 8192", rows, D)) return A3_ERR_ARG;
  if (rows == 0) return A3_OK;
  if ((ldx & 7) || (ldy & 7) || (group_stride & 7) || group_rows < 0)
    return set_error(A3_ERR_ARG, "a3_rmsnorm_fwd: ld % 8 != 0");
  rmsnorm_fwd_kernel<<<rows, ROW_THREADS, 0, ST(stream)>>>(BF(x), BF(w), BFM(y), rstd, D, ldx, ldy, eps,
                                                           group_rows, group_stride);
  return check_launch("a3_rmsnorm_fwd");
}

extern "C" int a3_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd,
                              void* dx, float* dw_partial, int nparts, int rows, int D,
                              int add_to_dx, int group_rows, long long group_stride, void* stream) {
  if (check_rowshape("a3_rmsnorm_bwd: D must be a multiple of 8 and <= 8192", rows, D)) return A3_ERR_ARG;
  if (nparts <= 0) return set_error(A3_ERR_ARG, "a3_rmsnorm_bwd: nparts <= 0");
  if ((group_stride & 7) || group_rows < 0) return set_error(A3_ERR_ARG, "a3_rmsnorm_bwd: group_stride % 8 != 0");
  const int nch = (D / 8 + ROW_THREADS - 1) / ROW_THREADS;
  if (nch <= 1)
    rmsnorm_bwd_kernel<1><<<nparts, ROW_THREADS, 0, ST(stream)>>>(BF(dy), BF(x), BF(w), rstd, BFM(dx),
                                                                  dw_partial, rows, D, add_to_dx, group_rows, group_stride);
  else if (nch == 2)
    rmsnorm_bwd_kernel<2><<<nparts, ROW_THREADS, 0, ST(stream)>>>(BF(dy), BF(x), BF(w), rstd, BFM(dx),
                                                                  dw_partial, rows, D, add_to_dx, group_rows, group_stride);
  else
    rmsnorm_bwd_kernel<4><<<nparts, ROW_THREADS, 0, ST(stream)>>>(BF(dy), BF(x), BF(w), rstd, BFM(dx),
                                                                  dw_partial, rows, D, add_to_dx, group_rows, group_stride);
  return check_launch("a3_rmsnorm_bwd");
}

extern "C" int a3_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean,
                                float* rstd, int rows, int D, long long ldx, long long ldy,
                                float eps, void* stream) {
  if (check_rowshape("a3_layernorm_fwd: D must be a multiple of 8 and <= 8192", rows, D)) return A3_ERR_ARG;
  if (rows == 0) return A3_OK;
  if ((ldx & 7) || (ldy & 7)) return set_error(A3_ERR_ARG, "a3_layernorm_fwd: ld % 8 != 0");
  layernorm_fwd_kernel<<<rows, ROW_THREADS, 0, ST(stream)>>>(BF(x), BF(w), BF(b), BFM(y), mean, rstd,
                                                             D, ldx, ldy, eps);
  return check_launch("a3_layernorm_fwd");
}

extern "C" int a3_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean,
                                const float* rstd, void* dx, float* dw_partial, float* db_partial,
                                int nparts, int rows, int D, void* stream) {
  if (check_rowshape("a3_layernorm_bwd: D must be a multiple of 8 and <= 8192", rows, D)) return A3_ERR_ARG;
  if (nparts <= 0) return set_error(A3_ERR_ARG, "a3_layernorm_bwd: nparts <= 0");
  layernorm_bwd_kernel<<<nparts, ROW_THREADS, 0, ST(stream)>>>(BF(dy), BF(x), BF(w), mean, rstd,
                                                               BFM(dx), dw_partial, db_partial, rows, D);
  return check_launch("a3_layernorm_bwd");
}

extern "C" int a3_colsum_partials(const float* partial, int nparts, int D, void* out, int out_fp32,
                                  int accumulate, void* stream) {
  if (D <= 0 || nparts <= 0) return A3_OK;
  colsum_partials_kernel<<<(D + 31) / 32, dim3(32, 8), 0, ST(stream)>>>(partial, nparts, D, out, out_fp32,
                                                                        accumulate);
  return check_launch("a3_colsum_partials");
}

extern "C" int a3_colsum_bf16(const void* X, int M, int N, long long ldx, void* out, int accumulate,
                              void* stream) {
  if (N <= 0) return A3_OK;
  colsum_bf16_kernel<<<(N + 31) / 32, dim3(32, 8), 0, ST(stream)>>>(BF(X), M, N, ldx, BFM(out),
                                                                    accumulate);
  return check_launch("a3_colsum_bf16");
}

extern "C" int a3_swiglu_fwd(const void* gu, void* y, int rows, int F, void* stream) {
  if (F & 7) return set_error(A3_ERR_ARG, "a3_swiglu_fwd: F % 8 != 0");
  const long long total = (long long)rows * (F / 8);
  if (total <= 0) return A3_OK;
  swiglu_fwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(BF(gu), BFM(y), total, F / 8);
  return check_launch("a3_swiglu_fwd");
}

extern "C" int a3_swiglu_bwd(const void* gu, const void* dy, void* dgu, int rows, int F, void* stream) {
  if (F & 7) return set_error(A3_ERR_ARG, "a3_swiglu_bwd: F % 8 != 0");
  const long long total = (long long)rows * (F / 8);
  if (total <= 0) return A3_OK;
  swiglu_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(BF(gu), BF(dy), BFM(dgu), total, F / 8);
  return check_launch("a3_swiglu_bwd");
}

extern "C" int a3_rope_apply(void* x, long long ld, int rows, int heads, int head_dim,
                             const float* table, int period, int pos0, int inverse, void* stream) {
  if ((head_dim & 7) || (ld & 7) || period <= 0)
    return set_error(A3_ERR_ARG, "a3_rope_apply: head_dim % 8, ld % 8, period > 0 required");
  const long long total = (long long)rows * heads * head_dim / 8;
  if (total <= 0) return A3_OK;
  rope_apply_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(
      BFM(x), ld, total, heads, head_dim, reinterpret_cast<const float2*>(table), period, pos0, inverse);
  return check_launch("a3_rope_apply");
}

// Device-side token selection and stop bookkeeping of the generation loop (reference
// model/accessory/model/meta.py:456-477,568-583), so that a decode step needs no host round trip:
//   sample_top_p_kernel   softmax(logits / T) + nucleus (top-p) sampling of one token per row
//   generate_step_kernel  prompt force-feeding, token write, stop-sequence matching, stop_pos / stopped
// Both are latency-bound row kernels (one CTA per row / one CTA per batch); the only HBM traffic is
// one read of the fp32 logits row (V * 4 bytes).
#include "common.cuh"

namespace a3 {

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_WARPS = SAMPLE_THREADS / 32;
constexpr int SAMPLE_ITEMS = 32;  // probabilities per thread kept in registers when V <= 32768

// Sum over the block in a FIXED order (xor-shuffle tree over the lanes, then the same tree over the 32
// warp partials, done redundantly by every warp so one barrier suffices): every thread gets the same
// bit pattern, and the result is monotone in its non-negative inputs, which the bisection below relies
// on.  `red` must hold SAMPLE_WARPS floats and alternate between two buffers from call to call.
__device__ __forceinline__ float block_sum_fixed(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  static_assert(SAMPLE_WARPS == 32, "one warp partial per lane");
  float t = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
__device__ __forceinline__ float block_max_fixed(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < SAMPLE_WARPS; ++w) t = fmaxf(t, red[w]);
  return t;
}

struct Crossing {
  uint32_t key;  // float bits of the probability value at which the sorted cumulative sum crosses theta
  float above;   // sum of the probabilities strictly greater than that value
  float at;      // tree sum of the probabilities >= that value
  int found;     // 0: the total never exceeds theta
};

// F(k) = sum of p_i with bits(p_i) >= k  (non-increasing in k; p_i >= 0 so the bit order is the value
// order).  Finds the largest k with F(k) > theta by bisection over the 31 value bits: in descending
// sorted order this is the value of the first element whose inclusive cumulative sum exceeds theta --
// the quantity both the nucleus cut (meta.py:576-578) and the inverse-CDF draw need -- without sorting.
template <bool REG>
__device__ Crossing find_crossing(const float* p, const float (&r)[SAMPLE_ITEMS], int V, float theta, float total,
                                  float* red2) {
  Crossing c;
  c.found = total > theta;
  c.key = 0;
  c.above = 0.f;
  c.at = total;
  if (!c.found) return c;
  uint32_t P = 0;
  float atP = total, aboveP = 0.f;
  int par = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t t = P | (1u << bit);
    float acc = 0.f;
    if (REG) {  // the row lives in registers (entries past V are 0 and never >= t, as t >= 1)
#pragma unroll
      for (int j = 0; j < SAMPLE_ITEMS; ++j) acc += (__float_as_uint(r[j]) >= t) ? r[j] : 0.f;
    } else {
      for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) {
        const float v = p[i];
        acc += (__float_as_uint(v) >= t) ? v : 0.f;
      }
    }
    const float f = block_sum_fixed(acc, red2 + par * SAMPLE_WARPS);
    par ^= 1;
    if (f > theta) { P = t; atP = f; } else { aboveP = f; }
    // invariant: F(P) = atP > theta, and F(P + 2^bit) = aboveP <= theta
  }
  c.key = P;
  c.at = atP;
  c.above = aboveP;  // after the last bit: F(P + 1) = sum of values strictly greater than P
  return c;
}

// Smallest m >= 1 with above + m * v > theta: how many entries of the crossing value v (ordered by
// token id) the cumulative sum has to take in before it exceeds theta.
__device__ __forceinline__ int ties_needed(float above, float v, float theta, int V) {
  if (!(v > 0.f)) return 1;
  const float q = floorf((theta - above) / v);
  int m = q < (float)V ? (int)q + 1 : V;
  if (m < 1) m = 1;
  while (m > 1 && above + (float)(m - 1) * v > theta) --m;
  while (m < V && !(above + (float)m * v > theta)) ++m;
  return m;
}

// m-th (1-based) smallest index i with bits(p_i) == key.  Every thread returns the answer.
// Warp w owns the contiguous segment [w * seg, (w + 1) * seg); it is read 32 consecutive floats at a time
// (bank-conflict free) and matches are counted with ballots, which keeps index order: row-major over the
// 32-wide rows, lane-minor inside a row.
__device__ int nth_index_of(const float* p, int V, uint32_t key, int m, int* scan, int* result) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int seg = ((V + SAMPLE_WARPS - 1) / SAMPLE_WARPS + 31) & ~31;  // per-warp segment, multiple of 32
  const int lo = w * seg, hi = min(lo + seg, V);
  int cnt = 0;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    cnt += __popc(__ballot_sync(0xffffffffu, i < hi && __float_as_uint(p[i]) == key));
  }
  __syncthreads();
  if (lane == 0) scan[w] = cnt;
  if (threadIdx.x == 0) *result = -1;
  __syncthreads();
  int excl = 0;
  for (int k = 0; k < w; ++k) excl += scan[k];
  if (excl < m && m <= excl + cnt) {  // warp-uniform: this warp holds the m-th match
    int seen = excl;
    for (int base = lo; base < hi; base += 32) {
      const int i = base + lane;
      const unsigned b = __ballot_sync(0xffffffffu, i < hi && __float_as_uint(p[i]) == key);
      const int c = __popc(b);
      if (seen + c >= m) {
        const int need = m - seen;  // the need-th set bit of b
        const bool mine = ((b >> lane) & 1u) && __popc(b & ((2u << lane) - 1u)) == need;
        if (mine) *result = i;
        break;
      }
      seen += c;
    }
  }
  __syncthreads();
  return *result;
}

// One CTA per row.  probs = softmax(logits / temperature) in fp32 (meta.py:457); sort descending,
// cumulative sum, drop the entries whose EXCLUSIVE cumulative sum exceeds top_p, renormalise
// (meta.py:574-579); draw from the kept entries by inverse CDF with the caller's uniform u[row] in
// [0,1) (torch.multinomial draws from the same distribution; its random stream is not reproducible
// across devices, so the uniform is an input).  Equal probabilities are ordered by token id.
template <bool REG>
__global__ void __launch_bounds__(SAMPLE_THREADS)
sample_top_p_kernel(const float* __restrict__ logits, long long ld, int V, float temperature, float top_p,
                    const float* __restrict__ u, long long* __restrict__ out) {
  extern __shared__ float p[];  // V floats
  __shared__ __align__(16) float red[4 * SAMPLE_WARPS];
  __shared__ int scan[SAMPLE_WARPS];
  __shared__ int result;
  const float* lr = logits + (long long)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) {
    const float x = __fdiv_rn(lr[i], temperature);
    p[i] = x;
    mx = fmaxf(mx, x);
  }
  mx = block_max_fixed(mx, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) {
    const float e = expf(p[i] - mx);
    p[i] = e;
    s += e;
  }
  s = block_sum_fixed(s, red + SAMPLE_WARPS);
  float tot = 0.f;
  float r[SAMPLE_ITEMS];
#pragma unroll
  for (int j = 0; j < SAMPLE_ITEMS; ++j) r[j] = 0.f;
  if (REG) {
#pragma unroll
    for (int j = 0; j < SAMPLE_ITEMS; ++j) {
      const int i = threadIdx.x + j * SAMPLE_THREADS;
      if (i < V) {
        r[j] = __fdiv_rn(p[i], s);
        p[i] = r[j];
        tot += r[j];
      }
    }
  } else {
    for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) {
      const float q = __fdiv_rn(p[i], s);
      p[i] = q;
      tot += q;
    }
  }
  __syncthreads();  // p[] complete before cross-thread reads
  tot = block_sum_fixed(tot, red + 2 * SAMPLE_WARPS);
  __syncthreads();  // red[0..] is reused by the bisection

  // 1. nucleus: first sorted position whose inclusive cumulative sum exceeds top_p is the last one kept
  const Crossing n = find_crossing<REG>(p, r, V, top_p, tot, red);
  float Z = tot;           // normaliser: sum of the kept probabilities
  uint32_t nkey = 0;       // smallest kept value
  int nm = 0x7fffffff;     // how many entries of that value are kept (ties ordered by token id)
  if (n.found) {
    const float v = __uint_as_float(n.key);
    const int m = ties_needed(n.above, v, top_p, V);
    const float zm = n.above + (float)m * v;
    Z = zm < n.at ? zm : n.at;  // a single entry of that value: Z is the tree sum the draw compares against
    nkey = n.key;
    nm = m;
  }
  __syncthreads();
  // 2. inverse-CDF draw: first sorted position whose inclusive cumulative sum exceeds u * Z
  float theta = u[blockIdx.x] * Z;
  if (!(theta < Z)) theta = __uint_as_float(__float_as_uint(Z) - 1);  // u -> 1: stay inside the nucleus
  if (theta < 0.f) theta = 0.f;
  Crossing d = find_crossing<REG>(p, r, V, theta, tot, red);
  uint32_t key = d.found ? d.key : nkey;
  int m = 1;
  if (d.found) {
    const float v = __uint_as_float(d.key);
    m = ties_needed(d.above, v, theta, V);
  }
  if (key < nkey) { key = nkey; m = nm; }          // rounding pushed the draw past the cut: clamp to it
  if (key == nkey && m > nm) m = nm;
  int idx = nth_index_of(p, V, key, m, scan, &result);
  if (idx < 0) {  // fewer than m ties exist (only through float rounding of m): take the last of them
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < V; i += SAMPLE_THREADS) cnt += __float_as_uint(p[i]) == key;
    float cf = block_sum_fixed((float)cnt, red + 2 * SAMPLE_WARPS);
    __syncthreads();
    idx = nth_index_of(p, V, key, max((int)cf, 1), scan, &result);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = idx < 0 ? 0 : idx;
}

// One CTA for the whole batch; the body of the reference loop after the token choice
// (meta.py:463-477), row by row:
//   next = prompt position ? tokens[b][cur] : next_token[b]        (force-feeding, :463-465)
//   tokens[b][cur] = next
//   stop_pos[b] = stopped[b] ? stop_pos[b] : cur + 1               (:467)
//   for each stop sequence s of length k with cur + 1 - k >= 0, in order:
//     hit = tokens[b][cur+1-k .. cur] == s  and  not prompt position  and  not stopped[b]
//     hit: stop_pos[b] = cur + 1 - k; stopped[b] = true             (:468-475)
//   all_stopped[0] = all(stopped)                                   (:476, read by the host later)
__global__ void __launch_bounds__(256)
generate_step_kernel(long long* __restrict__ tokens, long long ld, const int* __restrict__ prompt_len,
                     const long long* __restrict__ next_token, int cur, const long long* __restrict__ stop_tokens,
                     const int* __restrict__ stop_offsets, int n_stop, unsigned char* __restrict__ stopped,
                     long long* __restrict__ stop_pos, int* __restrict__ all_stopped,
                     long long* __restrict__ fed_token, int B) {
  int all = 1;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    long long* row = tokens + (long long)b * ld;
    const bool prompt = cur < prompt_len[b];
    const long long nt = prompt ? row[cur] : next_token[b];
    row[cur] = nt;
    if (fed_token) fed_token[b] = nt;
    bool st = stopped[b] != 0;
    long long sp = st ? stop_pos[b] : (long long)cur + 1;
    for (int s = 0; s < n_stop; ++s) {
      const int o = stop_offsets[s], k = stop_offsets[s + 1] - o;
      if (k <= 0 || cur + 1 - k < 0) continue;
      bool hit = !prompt && !st;
      for (int j = 0; hit && j < k; ++j) hit = row[cur + 1 - k + j] == stop_tokens[o + j];
      if (hit) { sp = cur + 1 - k; st = true; }
    }
    stop_pos[b] = sp;
    stopped[b] = st ? 1 : 0;
    all &= st ? 1 : 0;
  }
  all = __syncthreads_and(all);
  if (threadIdx.x == 0 && all_stopped) all_stopped[0] = all;
}

}  // namespace a3

using namespace a3;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int a3_sample_top_p(const float* logits, long long ld, int B, int V, float temperature,
                               float top_p, const float* u, long long* out, void* stream) {
  if (B <= 0) return A3_OK;
  if (V <= 0 || !(temperature > 0.f)) return set_error(A3_ERR_ARG, "a3_sample_top_p: V > 0 and temperature > 0");
  const size_t smem = (size_t)V * sizeof(float);
  if (smem > 200 * 1024) return set_error(A3_ERR_UNSUPPORTED, "a3_sample_top_p: V > 51200 (row must fit shared memory)");
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&sample_top_p_kernel<true>), 200 * 1024)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&sample_top_p_kernel<false>), 200 * 1024)) return rc;
  if (V <= SAMPLE_THREADS * SAMPLE_ITEMS)
    sample_top_p_kernel<true><<<B, SAMPLE_THREADS, smem, ST(stream)>>>(logits, ld, V, temperature, top_p, u, out);
  else
    sample_top_p_kernel<false><<<B, SAMPLE_THREADS, smem, ST(stream)>>>(logits, ld, V, temperature, top_p, u, out);
  return check_launch("a3_sample_top_p");
}

extern "C" int a3_generate_step(long long* tokens, long long ld_tokens, const int* prompt_len,
                                const long long* next_token, int cur_pos, const long long* stop_tokens,
                                const int* stop_offsets, int n_stop, unsigned char* stopped,
                                long long* stop_pos, int* all_stopped, long long* fed_token, int B,
                                void* stream) {
  if (B <= 0) return A3_OK;
  if (cur_pos < 0 || cur_pos >= ld_tokens) return set_error(A3_ERR_ARG, "a3_generate_step: cur_pos outside the token buffer");
  if (n_stop < 0 || (n_stop > 0 && (!stop_tokens || !stop_offsets)))
    return set_error(A3_ERR_ARG, "a3_generate_step: stop sequences missing");
  generate_step_kernel<<<1, 256, 0, ST(stream)>>>(tokens, ld_tokens, prompt_len, next_token, cur_pos, stop_tokens,
                                                  stop_offsets, n_stop, stopped, stop_pos, all_stopped, fed_token, B);
  return check_launch("a3_generate_step");
}

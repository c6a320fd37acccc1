// FlashAttention-style forward on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// One CTA = 128 query rows of one (batch, head); 4 warps, thread t owns query row t (TMEM lane t),
// so row max / row sum need no shuffles.  Per 64-key tile:
//     S = Q K^T        tcgen05.mma  (A = Q smem K-major, B = K tile smem K-major)  -> TMEM[0,64)
//     P = exp2(S*c - m)  threads read S from TMEM, write bf16 P to smem in the UMMA SW128 layout,
//                        rescale O in TMEM when the row max moved
//     O += P V         tcgen05.mma  (A = P smem K-major, B = V tile smem MN-major) -> TMEM[64,64+HD)
// K/V tiles are double-buffered by TMA; the S-MMA of tile j+1 is issued right behind the PV-MMA of
// tile j so one tcgen05.commit covers both.  2-3 CTAs are co-resident per SM (112 KB / 64 KB of
// smem) and fill the tensor pipe while their neighbours run softmax.
//
// Replaces flash_attn_func / F.scaled_dot_product_attention in reference
// model/accessory/model/LLM/llama_ens5.py:142-167 (causal, head_dim 128) and the attention core of
// open_clip's ResidualAttentionBlock (llama_ens5.py:366; no mask, head_dim 64).
#include "common.cuh"
#include "ptx.cuh"

namespace a3 {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 64;

template <int HD>
struct AttnLegacySmem {
  static constexpr int Q_BYTES = ATT_BM * HD * 2;
  static constexpr int KV_BYTES = ATT_BN * HD * 2;
  static constexpr int P_BYTES = ATT_BM * ATT_BN * 2;
  static constexpr int OFF_K = Q_BYTES;                 // [2]
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;    // [2]
  static constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 128;  // no alignment slack: 2 CTAs/SM must fit at HD=128
  static constexpr int TMEM_COLS = (ATT_BN + HD) <= 128 ? 128 : 256;
};

struct AttnLegacyParams {
  __nv_bfloat16* o;
  float* lse;
  long long ldo;
  int B, H, Tq, Tk;
  float scale;
  int causal;
};

template <int HD>
__global__ void __launch_bounds__(128)
attn_fwd_legacy_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnLegacyParams p) {
  using S = AttnLegacySmem<HD>;
  constexpr int NCH = HD / 64;  // 64-column (128-byte) chunks per head row
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // SW128 tiles need a 1024-byte aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;         // [1]
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* mma_bar = bars + 3;    // [1] "S_j ready (and every earlier MMA retired)"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const int off = p.Tk - p.Tq;  // right-aligned causal offset
  int kv_end = p.Tk;
  if (p.causal) kv_end = min(p.Tk, q0 + ATT_BM + off);
  const int n_tiles = (max(kv_end, 0) + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t t_S = t_lane;            // columns [0, 64)
  const uint32_t t_O = t_lane + ATT_BN;   // columns [64, 64 + HD)

  const uint32_t sQ = smem_u32(smem), sK = sQ + S::OFF_K, sV = sQ + S::OFF_V, sP = sQ + S::OFF_P;
  constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, HD, 0, 1);

  auto load_kv = [&](int j) {
    const int st = j & 1;
    mbar_arrive_expect_tx(&kv_full[st], 2 * S::KV_BYTES);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem + S::OFF_K + st * S::KV_BYTES + c * (ATT_BN * 128), &tmK, &kv_full[st],
                  h * HD + c * 64, j * ATT_BN, b);
      tma_load_3d(smem + S::OFF_V + st * S::KV_BYTES + c * (ATT_BN * 128), &tmV, &kv_full[st],
                  h * HD + c * 64, j * ATT_BN, b);
    }
  };
  auto issue_s = [&](int j) {  // S = Q K_j^T
    const int st = j & 1;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      const uint32_t coff = (kk >> 2) , inner = (kk & 3) * 32;
      const uint64_t ad = make_smem_desc_sw128(sQ + coff * (ATT_BM * 128) + inner, 16, 1024);
      const uint64_t bd =
          make_smem_desc_sw128(sK + st * S::KV_BYTES + coff * (ATT_BN * 128) + inner, 16, 1024);
      umma_bf16(tmem_base, ad, bd, idesc_s, kk > 0 ? 1u : 0u);
    }
  };
  auto issue_pv = [&](int j) {  // O += P V_j
    const int st = j & 1;
#pragma unroll
    for (int kk = 0; kk < ATT_BN / 16; ++kk) {
      const uint64_t ad = make_smem_desc_sw128(sP + kk * 32, 16, 1024);
      const uint64_t bd = make_smem_desc_sw128(sV + st * S::KV_BYTES + kk * (16 * 128),
                                               ATT_BN * 128, 1024);
      umma_bf16(tmem_base + ATT_BN, ad, bd, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
    }
  };

  if (n_tiles > 0 && threadIdx.x == 0) {
    mbar_arrive_expect_tx(q_full, S::Q_BYTES);
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      tma_load_3d(smem + c * (ATT_BM * 128), &tmQ, q_full, h * HD + c * 64, q0, b);
    load_kv(0);
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    umma_commit(mma_bar);
  }
  __syncwarp();

  const int row = q0 + warp * 32 + lane;  // query index of this thread
  const float c2 = p.scale * 1.4426950408889634f;
  float m_run = -INFINITY, l_run = 0.f;
  uint8_t* p_row = smem + S::OFF_P + (warp * 32 + lane) * 128;
  const int sw = (lane & 7);

  for (int j = 0; j < n_tiles; ++j) {
    mbar_wait(mma_bar, j & 1);  // S_j done, and with it every earlier MMA (PV_{j-1})
    tc_fence_after();
    if (threadIdx.x == 0 && j + 1 < n_tiles) load_kv(j + 1);  // stage (j+1)&1 was freed by PV_{j-1}
    __syncwarp();

    uint32_t s0[32], s1[32];
    tmem_ld_32x32(t_S, s0);
    tmem_ld_32x32(t_S + 32, s1);
    tmem_ld_wait();
    const int k0 = j * ATT_BN;
    const int limit = p.causal ? min(p.Tk - 1, row + off) : p.Tk - 1;  // last visible key
    float mx = m_run;
    if (k0 + ATT_BN - 1 > limit) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (k0 + i > limit) s0[i] = 0xff800000u;
        if (k0 + 32 + i > limit) s1[i] = 0xff800000u;
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
    const float m_safe = mx == -INFINITY ? 0.f : mx;
    const float alpha = exp2f((m_run - m_safe) * c2);  // m_run = -inf -> 0
    const float mc = m_safe * c2;
    float lsum = 0.f;
    // P (bf16) -> smem, K-major SW128: 16-byte unit u of row r lives at r*128 + ((u ^ (r&7)) * 16)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = u * 8 + t * 2;
        const float a = exp2f(__uint_as_float(i < 32 ? s0[i] : s1[i - 32]) * c2 - mc);
        const float bb = exp2f(__uint_as_float(i + 1 < 32 ? s0[i + 1] : s1[i + 1 - 32]) * c2 - mc);
        lsum += a + bb;
        w[t] = pack_bf16x2(a, bb);
      }
      *reinterpret_cast<uint4*>(p_row + ((u ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    l_run = l_run * alpha + lsum;
    m_run = mx;
    // rescale O when any row of this warp moved its max (warp-uniform branch: tcgen05.ld/st are
    // warp collectives)
    if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32(t_O + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32(t_O + c * 32, o);
      }
      tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_pv(j);
      if (j + 1 < n_tiles) {
        mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      umma_commit(mma_bar);
    }
    __syncwarp();
  }

  // ---- epilogue: O / l -> bf16, lse
  if (n_tiles > 0) {
    mbar_wait(mma_bar, n_tiles & 1);
    tc_fence_after();
  }
  const bool row_ok = row < p.Tq;
  const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
  __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.Tq + row) * p.ldo + h * HD;
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) {
    uint32_t o[32];
    if (n_tiles > 0) {
      tmem_ld_32x32(t_O + c * 32, o);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0u;
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q;
        q.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        q.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        q.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        q.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 32 + i) = q;
      }
    }
  }
  if (row_ok && p.lse)
    p.lse[(static_cast<long long>(b) * p.H + h) * p.Tq + row] =
        l_run > 0.f ? m_run * p.scale + logf(l_run) : -INFINITY;

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int HD>
static int launch_attn_fwd_legacy(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                           const AttnLegacyParams& p, cudaStream_t stream) {
  using S = AttnLegacySmem<HD>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&attn_fwd_legacy_kernel<HD>), S::TOTAL)) return rc;
  dim3 grid((p.Tq + ATT_BM - 1) / ATT_BM, p.H, p.B);
  attn_fwd_legacy_kernel<HD><<<grid, 128, S::TOTAL, stream>>>(tmQ, tmK, tmV, p);
  return check_launch("a3_attn_fwd");
}

}  // namespace a3

namespace a3 {
int attn_fwd_legacy(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                           int H, int Tq, int Tk, int head_dim, long long ldq, long long ldk,
                           long long ldv, long long ldo, long long q_bstride, long long kv_bstride,
                           float scale, int causal, void* stream) {
  if (B <= 0 || H <= 0 || Tq <= 0) return A3_OK;
  if (Tk <= 0) return set_error(A3_ERR_ARG, "a3_attn_fwd: Tk must be positive");
  if (head_dim != 64 && head_dim != 128)
    return set_error(A3_ERR_UNSUPPORTED, "a3_attn_fwd: head_dim must be 64 or 128");
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7) || (q_bstride & 7) || (kv_bstride & 7) ||
      (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
    return set_error(A3_ERR_ARG, "a3_attn_fwd: strides % 8 == 0 and 16-byte aligned pointers required");
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t width = (uint64_t)H * head_dim;
  int rc = make_tmap_3d_bf16(&tmQ, q, width, (uint64_t)Tq, (uint64_t)B, (uint64_t)ldq * 2,
                             (uint64_t)q_bstride * 2, 64, ATT_BM, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmK, k, width, (uint64_t)Tk, (uint64_t)B, (uint64_t)ldk * 2,
                         (uint64_t)kv_bstride * 2, 64, ATT_BN, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmV, v, width, (uint64_t)Tk, (uint64_t)B, (uint64_t)ldv * 2,
                         (uint64_t)kv_bstride * 2, 64, ATT_BN, 1);
  if (rc) return rc;
  AttnLegacyParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.ldo = ldo;
  p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk;
  p.scale = scale;
  p.causal = causal ? 1 : 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return head_dim == 64 ? launch_attn_fwd_legacy<64>(tmQ, tmK, tmV, p, s)
                        : launch_attn_fwd_legacy<128>(tmQ, tmK, tmV, p, s);
}
}  // namespace a3

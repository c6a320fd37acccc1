// Which 4-D TMA box / swizzle combinations does sm_100a accept for the patch loader?  (stand-alone probe)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma4d_probe tma4d_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#include <cstring>
__global__ void probe(const __grid_constant__ CUtensorMap tm, uint32_t bytes, uint16_t* out, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sd), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(sb), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(sb) : "memory");
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
__global__ void probe3(const __grid_constant__ CUtensorMap tm, uint32_t bytes, uint16_t* out, int c0, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(sd), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(sb), "r"(c0), "r"(c1), "r"(c2) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(sb) : "memory");
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1; int vi = -1;
  const int S = 224, P = 14, G = 16, B = 2;
  std::vector<__nv_bfloat16> h((size_t)B * 3 * S * S);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 251));
  __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint16_t* out; cudaMalloc(&out, 65536);
  struct V { const char* name; CUtensorMapSwizzle sw; cuuint32_t box[4]; } vs[] = {
      {"SW128 box{16,4,8,1}", CU_TENSOR_MAP_SWIZZLE_128B, {16, 4, 8, 1}}, {"NONE  box{16,4,8,1}", CU_TENSOR_MAP_SWIZZLE_NONE, {16, 4, 8, 1}},
      {"SW32  box{16,4,8,1}", CU_TENSOR_MAP_SWIZZLE_32B, {16, 4, 8, 1}},  {"SW32  box{16,1,8,1}", CU_TENSOR_MAP_SWIZZLE_32B, {16, 1, 8, 1}},
      {"SW128 box{64,1,8,1}", CU_TENSOR_MAP_SWIZZLE_128B, {64, 1, 8, 1}}, {"SW64  box{16,4,8,1}", CU_TENSOR_MAP_SWIZZLE_64B, {16, 4, 8, 1}},
      {"SW128 box{16,4,1,1}", CU_TENSOR_MAP_SWIZZLE_128B, {16, 4, 1, 1}}, {"NONE  box{16,14,8,1}", CU_TENSOR_MAP_SWIZZLE_NONE, {16, 14, 8, 1}}};
  for (auto& v : vs) {
    if (++vi != only && only >= 0) continue;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)S, (cuuint64_t)P, (cuuint64_t)G, (cuuint64_t)3 * B};
    cuuint64_t st[3] = {(cuuint64_t)S * 2, (cuuint64_t)P * S * 2, (cuuint64_t)S * S * 2};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, v.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", v.name, (int)r); continue; }
    uint32_t bytes = v.box[0] * v.box[1] * v.box[2] * v.box[3] * 2;
    cudaMemset(out, 0, 65536);
    probe<<<1, 128, 32768>>>(tm, bytes, out, 14 /*px=1*/, 4, 8, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: RUN FAILED %s\n", v.name, cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> o(bytes / 2); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    printf("%s: ok, first 32-byte rows:", v.name);
    for (int rrow = 0; rrow < 6; ++rrow) { __nv_bfloat16 b; memcpy(&b, &o[rrow * 16], 2); printf(" %g", __bfloat162float(b)); }
    printf("\n");
  }
  if (only == 8 || only == 9) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)S, (cuuint64_t)P, (cuuint64_t)G * 3 * B};
    cuuint64_t st[2] = {(cuuint64_t)S * 2, (cuuint64_t)P * S * 2};
    cuuint32_t box[3] = {16, 4, 8}, es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        only == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("3-D patch map (%s): encode %d\n", only == 8 ? "SW128" : "NONE", (int)r);
    const int c0 = argc > 2 ? atoi(argv[2]) : 14;
    probe3<<<1, 128, 32768>>>(tm, 16 * 4 * 8 * 2, out, c0, 4, 8 + 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("3-D patch map run: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
      std::vector<uint16_t> o(16 * 4 * 8); cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
      // expected (no swizzle): element (x, ky, py) = image1.. value at channel 1 (c2 = 8 + 16 -> py 8 of channel index 1): idx = ((1*S*S) + (8*14+4+ky)*S + 14 + x) % 251
      int bad = 0;
      for (int py = 0; py < 8; ++py) for (int ky = 0; ky < 4; ++ky) for (int x = 0; x < 16; ++x) {
        size_t gi = (size_t)1 * S * S + (size_t)((8 + py) * 14 + 4 + ky) * S + c0 + x;
        float want = (float)(gi % 251);
        int off = (py * 4 + ky) * 16 + x;
        if (only == 8) { int byte = off * 2; int row = byte / 128, ch = (byte % 128) / 16; byte = row * 128 + ((ch ^ (row & 7)) * 16) + byte % 16; off = byte / 2; }
        __nv_bfloat16 b; memcpy(&b, &o[off], 2);
        if (__bfloat162float(b) != __bfloat162float(__float2bfloat16(want))) ++bad;
      }
      printf("layout check: %d mismatches of 512\n", bad);
      // where does element (x 0, ky 1, py 0) and (x 0, ky 0, py 1) land?
      for (int t = 0; t < 2; ++t) {
        size_t gi = (size_t)1 * S * S + (size_t)((8 + (t ? 1 : 0)) * 14 + 4 + (t ? 0 : 1)) * S + c0;
        float want = __bfloat162float(__float2bfloat16((float)(gi % 251)));
        for (int off = 0; off < 512; ++off) { __nv_bfloat16 b; memcpy(&b, &o[off], 2); if (__bfloat162float(b) == want) printf("  %s found at element offset %d (byte %d)\n", t ? "(ky0,py1)" : "(ky1,py0)", off, off * 2); }
      }
    }
  }
  return 0;
}

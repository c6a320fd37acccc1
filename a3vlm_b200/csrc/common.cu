// Host-side plumbing behind the C ABI: error slot, launch accounting, device queries and TMA
// tensor-map encoding through the driver entry point (resolved at run time: no libcuda link).
#include <cstdlib>
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace a3 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return A3_ERR_CUDA;
  }
  return A3_OK;
}

int ensure_dyn_smem(const void* kernel, size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(A3_ERR_CUDA, "cudaGetDevice failed");
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = done[std::make_pair(kernel, dev)];
  if (bytes <= cur) return A3_OK;
  const size_t want = bytes > 48 * 1024 ? bytes : 48 * 1024;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
  if (e != cudaSuccess) return set_error(A3_ERR_CUDA, cudaGetErrorString(e));
  cur = want;
  return A3_OK;
}

// Reports the alignment of the dynamic shared-memory window of a launch (one thread).
__global__ void smem_probe_kernel(unsigned int* out) {
  extern __shared__ __align__(1024) uint8_t probe_smem[];
  *out = static_cast<unsigned int>(__cvta_generic_to_shared(probe_smem)) & 1023u;
}

int check_dyn_smem_alignment(const char* who) {
  static std::atomic<int> state[64];  // 0 unknown, 1 aligned, 2 misaligned
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return set_error(A3_ERR_CUDA, "cudaGetDevice failed");
  int st = state[dev].load(std::memory_order_relaxed);
  if (st == 0) {
    unsigned int* d = nullptr;
    unsigned int h = 1;
    if (cudaMalloc(&d, sizeof(unsigned int)) != cudaSuccess) return set_error(A3_ERR_CUDA, "shared-memory probe: allocation failed");
    smem_probe_kernel<<<1, 1, 1024>>>(d);
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(A3_ERR_CUDA, cudaGetErrorString(e));
    st = h == 0 ? 1 : 2;
    state[dev].store(st, std::memory_order_relaxed);
  }
  if (st == 2) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "%s: dynamic shared memory is not 1024-byte aligned on this device / driver", who);
    return set_error(A3_ERR_UNSUPPORTED, buf);
  }
  return A3_OK;
}

static std::atomic<int> g_sm_limit{0};

int num_sms() {
  const int lim = g_sm_limit.load(std::memory_order_relaxed);
  if (lim > 0) return lim;
  return device_sms();
}

int device_sms() {
  static std::atomic<int> n[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// A tensor map is a pure function of (address, rank, dims, strides, box, dtype, swizzle), and a training step asks for
// the same ~1 500 of them every step (every GEMM operand, every attention Q / K / V / dO / dQ view).  A per-thread
// direct-mapped cache (no lock, no allocation; a colliding key simply re-encodes) answers the repeats.
struct TmapKey {
  const void* ptr;
  int rank, dtype, swz, pad;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], pad2;
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int TMAP_SLOTS = 2048;

static inline uint64_t tmap_hash(const TmapKey& k) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;  // FNV-1a
  for (size_t i = 0; i < sizeof(TmapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

static int encode_uncached(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                           const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapDataType dtype,
                           CUtensorMapSwizzle swz);

static int encode(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box,
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  static const bool cache_off = getenv("A3_TMAP_CACHE") && atoi(getenv("A3_TMAP_CACHE")) == 0;
  if (cache_off || rank < 1 || rank > 5) return encode_uncached(tm, ptr, rank, dims, strides, box, dtype, swz);
  thread_local static TmapSlot* slots = nullptr;
  if (!slots) slots = new TmapSlot[TMAP_SLOTS]();  // zero-initialised: valid = false
  TmapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.rank = rank; k.dtype = (int)dtype; k.swz = (int)swz;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
  TmapSlot& s = slots[tmap_hash(k) & (TMAP_SLOTS - 1)];
  if (s.valid && std::memcmp(&s.key, &k, sizeof(k)) == 0) {
    *tm = s.map;
    return A3_OK;
  }
  if (int rc = encode_uncached(tm, ptr, rank, dims, strides, box, dtype, swz)) return rc;
  s.key = k; s.map = *tm; s.valid = true;
  return A3_OK;
}

static int encode_uncached(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                           const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapDataType dtype,
                           CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(A3_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, dtype, (cuuint32_t)rank, const_cast<void*>(ptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    std::snprintf(buf, sizeof(buf),
                  "cuTensorMapEncodeTiled failed (%d): rank %d dims {%llu,%llu,%llu} stride1 %llu "
                  "box {%u,%u}",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides[0],
                  box[0], box[1]);
    return set_error(A3_ERR_CUDA, buf);
  }
  return A3_OK;
}

int make_tmap_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  return encode(tm, ptr, 2, dims, strides, box);
}

int make_tmap_3d_bf16(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                      uint32_t box0, uint32_t box1, uint32_t box2) {
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  return encode(tm, ptr, 3, dims, strides, box);
}

int make_tmap_4d_bf16(CUtensorMap* tm, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&strides_bytes)[3],
                      const uint32_t (&box)[4], int swizzle32) {
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  return encode(tm, ptr, 4, d, st, bx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_3d_f32_linear(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                            uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                            uint32_t box0, uint32_t box1, uint32_t box2) {
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  return encode(tm, ptr, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace a3

extern "C" int a3_set_sm_limit(int n) {
  if (n < 0 || n > a3::device_sms()) return a3::set_error(A3_ERR_ARG, "a3_set_sm_limit: 0 <= n <= SM count");
  a3::g_sm_limit.store(n & ~1);  // even: CTA pairs
  return A3_OK;
}

extern "C" const char* a3_last_error(void) { return a3::g_err; }
extern "C" int a3_version(void) { return 100; }
extern "C" long long a3_launch_count(void) { return a3::g_launches.load(); }
extern "C" void a3_reset_launch_count(void) { a3::g_launches.store(0); }

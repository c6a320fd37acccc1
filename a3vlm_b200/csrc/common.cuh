// Host-side helpers shared by every translation unit of liba3vlm_b200.so: error reporting behind
// the C ABI, launch checking, device properties, and TMA tensor-map encoding (the driver entry
// point is resolved at run time so the library carries no link-time dependency on libcuda).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "a3vlm_b200.h"

namespace a3 {

int set_error(int code, const char* msg);
int check_launch(const char* what);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: remembered per (kernel, current device),
// so a process that drives several GPUs configures each of them (raised only, never lowered).
int ensure_dyn_smem(const void* kernel, size_t bytes);
// The attention kernels place 128-byte-swizzled tiles at fixed offsets of the dynamic shared-memory base and have
// no room for alignment slack (224 of the 227 KB are tiles).  Probed once per device with a one-thread kernel
// (the first call synchronises); A3_ERR_UNSUPPORTED if the base is not 1024-byte aligned.
int check_dyn_smem_alignment(const char* who);
int num_sms();     // SMs the persistent kernels size their grids for (a3_set_sm_limit)
int device_sms();  // physical SM count

// 2-D bf16 tensor map with 128-byte swizzle.  dim0 is the contiguous dimension.
int make_tmap_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t stride1_bytes, uint32_t box0, uint32_t box1);
// 3-D bf16 tensor map with 128-byte swizzle (dim0 contiguous; strides for dim1, dim2 in bytes).
int make_tmap_3d_bf16(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                      uint32_t box0, uint32_t box1, uint32_t box2);

// 4-D bf16 tensor map with 128-byte (or 32-byte) swizzle (dims[0] contiguous; strides of dims 1..3 in bytes,
// multiples of 16; box origins must be 16-byte aligned in dim 0).
int make_tmap_4d_bf16(CUtensorMap* tm, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&strides_bytes)[3],
                      const uint32_t (&box)[4], int swizzle32 = 0);

// 3-D fp32 tensor map without swizzle (row-major smem box), used for TMA reduce-add stores.
int make_tmap_3d_f32_linear(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1,
                            uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                            uint32_t box0, uint32_t box1, uint32_t box2);

}  // namespace a3

// cds_common.cuh — shared host/device helpers for libcds_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/cds.h"

namespace cds {

void set_error(const char* fmt, ...);
int ensure_device();            // CDS_OK or CDS_ENODEVICE; never falls back to the CPU
void count_launch(int n = 1);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-(function, device) attribute: remember the largest value set for each pair
// (thread-safe; a second device or a second host thread in the same process gets its own cudaFuncSetAttribute).
int ensure_dyn_smem_impl(const void* func, size_t bytes);
template <typename K> inline int ensure_dyn_smem(K kernel, size_t bytes) { return ensure_dyn_smem_impl(reinterpret_cast<const void*>(kernel), bytes); }
int device_sm_count();          // multiprocessors of the current device (cached per device)

#ifdef __CUDACC__
// One 256-bit store (sm_100: STG.256): a thread that owns 32 contiguous bytes writes a whole 32-byte sector with one instruction --
// two 128-bit stores from lanes that are 64 bytes or a picture row apart leave every sector half written per instruction.
// `p` must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
#endif

#define CDS_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      cds::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CDS_ECUDA;                                                             \
    }                                                                               \
  } while (0)

#define CDS_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cds::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      cds::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return CDS_ECUDA;                                                             \
    }                                                                               \
  } while (0)

// Grow-only device scratch buffer (host-buffer entry points reuse it across calls).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
  ~DevBuf();
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// floats >= 0 order like their bit patterns -> atomicMax on unsigned gives a float max
__device__ __forceinline__ void atomic_max_nonneg(unsigned* addr, float v) {
  if (v > 0.f) atomicMax(addr, __float_as_uint(v));
}
// order-preserving float <-> unsigned key (all finite floats), for min/max of signed data
__device__ __forceinline__ unsigned float_key(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cds

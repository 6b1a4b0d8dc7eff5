// Shared helpers for the cellbender_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cb {

constexpr int kWarp = 32;

// ---- error plumbing for the C ABI (status int + last-error string) ----------------
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> status
int sm_count();                      // SMs of the current device (cached per device)

#define CB_REQUIRE(cond, ...)              \
  do {                                     \
    if (!(cond)) {                         \
      cb::set_last_error(__VA_ARGS__);     \
      return 1;                            \
    }                                      \
  } while (0)

// ---- warp / block reductions -----------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of N values per thread; result valid in every thread.  `smem` must
// hold N * 32 elements of T.  Deterministic (fixed tree).
template <typename T, int N>
__device__ __forceinline__ void block_sum(T (&v)[N], T* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect smem reuse across calls
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T x = (lane < nw) ? smem[i * 32 + lane] : T(0);
    v[i] = warp_sum(x);
  }
}

__device__ __forceinline__ float block_max(float v, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float x = (lane < nw) ? smem[lane] : -INFINITY;
  return warp_max(x);
}

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace cb

// Common device/host helpers for the dinox_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dinox_b200.h"

namespace dinox {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local last error string, returned through dinox_last_error_string)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what, cudaStream_t stream);
int require_sm100(void);
int num_sms(void);

#define DINOX_REQUIRE(cond, code, ...)                                                   \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::dinox::set_error(__VA_ARGS__);                                                   \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

#define DINOX_CUDA(call)                                                                 \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::dinox::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DINOX_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
#define DINOX_LOG2E 1.4426950408889634f
#define DINOX_LN2 0.6931471805599453f

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

// online (max, sum-of-exp2) pair merge; all quantities in log2 units
struct MaxSum {
  float m, s;
};
__device__ __forceinline__ MaxSum maxsum_merge(MaxSum a, MaxSum b) {
  float m = fmaxf(a.m, b.m);
  MaxSum r;
  r.m = m;
  // exp2f(-inf - -inf) guard: if m == -inf both sums are zero
  float sa = (a.m == -INFINITY) ? 0.f : a.s * exp2f(a.m - m);
  float sb = (b.m == -INFINITY) ? 0.f : b.s * exp2f(b.m - m);
  r.s = sa + sb;
  return r;
}
__device__ __forceinline__ MaxSum warp_maxsum(MaxSum v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum w;
    w.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    w.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    v = maxsum_merge(v, w);
  }
  return v;
}

// block-wide reductions through shared memory (blockDim.x multiple of 32, <= 1024).
// Fixed order => deterministic.
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float r = (threadIdx.x < kThreads / 32) ? smem[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) smem[0] = r;
  __syncthreads();
  return smem[0];
}
template <int kThreads>
__device__ __forceinline__ MaxSum block_maxsum(MaxSum v, float* smem /* >= 64 floats */) {
  v = warp_maxsum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { smem[w] = v.m; smem[32 + w] = v.s; }
  __syncthreads();
  MaxSum r;
  r.m = (threadIdx.x < kThreads / 32) ? smem[threadIdx.x] : -INFINITY;
  r.s = (threadIdx.x < kThreads / 32) ? smem[32 + threadIdx.x] : 0.f;
  if (w == 0) r = warp_maxsum(r);
  __syncthreads();
  if (threadIdx.x == 0) { smem[0] = r.m; smem[32] = r.s; }
  __syncthreads();
  r.m = smem[0];
  r.s = smem[32];
  return r;
}

// element loads as float for the three logit dtypes of the materialised-logit path
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// 128-bit streaming loads / stores
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace dinox

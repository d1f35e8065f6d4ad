// Shared helpers for the lt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/lt_b200.h"

namespace lt {

// ---- error reporting ------------------------------------------------------------------------
char* err_buf();  // thread-local 512-byte buffer (capi.cu)
int fail(int code, const char* fmt, ...);

#define LT_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::lt::fail(LT_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define LT_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) return ::lt::fail(LT_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ---- split-bf16 ("S32") format ----------------------------------------------------------------
// x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi): 16 significand bits, fp32 range.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}
// element offset (in bf16 units) of the high part of channel c of a pixel whose row starts at 0;
// low part is +32.
__device__ __host__ __forceinline__ int s32_off(int c) { return ((c >> 5) << 6) + (c & 31); }

// 4 consecutive channels (c % 4 == 0) <-> split storage
__device__ __forceinline__ void store_s32x4(__nv_bfloat16* row, int c, float4 v) {
  __nv_bfloat16 h[4], l[4];
  split_bf16(v.x, h[0], l[0]);
  split_bf16(v.y, h[1], l[1]);
  split_bf16(v.z, h[2], l[2]);
  split_bf16(v.w, h[3], l[3]);
  __nv_bfloat16* p = row + s32_off(c);
  *reinterpret_cast<uint2*>(p) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(p + 32) = *reinterpret_cast<uint2*>(l);
}
__device__ __forceinline__ float4 load_s32x4(const __nv_bfloat16* row, int c) {
  const __nv_bfloat16* p = row + s32_off(c);
  uint2 hu = *reinterpret_cast<const uint2*>(p);
  uint2 lu = *reinterpret_cast<const uint2*>(p + 32);
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&hu);
  const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&lu);
  return make_float4(join_bf16(h[0], l[0]), join_bf16(h[1], l[1]), join_bf16(h[2], l[2]), join_bf16(h[3], l[3]));
}

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

}  // namespace lt

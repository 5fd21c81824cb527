// common.cuh - error plumbing and small device helpers shared by every libgvit translation unit.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gvit.h"

namespace gvit {

// thread-local last-error text (autograd calls us from a second thread)
void set_error(const char* fmt, ...);
int fail(gvit_status s, const char* fmt, ...);

#define GVIT_CHECK_CUDA(expr)                                                                     \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return ::gvit::fail(GVIT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                                    \
  } while (0)

#define GVIT_CHECK_LAUNCH() GVIT_CHECK_CUDA(cudaGetLastError())

#define GVIT_REQUIRE(cond, status, ...)                  \
  do {                                                   \
    if (!(cond)) return ::gvit::fail(status, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- element access -------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats, 16-byte (bf16) or 2x16-byte (fp32) transactions
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
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

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Launch a grid-stride streaming kernel (256 threads, one 8-element group per thread and trip) as exactly ONE wave:
// grid = min(blocks needed, SMs x resident CTAs of THIS kernel).  A fixed "8 CTAs per SM" guess put a 40-register kernel
// (6 resident CTAs) at 1.33 waves: the last third of the CTAs ran alone at a third of the occupancy.
template <typename... KArgs, typename... Args>
inline void stream_launch(void (*kernel)(KArgs...), int64_t groups, cudaStream_t st, Args... args) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  const int64_t blocks = (groups + 255) / 256, cap = (int64_t)num_sms() * per_sm;
  const int grid = (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  kernel<<<grid, 256, 0, st>>>(args...);
}

}  // namespace gvit

// edges.cu - the bandwidth-bound edges of the block: LayerNorm (nn.LayerNorm at
// /root/reference/src/models/vit.py:103,108,154) and proj_drop + residual add (vit.py:71,117).
// One warp per token row, 16-byte loads/stores, fp32 statistics, no shared-memory staging
// (each element is touched once).
#include "kernels.cuh"

namespace gvit {
namespace {

constexpr int MAXC = 4;  // D <= 1024

template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                     const T* __restrict__ beta, int64_t rows, int D, float eps,
                                                     T* __restrict__ y, float* __restrict__ mean,
                                                     float* __restrict__ rstd) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[MAXC][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
      load8(x + row * D + d0, v[c]);
#pragma unroll
      for (int t = 0; t < 8; ++t) s += v[c][t];
    }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float dlt = v[c][t] - mu; q = fmaf(dlt, dlt, q); }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
      float g[8], bt[8], o[8];
      load8(gamma + d0, g);
      load8(beta + d0, bt);
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = fmaf((v[c][t] - mu) * rs, g[t], bt[t]);
      store8(y + row * D + d0, o);
    }
  }
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// dx for every row; per-CTA partial dgamma/dbeta (fixed row -> CTA assignment: deterministic)
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                     const T* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, int64_t rows, int D,
                                                     T* __restrict__ dx, float* __restrict__ partial) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dg[MAXC][8] = {}, db[MAXC][8] = {}, g[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) load8(gamma + d0, g[c]);
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[MAXC][8], gy[MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float xv[8], dyv[8];
        load8(x + row * D + d0, xv);
        load8(dy + row * D + d0, dyv);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          xh[c][t] = (xv[t] - mu) * rs;
          gy[c][t] = dyv[t] * g[c][t];
          s1 += gy[c][t];
          s2 = fmaf(gy[c][t], xh[c][t], s2);
          dg[c][t] = fmaf(dyv[t], xh[c][t], dg[c][t]);
          db[c][t] += dyv[t];
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = rs * (gy[c][t] - s1 - xh[c][t] * s2);
        store8(dx + row * D + d0, o);
      }
    }
  }
  // reduce the 8 warps' partials column by column through shared memory, 256 columns at a time
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c * 256 >= D) break;
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 8; ++t) red[warp][lane * 8 + t] = pass == 0 ? dg[c][t] : db[c][t];
      __syncthreads();
      const int col = c * 256 + threadIdx.x;
      if (col < D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        partial[((int64_t)pass * gridDim.x + blockIdx.x) * D + col] = s;
      }
    }
  }
}

__global__ void ln_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int D, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= D) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < nblk; ++i) {
    a += partial[(int64_t)i * D + col];
    b += partial[((int64_t)nblk + i) * D + col];
  }
  dgamma[col] = a;
  dbeta[col] = b;
}

// ---- Philox-4x32-10 (Salmon et al.), counter = (offset + i/4), key = seed --------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_residual_fwd_kernel(const T* __restrict__ y, const T* __restrict__ resid,
                                                                   int64_t n, float p, uint64_t seed, uint64_t offset,
                                                                   T* __restrict__ out, uint8_t* __restrict__ mask) {
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += (int64_t)gridDim.x * blockDim.x * 8) {
    float a[8], r[8] = {};
    load8(y + i8, a);
    if (resid) load8(resid + i8, r);
    if (p > 0.f) {
      uint8_t keep[8];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const uint64_t c = offset + (uint64_t)(i8 / 4 + hlf);
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t u[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) keep[hlf * 4 + t] = (u[t] >> 8) * (1.0f / 16777216.0f) >= p;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) a[t] = keep[t] ? a[t] * scale : 0.f;
      *reinterpret_cast<uint2*>(mask + i8) = *reinterpret_cast<const uint2*>(keep);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] += r[t];
    store8(out + i8, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ mask,
                                                          int64_t n, float p, T* __restrict__ dy) {
  const float scale = 1.0f / (1.0f - p);
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += (int64_t)gridDim.x * blockDim.x * 8) {
    float a[8];
    load8(dout + i8, a);
    const uint2 raw = *reinterpret_cast<const uint2*>(mask + i8);
    const uint8_t* keep = reinterpret_cast<const uint8_t*>(&raw);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = keep[t] ? a[t] * scale : 0.f;
    store8(dy + i8, a);
  }
}

inline int stream_grid(int64_t n8) {
  const int64_t blocks = (n8 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, int dtype,
                  void* y, float* mean, float* rstd, cudaStream_t st) {
  const int blocks = (int)((rows + 7) / 8);
  if (dtype == GVIT_F32)
    ln_fwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(gamma),
                                                 static_cast<const float*>(beta), rows, D, eps, static_cast<float*>(y),
                                                 mean, rstd);
  else
    ln_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gamma),
        static_cast<const __nv_bfloat16*>(beta), rows, D, eps, static_cast<__nv_bfloat16*>(y), mean, rstd);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, int64_t rows,
                  int D, int dtype, void* dx, float* dgamma, float* dbeta, float* partial_ws, cudaStream_t st) {
  int nblk = (int)((rows + 7) / 8);
  if (nblk > GVIT_LN_PARTIALS) nblk = GVIT_LN_PARTIALS;
  if (dtype == GVIT_F32)
    ln_bwd_kernel<float><<<nblk, 256, 0, st>>>(static_cast<const float*>(dy), static_cast<const float*>(x),
                                               static_cast<const float*>(gamma), mean, rstd, rows, D,
                                               static_cast<float*>(dx), partial_ws);
  else
    ln_bwd_kernel<__nv_bfloat16><<<nblk, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
        static_cast<const __nv_bfloat16*>(gamma), mean, rstd, rows, D, static_cast<__nv_bfloat16*>(dx), partial_ws);
  GVIT_CHECK_LAUNCH();
  ln_bwd_reduce_kernel<<<(D + 255) / 256, 256, 0, st>>>(partial_ws, nblk, D, dgamma, dbeta);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dropout_residual_fwd(const void* y, const void* resid, int64_t n, float p, uint64_t seed, uint64_t offset,
                         int dtype, void* out, uint8_t* keep_mask, cudaStream_t st) {
  const int grid = stream_grid(n / 8);
  if (dtype == GVIT_F32)
    dropout_residual_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(y),
                                                             static_cast<const float*>(resid), n, p, seed, offset,
                                                             static_cast<float*>(out), keep_mask);
  else
    dropout_residual_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(resid), n, p, seed, offset,
        static_cast<__nv_bfloat16*>(out), keep_mask);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dropout_bwd(const void* dout, const uint8_t* keep_mask, int64_t n, float p, int dtype, void* dy, cudaStream_t st) {
  const int grid = stream_grid(n / 8);
  if (dtype == GVIT_F32)
    dropout_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dout), keep_mask, n, p,
                                                    static_cast<float*>(dy));
  else
    dropout_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), keep_mask, n, p,
                                                            static_cast<__nv_bfloat16*>(dy));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

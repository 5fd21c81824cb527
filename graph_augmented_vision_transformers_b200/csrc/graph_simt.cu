// graph_simt.cu - the gather / scatter side of the graph sub-layer:
//   G4+G5 forward (softmax over the k selected similarities, weighted gather of neighbour tokens),
//   the reverse adjacency (CSR) that turns every backward scatter into a deterministic gather,
//   and the backward of G4+G5 and of G1-G3.
// All of these are HBM/L2-bound row operations: one warp per token row, 16-byte loads, no atomics.
#include <float.h>

#include "kernels.cuh"

namespace gvit {
namespace {

constexpr int MAXC = 4;  // D <= 1024: a lane owns up to 4 chunks of 8 consecutive features

template <typename T> __device__ __forceinline__ float round_like(float v) { return to_f32(from_f32<T>(v)); }

// ---- G4 + G5 forward -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) agg_gather_fwd_kernel(const T* __restrict__ p, int64_t bs, int64_t rs, int B,
                                                             int Np, int D, int k, const int32_t* __restrict__ idx,
                                                             const float* __restrict__ vals, float* __restrict__ w,
                                                             T* __restrict__ z) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * Np) return;
  const int b = row / Np;
  const float v = lane < k ? vals[(int64_t)row * k + lane] : -FLT_MAX;
  const int nb = lane < k ? idx[(int64_t)row * k + lane] : 0;
  const float m = warp_max(v);
  const float e = lane < k ? expf(v - m) : 0.f;
  const float wl = e / warp_sum(e);
  if (lane < k) w[(int64_t)row * k + lane] = wl;
  const float wq = round_like<T>(wl);   // bf16 path: the GEMM-shaped aggregation sees bf16 weights (autocast)
  const T* img = p + b * bs;
  for (int d0 = lane * 8; d0 < D; d0 += 256) {
    float acc[8] = {};
    for (int j = 0; j < k; ++j) {
      const float wj = __shfl_sync(0xffffffffu, wq, j);
      const int nj = __shfl_sync(0xffffffffu, nb, j);
      float x[8];
      load8(img + (int64_t)nj * rs + d0, x);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(wj, x[t], acc[t]);
    }
    store8(z + (int64_t)row * D + d0, acc);
  }
}

// ---- reverse adjacency -------------------------------------------------------------------------
// One CTA per image.  Thread j scans the image's Np*k edges in ascending edge id, so each reverse list
// is sorted and every backward reduction has a fixed summation order.
__global__ void graph_reverse_kernel(const int32_t* __restrict__ idx, int Np, int k, int32_t* __restrict__ rev_ptr,
                                     int32_t* __restrict__ rev_src) {
  extern __shared__ int32_t sm[];
  int32_t* e_dst = sm;                 // Np*k
  int32_t* cnt = sm + (int64_t)Np * k; // Np + 1
  const int b = blockIdx.x, E = Np * k;
  const int32_t* src = idx + (int64_t)b * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) e_dst[e] = src[e];
  __syncthreads();
  for (int j = threadIdx.x; j < Np; j += blockDim.x) {
    int c = 0;
    for (int e = 0; e < E; ++e) c += (e_dst[e] == j);
    cnt[j] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int j = 0; j < Np; ++j) { const int c = cnt[j]; cnt[j] = run; run += c; }
    cnt[Np] = run;
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= Np; j += blockDim.x) rev_ptr[(int64_t)b * (Np + 1) + j] = cnt[j];
  for (int j = threadIdx.x; j < Np; j += blockDim.x) {
    int o = cnt[j];
    int32_t* dst = rev_src + (int64_t)b * E;
    for (int e = 0; e < E; ++e)
      if (e_dst[e] == j) dst[o++] = e;
  }
}

// ---- backward of G4 + G5 -------------------------------------------------------------------------
// phase A (warp per destination row i): dw_ij = dz_i . p[idx_ij];  dvals = softmax backward.
template <typename T>
__global__ void __launch_bounds__(256) agg_bwd_dvals_kernel(const T* __restrict__ p, int64_t bs, int64_t rs, int B,
                                                            int Np, int D, int k, const int32_t* __restrict__ idx,
                                                            const float* __restrict__ w, const T* __restrict__ dz,
                                                            float* __restrict__ dvals) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * Np) return;
  const int b = row / Np;
  float g[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) load8(dz + (int64_t)row * D + d0, g[c]);
  }
  const int nb = lane < k ? idx[(int64_t)row * k + lane] : 0;
  const T* img = p + b * bs;
  float dw = 0.f;
  for (int j = 0; j < k; ++j) {
    const int nj = __shfl_sync(0xffffffffu, nb, j);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float x[8];
        load8(img + (int64_t)nj * rs + d0, x);
#pragma unroll
        for (int t = 0; t < 8; ++t) dot = fmaf(g[c][t], x[t], dot);
      }
    }
    dot = warp_sum(dot);
    if (lane == j) dw = dot;
  }
  const float wl = lane < k ? w[(int64_t)row * k + lane] : 0.f;
  const float s = warp_sum(wl * dw);
  if (lane < k) dvals[(int64_t)row * k + lane] = wl * (dw - s);
}

// phase B (warp per source row j): dp_j = sum over edges e=(i,s) that point at j of w_e * dz_i.
template <typename T>
__global__ void __launch_bounds__(256) agg_bwd_dp_kernel(int64_t bs, int64_t rs, int B, int Np, int D, int k,
                                                         const float* __restrict__ w, const T* __restrict__ dz,
                                                         const int32_t* __restrict__ rev_ptr,
                                                         const int32_t* __restrict__ rev_src, T* __restrict__ dp) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * Np) return;
  const int b = row / Np, j = row % Np;
  const int beg = rev_ptr[(int64_t)b * (Np + 1) + j], end = rev_ptr[(int64_t)b * (Np + 1) + j + 1];
  const int32_t* edges = rev_src + (int64_t)b * Np * k;
  const float* wimg = w + (int64_t)b * Np * k;
  float acc[MAXC][8] = {};
  for (int q = beg; q < end; ++q) {
    const int e = edges[q];
    const float we = round_like<T>(wimg[e]);
    const T* src = dz + ((int64_t)b * Np + e / k) * D;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float x[8];
        load8(src + d0, x);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[c][t] = fmaf(we, x[t], acc[c][t]);
      }
    }
  }
  T* dst = dp + b * bs + (int64_t)j * rs;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) store8(dst + d0, acc[c]);
  }
}

// ---- backward of G1-G3 -------------------------------------------------------------------------
// S_ij = p^_i . p^_j with p^ = p * rnorm.  With dS sparse (k entries per row):
//   dp^_j = sum_s dS_js p^_{idx_js}  +  sum_{e=(i,s) -> j} dS_e p^_i
//   dp_j += rnorm_j * (dp^_j - p^_j (p^_j . dp^_j))
template <typename T>
__global__ void __launch_bounds__(256) knn_bwd_kernel(const T* __restrict__ p, int64_t bs, int64_t rs, int B, int Np,
                                                      int D, int k, const int32_t* __restrict__ idx,
                                                      const float* __restrict__ rnorm, const float* __restrict__ dvals,
                                                      const int32_t* __restrict__ rev_ptr,
                                                      const int32_t* __restrict__ rev_src, T* __restrict__ dp) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * Np) return;
  const int b = row / Np, j = row % Np;
  const T* img = p + b * bs;
  const float* rn = rnorm + (int64_t)b * Np;
  const float* dsi = dvals + (int64_t)b * Np * k;
  const int32_t* ii = idx + (int64_t)b * Np * k;
  float acc[MAXC][8] = {};
  auto add_row = [&](int src_row, float coeff) {
    const float c2 = coeff * rn[src_row];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float x[8];
        load8(img + (int64_t)src_row * rs + d0, x);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[c][t] = fmaf(c2, x[t], acc[c][t]);
      }
    }
  };
  for (int s = 0; s < k; ++s) add_row(ii[(int64_t)j * k + s], dsi[(int64_t)j * k + s]);
  const int beg = rev_ptr[(int64_t)b * (Np + 1) + j], end = rev_ptr[(int64_t)b * (Np + 1) + j + 1];
  const int32_t* edges = rev_src + (int64_t)b * Np * k;
  for (int q = beg; q < end; ++q) { const int e = edges[q]; add_row(e / k, dsi[e]); }

  const float rnj = rn[j];
  float own[MAXC][8];
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
      load8(img + (int64_t)j * rs + d0, own[c]);
#pragma unroll
      for (int t = 0; t < 8; ++t) { own[c][t] *= rnj; dot = fmaf(own[c][t], acc[c][t], dot); }
    }
  }
  dot = warp_sum(dot);
  T* dst = dp + b * bs + (int64_t)j * rs;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
      float cur[8];
      load8(dst + d0, cur);
#pragma unroll
      for (int t = 0; t < 8; ++t) cur[t] += rnj * (acc[c][t] - own[c][t] * dot);
      store8(dst + d0, cur);
    }
  }
}

inline int row_blocks(int rows) { return (rows + 7) / 8; }

}  // namespace

int agg_gather_fwd_simt(const Tokens& t, int k, int dtype, const int32_t* idx, const float* vals, float* w, void* z,
                        cudaStream_t st) {
  const int rows = t.B * t.Np;
  if (dtype == GVIT_F32)
    agg_gather_fwd_kernel<float><<<row_blocks(rows), 256, 0, st>>>(static_cast<const float*>(t.ptr), t.batch_stride,
                                                                   t.row_stride, t.B, t.Np, t.D, k, idx, vals, w,
                                                                   static_cast<float*>(z));
  else
    agg_gather_fwd_kernel<__nv_bfloat16><<<row_blocks(rows), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(t.ptr), t.batch_stride, t.row_stride, t.B, t.Np, t.D, k, idx, vals, w,
        static_cast<__nv_bfloat16*>(z));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int graph_reverse(const int32_t* idx, int B, int Np, int k, int32_t* rev_ptr, int32_t* rev_src, cudaStream_t st) {
  const size_t smem = ((size_t)Np * k + Np + 1) * sizeof(int32_t);
  GVIT_REQUIRE(smem <= 200 * 1024, GVIT_ERR_SHAPE, "graph_reverse: Np*k=%d edges exceed shared memory", Np * k);
  if (smem > 48 * 1024)
    GVIT_CHECK_CUDA(cudaFuncSetAttribute(graph_reverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = Np <= 256 ? 256 : (Np <= 512 ? 512 : 1024);
  graph_reverse_kernel<<<B, threads, smem, st>>>(idx, Np, k, rev_ptr, rev_src);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int agg_bwd_simt(const Tokens& t, int k, int dtype, const int32_t* idx, const float* w, const void* dz,
                 const int32_t* rev_ptr, const int32_t* rev_src, float* dvals, void* dp, cudaStream_t st) {
  const int rows = t.B * t.Np;
  if (dtype == GVIT_F32) {
    agg_bwd_dvals_kernel<float><<<row_blocks(rows), 256, 0, st>>>(static_cast<const float*>(t.ptr), t.batch_stride,
                                                                  t.row_stride, t.B, t.Np, t.D, k, idx, w,
                                                                  static_cast<const float*>(dz), dvals);
    GVIT_CHECK_LAUNCH();
    agg_bwd_dp_kernel<float><<<row_blocks(rows), 256, 0, st>>>(t.batch_stride, t.row_stride, t.B, t.Np, t.D, k, w,
                                                               static_cast<const float*>(dz), rev_ptr, rev_src,
                                                               static_cast<float*>(dp));
  } else {
    using bf = __nv_bfloat16;
    agg_bwd_dvals_kernel<bf><<<row_blocks(rows), 256, 0, st>>>(static_cast<const bf*>(t.ptr), t.batch_stride,
                                                               t.row_stride, t.B, t.Np, t.D, k, idx, w,
                                                               static_cast<const bf*>(dz), dvals);
    GVIT_CHECK_LAUNCH();
    agg_bwd_dp_kernel<bf><<<row_blocks(rows), 256, 0, st>>>(t.batch_stride, t.row_stride, t.B, t.Np, t.D, k, w,
                                                            static_cast<const bf*>(dz), rev_ptr, rev_src,
                                                            static_cast<bf*>(dp));
  }
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int knn_bwd_simt(const Tokens& t, int k, int dtype, const int32_t* idx, const float* rnorm, const float* dvals,
                 const int32_t* rev_ptr, const int32_t* rev_src, void* dp, cudaStream_t st) {
  const int rows = t.B * t.Np;
  if (dtype == GVIT_F32)
    knn_bwd_kernel<float><<<row_blocks(rows), 256, 0, st>>>(static_cast<const float*>(t.ptr), t.batch_stride,
                                                            t.row_stride, t.B, t.Np, t.D, k, idx, rnorm, dvals,
                                                            rev_ptr, rev_src, static_cast<float*>(dp));
  else
    knn_bwd_kernel<__nv_bfloat16><<<row_blocks(rows), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(t.ptr), t.batch_stride, t.row_stride, t.B, t.Np, t.D, k, idx, rnorm, dvals,
        rev_ptr, rev_src, static_cast<__nv_bfloat16*>(dp));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

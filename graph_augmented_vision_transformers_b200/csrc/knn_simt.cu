// knn_simt.cu - graph construction G1-G3 with exact fp32 FMA arithmetic.
//
// The fp32 parity path follows the STRICT specification oracle/knn_strict.c (GRAPH_SPEC_VERSION 2) operation by
// operation, so neighbour indices AND similarities are bit-identical to it on every row:
//   G1  ss_i = sequential FMA chain over d of p_id^2;  n_i = max(sqrt(ss_i), 1e-12);  ph_id = p_id / n_i  (IEEE division)
//   G2  S_ij = sequential FMA chain over d = 0..D-1 of ph_id * ph_jd - the same order for every (i, j), so S is exactly
//       symmetric and duplicated token rows give bit-equal similarities
//   G3  strict-">" insertion while sweeping the columns in ascending order: ties resolve to the lowest index, exactly
//       like the oracle's stable sort.
// (Compiled without --use_fast_math: `/` and sqrtf are the correctly rounded IEEE operations.)
// S is never written to HBM: a CTA owns 64 rows of one image, sweeps the columns in 64-wide tiles
// and keeps a running top-k per row in shared memory.
#include <float.h>

#include "kernels.cuh"

namespace gvit {
namespace {

constexpr int TM = 64, TN = 64, BK = 16, PAD = 4;
constexpr int KSTRIDE = GVIT_MAX_K + 1;  // +1: row-strided top-k lists would otherwise share a bank

// ss_i of G1: ONE thread per token row, sequential FMA chain over d (the order the specification fixes)
template <typename T>
__global__ void __launch_bounds__(128) sumsq_kernel(const T* __restrict__ p, int64_t bs, int64_t rs, int B, int Np, int D,
                                                    float* __restrict__ ss_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * Np) return;
  const int b = r / Np, i = r % Np;
  const T* row = p + b * bs + i * rs;
  float acc = 0.f;
  int d = 0;
  for (; d + 8 <= D; d += 8) {
    float v[8];
    load8(row + d, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc = fmaf(v[t], v[t], acc);
  }
  for (; d < D; ++d) { const float v = to_f32(row[d]); acc = fmaf(v, v, acc); }
  ss_out[r] = acc;
}

// n_i = max(sqrt(ss_i), eps) (F.normalize's clamp)
__device__ __forceinline__ float norm_of(float ss) { return fmaxf(sqrtf(ss), 1e-12f); }

// the ABI's rnorm output: 1 / n_i, written in place over ss once every similarity tile has been formed
__global__ void __launch_bounds__(256) rnorm_finish_kernel(float* __restrict__ r, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = 1.0f / norm_of(r[i]);
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T>
__global__ void __launch_bounds__(256) sim_topk_kernel(const T* __restrict__ p, int64_t bs, int64_t rs, int Np, int D,
                                                       int k, const float* __restrict__ sumsq,
                                                       int32_t* __restrict__ idx, float* __restrict__ vals) {
  __shared__ float As[BK][TM + PAD];
  __shared__ float Bs[BK][TN + PAD];
  __shared__ float Ss[TM][TN + 1];
  __shared__ float topv[TM][KSTRIDE];
  __shared__ int topi[TM][KSTRIDE];

  const int b = blockIdx.y, r0 = blockIdx.x * TM, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const T* img = p + b * bs;
  const float* ss = sumsq + (int64_t)b * Np;

  if (tid < TM)
    for (int s = 0; s < k; ++s) { topv[tid][s] = -FLT_MAX; topi[tid][s] = 0x7fffffff; }

  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // tile loader: 64 rows x 16 k, 4 elements per thread
  const float na = r0 + lrow < Np ? norm_of(ss[r0 + lrow]) : 1.f;
  for (int c0 = 0; c0 < Np; c0 += TN) {
    const float nb = c0 + lrow < Np ? norm_of(ss[c0 + lrow]) : 1.f;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < D; k0 += BK) {
      float a[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
      if (r0 + lrow < Np && k0 + lk < D) load4<T>(img + (int64_t)(r0 + lrow) * rs + k0 + lk, a);
      if (c0 + lrow < Np && k0 + lk < D) load4<T>(img + (int64_t)(c0 + lrow) * rs + k0 + lk, bb);
#pragma unroll
      for (int t = 0; t < 4; ++t) { a[t] = a[t] / na; bb[t] = bb[t] / nb; }   // G1: normalise FIRST (IEEE division)
      __syncthreads();   // previous tile fully consumed
#pragma unroll
      for (int t = 0; t < 4; ++t) { As[lk + t][lrow] = a[t]; Bs[lk + t][lrow] = bb[t]; }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[4], bv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { av[t] = As[kk][ty * 4 + t]; bv[t] = Bs[kk][tx * 4 + t]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + tx * 4 + j;
        Ss[ty * 4 + i][tx * 4 + j] = c < Np ? acc[i][j] : -FLT_MAX;
      }
    }
    __syncthreads();
    if (tid < TM && r0 + tid < Np) {
      float* tv = topv[tid];
      int* ti = topi[tid];
      const int cmax = min(TN, Np - c0);
      for (int c = 0; c < cmax; ++c) {
        const float v = Ss[tid][c];
        if (v > tv[k - 1]) {          // strict: an equal later column never displaces an earlier one
          int pos = k - 1;
          while (pos > 0 && v > tv[pos - 1]) { tv[pos] = tv[pos - 1]; ti[pos] = ti[pos - 1]; --pos; }
          tv[pos] = v;
          ti[pos] = c0 + c;
        }
      }
    }
    // the next tile's first __syncthreads() orders these reads before Ss is rewritten
  }
  __syncthreads();
  if (tid < TM && r0 + tid < Np) {
    const int64_t o = ((int64_t)b * Np + r0 + tid) * k;
    for (int s = 0; s < k; ++s) { idx[o + s] = topi[tid][s]; vals[o + s] = topv[tid][s]; }
  }
}

template <typename T>
int launch(const Tokens& t, int k, int32_t* idx, float* vals, float* rnorm, cudaStream_t st) {
  const T* p = static_cast<const T*>(t.ptr);
  const int rows = t.B * t.Np;
  sumsq_kernel<T><<<(rows + 127) / 128, 128, 0, st>>>(p, t.batch_stride, t.row_stride, t.B, t.Np, t.D, rnorm);   // rnorm holds ss_i for now
  GVIT_CHECK_LAUNCH();
  dim3 grid((t.Np + TM - 1) / TM, t.B);
  sim_topk_kernel<T><<<grid, 256, 0, st>>>(p, t.batch_stride, t.row_stride, t.Np, t.D, k, rnorm, idx, vals);
  GVIT_CHECK_LAUNCH();
  rnorm_finish_kernel<<<(rows + 255) / 256, 256, 0, st>>>(rnorm, rows);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

int knn_fwd_simt(const Tokens& p, int k, int dtype, int32_t* idx, float* vals, float* rnorm, cudaStream_t st) {
  return dtype == GVIT_F32 ? launch<float>(p, k, idx, vals, rnorm, st) : launch<__nv_bfloat16>(p, k, idx, vals, rnorm, st);
}

}  // namespace gvit

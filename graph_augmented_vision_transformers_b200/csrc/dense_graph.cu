// dense_graph.cu - the row-wise stages of the DENSE-adjacency graph layer (SURVEY.md section 9 with G3 skipped: every patch
// token attends to every patch token of its image; BASELINE configs[3]).  The matrix products around them are
// bgemm_tc.cu; everything here is one warp per row, 16-byte accesses, fp32 arithmetic.
//
//   forward : rn_i = 1 / max(||p_i||, 1e-12)                                   (G1)            dense_rownorm
//             S_ij = G_ij rn_i rn_j,  A~_i. = softmax_j(S_i.)  -> bf16          (G2, G4)        dense_softmax_fwd
//   backward: delta_i = sum_j dA~_ij A~_ij,  dS_ij = A~_ij (dA~_ij - delta_i),  dG_ij = dS_ij rn_i rn_j   dense_softmax_bwd
//             dp_i = T_i + V_i - rn_i^2 (p_i . V_i) p_i                                        dense_combine_bwd
//             with T = A~^T dZ (aggregation backward) and V = (dG + dG^T) P (similarity backward; the last term is the
//             backward of the L2 normalisation: the component of V_i along p_i is removed).
#include <float.h>

#include "kernels.cuh"

namespace gvit {
namespace {

using bf = __nv_bfloat16;

__global__ void __launch_bounds__(256) dense_rownorm_kernel(const bf* __restrict__ p, int64_t bs, int64_t rs, int B, int Np, int D,
                                                            float* __restrict__ rn) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Np) return;
  const int b = w / Np, i = w % Np;
  const bf* row = p + b * bs + i * rs;
  float acc = 0.f;
  for (int d0 = lane * 8; d0 < D; d0 += 256) {
    float v[8];
    load8(row + d0, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc = fmaf(v[t], v[t], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) rn[w] = 1.0f / fmaxf(sqrtf(acc), 1e-12f);
}

// one warp per row i of image b: columns j = lane, lane + 32, ... (coalesced 128-byte segments of the fp32 Gram row)
__global__ void __launch_bounds__(256) dense_softmax_fwd_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ rn,
                                                                int B, int Np, int ldA, bf* __restrict__ A) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Np) return;
  const int b = w / Np;
  const float* g = G + (int64_t)w * ldg;
  const float* rnb = rn + (int64_t)b * Np;
  const float rni = rn[w];
  constexpr int MAXC = 32;                                   // Np <= 1024
  float s[MAXC];
  float mx = -FLT_MAX;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    s[c] = j < Np ? g[j] * rni * rnb[j] : -FLT_MAX;
    mx = fmaxf(mx, s[c]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    s[c] = j < Np ? __expf(s[c] - mx) : 0.f;
    sum += s[c];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  bf* a = A + (int64_t)w * ldA;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    if (j < ldA) a[j] = __float2bfloat16_rn(s[c] * inv);     // the pad columns [Np, ldA) are written as zeros
  }
}

__global__ void __launch_bounds__(256) dense_softmax_bwd_kernel(const float* __restrict__ dA, int ldg, const bf* __restrict__ A, int ldA,
                                                                const float* __restrict__ rn, int B, int Np, bf* __restrict__ dG) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Np) return;
  const int b = w / Np;
  const float* da = dA + (int64_t)w * ldg;
  const bf* a = A + (int64_t)w * ldA;
  const float* rnb = rn + (int64_t)b * Np;
  const float rni = rn[w];
  constexpr int MAXC = 32;
  float av[MAXC], dv[MAXC];
  float delta = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    av[c] = j < Np ? __bfloat162float(a[j]) : 0.f;
    dv[c] = j < Np ? da[j] : 0.f;
    delta = fmaf(av[c], dv[c], delta);
  }
  delta = warp_sum(delta);
  bf* o = dG + (int64_t)w * ldA;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    if (j < ldA) o[j] = __float2bfloat16_rn(j < Np ? av[c] * (dv[c] - delta) * rni * rnb[j] : 0.f);
  }
}

// dp_i = T_i + V_i - rn_i^2 (p_i . V_i) p_i ; T, V contiguous (B, Np, D) bf16; p / dp strided token views
__global__ void __launch_bounds__(256) dense_combine_bwd_kernel(const bf* __restrict__ T, const bf* __restrict__ V, const bf* __restrict__ p,
                                                                int64_t bs, int64_t rs, const float* __restrict__ rn, int B, int Np, int D,
                                                                bf* __restrict__ dp) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Np) return;
  const int b = w / Np, i = w % Np;
  const bf* prow = p + b * bs + i * rs;
  const bf* vrow = V + (int64_t)w * D;
  const bf* trow = T + (int64_t)w * D;
  bf* drow = dp + b * bs + i * rs;
  constexpr int MAXG = 4;                                    // D <= 1024: up to 4 groups of 8 elements per lane
  float pv[MAXG][8], vv[MAXG][8];
  float c = 0.f;
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    const int d0 = (g * 32 + lane) * 8;
    if (d0 < D) {
      load8(prow + d0, pv[g]);
      load8(vrow + d0, vv[g]);
#pragma unroll
      for (int t = 0; t < 8; ++t) c = fmaf(pv[g][t], vv[g][t], c);
    }
  }
  c = warp_sum(c);
  const float k = rn[w] * rn[w] * c;
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    const int d0 = (g * 32 + lane) * 8;
    if (d0 < D) {
      float tv[8], o[8];
      load8(trow + d0, tv);
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = tv[t] + vv[g][t] - k * pv[g][t];
      store8(drow + d0, o);
    }
  }
}

// G3 for more than 256 patch tokens: per-row top-k of S_ij = (G_ij rn_i) rn_j over a materialised fp32 Gram row (gvit_bgemm),
// one warp per row.  Lane l holds columns l, l + 32, ...; k rounds of (lane-local best, warp arg-max, winner retires).  Order:
// larger similarity first, at equal similarity the LOWER column (section 9 G3: the stable descending sort) - the lane-local
// scan keeps the first maximum (columns ascend with the slot) and the butterfly prefers the lower column on ties.
__global__ void __launch_bounds__(256) knn_select_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ rn, int B, int Np,
                                                         int k, int32_t* __restrict__ idx, float* __restrict__ vals) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Np) return;
  const int b = w / Np;
  const float* g = G + (int64_t)w * ldg;
  const float* rnb = rn + (int64_t)b * Np;
  const float rni = rn[w];
  constexpr int MAXC = 32;                                   // Np <= 1024
  float s[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = c * 32 + lane;
    s[c] = j < Np ? (g[j] * rni) * rnb[j] : -FLT_MAX;
  }
  uint32_t taken = 0;
  const int64_t o = (int64_t)w * k;
  for (int r = 0; r < k; ++r) {
    float bv = -FLT_MAX;
    int bj = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const bool free_ = !((taken >> c) & 1u) && c * 32 + lane < Np;
      if (free_ && (s[c] > bv || bj == 0x7fffffff)) { bv = s[c]; bj = c * 32 + lane; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
      if (oj != 0x7fffffff && (bj == 0x7fffffff || ov > bv || (ov == bv && oj < bj))) { bv = ov; bj = oj; }
    }
    if ((bj & 31) == lane) taken |= 1u << (bj >> 5);
    if (lane == 0) { idx[o + r] = bj; vals[o + r] = bv; }
  }
}

inline unsigned warps_grid(int64_t rows) { return (unsigned)((rows * 32 + 255) / 256); }

}  // namespace

int dense_rownorm(const Tokens& t, float* rn, cudaStream_t st) {
  dense_rownorm_kernel<<<warps_grid((int64_t)t.B * t.Np), 256, 0, st>>>(static_cast<const bf*>(t.ptr), t.batch_stride, t.row_stride, t.B,
                                                                        t.Np, t.D, rn);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int knn_select(const float* G, int ldg, const float* rn, int B, int Np, int k, int32_t* idx, float* vals, cudaStream_t st) {
  knn_select_kernel<<<warps_grid((int64_t)B * Np), 256, 0, st>>>(G, ldg, rn, B, Np, k, idx, vals);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dense_softmax_fwd(const float* G, int ldg, const float* rn, int B, int Np, int ldA, void* A, cudaStream_t st) {
  dense_softmax_fwd_kernel<<<warps_grid((int64_t)B * Np), 256, 0, st>>>(G, ldg, rn, B, Np, ldA, static_cast<bf*>(A));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dense_softmax_bwd(const float* dA, int ldg, const void* A, int ldA, const float* rn, int B, int Np, void* dG, cudaStream_t st) {
  dense_softmax_bwd_kernel<<<warps_grid((int64_t)B * Np), 256, 0, st>>>(dA, ldg, static_cast<const bf*>(A), ldA, rn, B, Np,
                                                                        static_cast<bf*>(dG));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dense_combine_bwd(const void* T, const void* V, const Tokens& t, const float* rn, void* dp, cudaStream_t st) {
  dense_combine_bwd_kernel<<<warps_grid((int64_t)t.B * t.Np), 256, 0, st>>>(static_cast<const bf*>(T), static_cast<const bf*>(V),
                                                                            static_cast<const bf*>(t.ptr), t.batch_stride, t.row_stride, rn,
                                                                            t.B, t.Np, t.D, static_cast<bf*>(dp));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

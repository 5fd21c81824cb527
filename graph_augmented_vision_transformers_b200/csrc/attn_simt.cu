// attn_simt.cu - exact-fp32 attention core, replaces /root/reference/src/models/vit.py:59-69 on the
// fp32 parity path.  Consumes the packed (B,N,3,H,dh) projection output in place and writes the
// head-major (B,N,H*dh) layout that vit.py:69's transpose+reshape produces, so neither the permute of
// vit.py:60 nor the copy of vit.py:69 exists here.  The (B,H,N,N) score tensor of vit.py:64 is never
// materialised: scores live in registers with an online softmax; the backward recomputes them from the
// saved log-sum-exp.
//
// Shape of the work: one CTA per (32-query tile, head, image), 4 warps x 8 rows; keys/values stream
// through shared memory 32 at a time.  lane <-> key for the q.k and dO.v dot products, lane <-> feature
// for the P.V / dS.K accumulations (probabilities are broadcast with shuffles).
#include <float.h>

#include "kernels.cuh"

namespace gvit {
namespace {

constexpr int TQ = 32, TK = 32, ROWS_PER_WARP = 8;

// feature owned by (lane, t); head dims below 32 leave the upper lanes idle (reads clamped, writes predicated)
template <int DH> __device__ __forceinline__ int feat(int lane, int t) { return min(lane + 32 * t, DH - 1); }
template <int DH> __device__ __forceinline__ bool feat_ok(int lane, int t) { return lane + 32 * t < DH; }

template <typename T>
__device__ __forceinline__ const T* qkv_ptr(const T* qkv, int b, int n, int which, int h, int N, int H, int DH) {
  return qkv + ((((int64_t)b * N + n) * 3 + which) * H + h) * DH;
}

// cooperative load of up to 32 rows x DH into smem[32][LD] as fp32; rows past N are zero-filled
template <typename T, int DH, int LD>
__device__ __forceinline__ void load_rows(float (*dst)[LD], const T* base, int64_t row_stride, int n0, int N) {
  constexpr int CH = DH / 8;
  for (int c = threadIdx.x; c < 32 * CH; c += blockDim.x) {
    const int r = c / CH, d0 = (c % CH) * 8;
    float v[8] = {};
    if (n0 + r < N) load8(base + (int64_t)(n0 + r) * row_stride + d0, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) dst[r][d0 + t] = v[t];
  }
}

template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const T* __restrict__ qkv, int N, int H, float scale,
                                                       T* __restrict__ out, float* __restrict__ lse) {
  constexpr int DD = (DH + 31) / 32;
  __shared__ float Qs[TQ][DH];
  __shared__ float Ks[TK][DH + 1];
  __shared__ float Vs[TK][DH + 1];
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t rs = (int64_t)3 * H * DH;
  load_rows<T, DH, DH>(Qs, qkv_ptr(qkv, b, 0, 0, h, N, H, DH), rs, q0, N);

  float m[ROWS_PER_WARP], l[ROWS_PER_WARP], o[ROWS_PER_WARP][DD];
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    m[r] = -FLT_MAX; l[r] = 0.f;
#pragma unroll
    for (int d = 0; d < DD; ++d) o[r][d] = 0.f;
  }
  for (int k0 = 0; k0 < N; k0 += TK) {
    __syncthreads();
    load_rows<T, DH, DH + 1>(Ks, qkv_ptr(qkv, b, 0, 1, h, N, H, DH), rs, k0, N);
    load_rows<T, DH, DH + 1>(Vs, qkv_ptr(qkv, b, 0, 2, h, N, H, DH), rs, k0, N);
    __syncthreads();
    const bool kvalid = k0 + lane < N;
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
      const float* q = Qs[warp * ROWS_PER_WARP + r];
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], Ks[lane][d], s);
      s = kvalid ? s * scale : -FLT_MAX;
      const float mn = fmaxf(m[r], warp_max(s));
      const float pr = kvalid ? expf(s - mn) : 0.f;
      const float corr = expf(m[r] - mn);
      l[r] = l[r] * corr + warp_sum(pr);
      m[r] = mn;
      float acc[DD];
#pragma unroll
      for (int d = 0; d < DD; ++d) acc[d] = o[r][d] * corr;
      for (int kk = 0; kk < TK; ++kk) {
        const float pk = __shfl_sync(0xffffffffu, pr, kk);
#pragma unroll
        for (int d = 0; d < DD; ++d) acc[d] = fmaf(pk, Vs[kk][feat<DH>(lane, d)], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < DD; ++d) o[r][d] = acc[d];
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    const int q = q0 + warp * ROWS_PER_WARP + r;
    if (q >= N) continue;
    const float inv = 1.0f / l[r];
    T* dst = out + ((int64_t)b * N + q) * H * DH + h * DH;
#pragma unroll
    for (int d = 0; d < DD; ++d)
      if (feat_ok<DH>(lane, d)) dst[lane + 32 * d] = from_f32<T>(o[r][d] * inv);
    if (lane == 0) lse[((int64_t)b * H + h) * N + q] = m[r] + logf(l[r]);
  }
}

// dQ (and delta = rowsum(dO*O), written for the dK/dV pass)
template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                          const T* __restrict__ dout, const float* __restrict__ lse,
                                                          int N, int H, float scale, float* __restrict__ delta,
                                                          T* __restrict__ dqkv) {
  constexpr int DD = (DH + 31) / 32;
  __shared__ float Qs[TQ][DH];
  __shared__ float dOs[TQ][DH];
  __shared__ float Ks[TK][DH + 1];
  __shared__ float Vs[TK][DH + 1];
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t rs = (int64_t)3 * H * DH, os = (int64_t)H * DH;
  load_rows<T, DH, DH>(Qs, qkv_ptr(qkv, b, 0, 0, h, N, H, DH), rs, q0, N);
  load_rows<T, DH, DH>(dOs, dout + (int64_t)b * N * os + h * DH, os, q0, N);
  __syncthreads();

  float lse_r[ROWS_PER_WARP], del_r[ROWS_PER_WARP], dq[ROWS_PER_WARP][DD];
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    const int lr = warp * ROWS_PER_WARP + r, q = q0 + lr;
    float d = 0.f;
    if (q < N) {
      const T* orow = out + ((int64_t)b * N + q) * os + h * DH;
#pragma unroll
      for (int t = 0; t < DD; ++t)
        if (feat_ok<DH>(lane, t)) d = fmaf(to_f32(orow[lane + 32 * t]), dOs[lr][lane + 32 * t], d);
    }
    d = warp_sum(d);
    del_r[r] = d;
    lse_r[r] = q < N ? lse[((int64_t)b * H + h) * N + q] : 0.f;
    if (lane == 0 && q < N) delta[((int64_t)b * H + h) * N + q] = d;
#pragma unroll
    for (int t = 0; t < DD; ++t) dq[r][t] = 0.f;
  }
  for (int k0 = 0; k0 < N; k0 += TK) {
    __syncthreads();
    load_rows<T, DH, DH + 1>(Ks, qkv_ptr(qkv, b, 0, 1, h, N, H, DH), rs, k0, N);
    load_rows<T, DH, DH + 1>(Vs, qkv_ptr(qkv, b, 0, 2, h, N, H, DH), rs, k0, N);
    __syncthreads();
    const bool kvalid = k0 + lane < N;
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
      const int lr = warp * ROWS_PER_WARP + r;
      float s = 0.f, dpv = 0.f;
#pragma unroll 16
      for (int d = 0; d < DH; ++d) {
        s = fmaf(Qs[lr][d], Ks[lane][d], s);
        dpv = fmaf(dOs[lr][d], Vs[lane][d], dpv);
      }
      const float pr = kvalid ? expf(s * scale - lse_r[r]) : 0.f;
      const float ds = pr * (dpv - del_r[r]) * scale;
      for (int kk = 0; kk < TK; ++kk) {
        const float dk = __shfl_sync(0xffffffffu, ds, kk);
#pragma unroll
        for (int t = 0; t < DD; ++t) dq[r][t] = fmaf(dk, Ks[kk][feat<DH>(lane, t)], dq[r][t]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    const int q = q0 + warp * ROWS_PER_WARP + r;
    if (q >= N) continue;
    T* dst = const_cast<T*>(qkv_ptr(dqkv, b, q, 0, h, N, H, DH));
#pragma unroll
    for (int t = 0; t < DD; ++t)
      if (feat_ok<DH>(lane, t)) dst[lane + 32 * t] = from_f32<T>(dq[r][t]);
  }
}

// dK, dV: one CTA per 32-key tile; queries stream through shared memory
template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                           const float* __restrict__ lse,
                                                           const float* __restrict__ delta, int N, int H, float scale,
                                                           T* __restrict__ dqkv) {
  constexpr int DD = (DH + 31) / 32;
  __shared__ float Ks[TK][DH];
  __shared__ float Vs[TK][DH];
  __shared__ float Qs[TQ][DH + 1];
  __shared__ float dOs[TQ][DH + 1];
  __shared__ float lse_s[TQ], del_s[TQ];
  const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t rs = (int64_t)3 * H * DH, os = (int64_t)H * DH;
  load_rows<T, DH, DH>(Ks, qkv_ptr(qkv, b, 0, 1, h, N, H, DH), rs, k0, N);
  load_rows<T, DH, DH>(Vs, qkv_ptr(qkv, b, 0, 2, h, N, H, DH), rs, k0, N);

  float dk[ROWS_PER_WARP][DD], dv[ROWS_PER_WARP][DD];
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r)
#pragma unroll
    for (int t = 0; t < DD; ++t) dk[r][t] = dv[r][t] = 0.f;

  for (int q0 = 0; q0 < N; q0 += TQ) {
    __syncthreads();
    load_rows<T, DH, DH + 1>(Qs, qkv_ptr(qkv, b, 0, 0, h, N, H, DH), rs, q0, N);
    load_rows<T, DH, DH + 1>(dOs, dout + (int64_t)b * N * os + h * DH, os, q0, N);
    if (threadIdx.x < TQ) {
      const int q = q0 + threadIdx.x;
      lse_s[threadIdx.x] = q < N ? lse[((int64_t)b * H + h) * N + q] : 0.f;
      del_s[threadIdx.x] = q < N ? delta[((int64_t)b * H + h) * N + q] : 0.f;
    }
    __syncthreads();
    const bool qvalid = q0 + lane < N;
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
      const int lk = warp * ROWS_PER_WARP + r;
      float s = 0.f, dpv = 0.f;
#pragma unroll 16
      for (int d = 0; d < DH; ++d) {
        s = fmaf(Qs[lane][d], Ks[lk][d], s);
        dpv = fmaf(dOs[lane][d], Vs[lk][d], dpv);
      }
      const float pr = qvalid ? expf(s * scale - lse_s[lane]) : 0.f;
      const float ds = pr * (dpv - del_s[lane]) * scale;
      for (int qq = 0; qq < TQ; ++qq) {
        const float pq = __shfl_sync(0xffffffffu, pr, qq), dq = __shfl_sync(0xffffffffu, ds, qq);
#pragma unroll
        for (int t = 0; t < DD; ++t) {
          dv[r][t] = fmaf(pq, dOs[qq][feat<DH>(lane, t)], dv[r][t]);
          dk[r][t] = fmaf(dq, Qs[qq][feat<DH>(lane, t)], dk[r][t]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    const int kr = k0 + warp * ROWS_PER_WARP + r;
    if (kr >= N) continue;
    T* dkp = const_cast<T*>(qkv_ptr(dqkv, b, kr, 1, h, N, H, DH));
    T* dvp = const_cast<T*>(qkv_ptr(dqkv, b, kr, 2, h, N, H, DH));
#pragma unroll
    for (int t = 0; t < DD; ++t) {
      if (!feat_ok<DH>(lane, t)) continue;
      dkp[lane + 32 * t] = from_f32<T>(dk[r][t]);
      dvp[lane + 32 * t] = from_f32<T>(dv[r][t]);
    }
  }
}

template <typename T, int DH>
int launch_fwd(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, cudaStream_t st) {
  dim3 grid((N + TQ - 1) / TQ, H, B);
  attn_fwd_kernel<T, DH><<<grid, 128, 0, st>>>(static_cast<const T*>(qkv), N, H, scale, static_cast<T*>(out), lse);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

template <typename T, int DH>
int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, float scale,
               float* delta, void* dqkv, cudaStream_t st) {
  dim3 grid((N + TQ - 1) / TQ, H, B);
  attn_bwd_dq_kernel<T, DH><<<grid, 128, 0, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(out),
                                                  static_cast<const T*>(dout), lse, N, H, scale, delta,
                                                  static_cast<T*>(dqkv));
  GVIT_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<T, DH><<<grid, 128, 0, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(dout), lse, delta,
                                                   N, H, scale, static_cast<T*>(dqkv));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

int attn_fwd_simt(const void* qkv, int B, int N, int H, int dh, float scale, int dtype, void* out, float* lse,
                  cudaStream_t st) {
  using bf = __nv_bfloat16;
  if (dh == 64) return dtype == GVIT_F32 ? launch_fwd<float, 64>(qkv, B, N, H, scale, out, lse, st)
                                         : launch_fwd<bf, 64>(qkv, B, N, H, scale, out, lse, st);
  if (dh == 32) return dtype == GVIT_F32 ? launch_fwd<float, 32>(qkv, B, N, H, scale, out, lse, st)
                                         : launch_fwd<bf, 32>(qkv, B, N, H, scale, out, lse, st);
  if (dh == 16) return dtype == GVIT_F32 ? launch_fwd<float, 16>(qkv, B, N, H, scale, out, lse, st)
                                         : launch_fwd<bf, 16>(qkv, B, N, H, scale, out, lse, st);
  return fail(GVIT_ERR_UNSUPPORTED, "attn_fwd: head dim %d (supported: 16, 32, 64)", dh);
}

int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, int dh,
                  float scale, int dtype, float* delta_ws, void* dqkv, cudaStream_t st) {
  using bf = __nv_bfloat16;
  if (dh == 64) return dtype == GVIT_F32 ? launch_bwd<float, 64>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st)
                                         : launch_bwd<bf, 64>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st);
  if (dh == 32) return dtype == GVIT_F32 ? launch_bwd<float, 32>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st)
                                         : launch_bwd<bf, 32>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st);
  if (dh == 16) return dtype == GVIT_F32 ? launch_bwd<float, 16>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st)
                                         : launch_bwd<bf, 16>(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st);
  return fail(GVIT_ERR_UNSUPPORTED, "attn_bwd: head dim %d (supported: 16, 32, 64)", dh);
}

}  // namespace gvit

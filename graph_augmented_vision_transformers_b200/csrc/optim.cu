// optim.cu - the optimiser side of the reference's training step (/root/reference/src/training/trainer.py) as three
// multi-tensor launches over ALL parameters of both AdamW groups:
//
//   trainer.py:114-116  clip_grad_norm_(params, 1.0)     -> mt_sqnorm (per-chunk partial sums of g^2, fixed order)
//   trainer.py:77-87    LambdaLR warm-up + cosine, /step  -> mt_step_prologue (ONE thread: total norm, clip coefficient,
//                                                            schedule factor of this step, step counter += 1; all on the device)
//   trainer.py:47-56,118 AdamW (2 groups), optimizer.step -> mt_adamw (clip coefficient and schedule factor folded in: the
//                                                            gradients are read once and never rewritten)
//
// torch's path is clip (norm kernels + one pass that rescales every gradient in place) + fused AdamW + a host-side scheduler
// that rewrites param_group['lr']; here nothing touches the host, so the whole step stays inside one CUDA graph, and the
// clip costs no extra pass over the 372 MB of ViT-B gradients.
// Tensors are addressed through device tables (addresses, sizes, per-tensor lr / weight decay) owned by the caller; a
// "chunk" is CHUNK consecutive elements of one tensor, one CTA per chunk.
#include "kernels.cuh"

namespace gvit {
namespace {

constexpr int CHUNK = 32768;

struct MtTables {
  const int64_t* p;        // parameter addresses (fp32)
  const int64_t* g;        // gradient addresses (fp32), 0 = no gradient this step (tensor skipped, as torch does)
  const int64_t* m;        // exp_avg
  const int64_t* v;        // exp_avg_sq
  const int64_t* numel;
  const float* lr;         // per-tensor base learning rate (its param group's)
  const float* wd;         // per-tensor weight decay
  const int32_t* chunk_tensor;
  const int32_t* chunk_index;   // chunk number inside its tensor
  int32_t* tstep;          // per-tensor AdamW step count (torch keeps `step` per parameter: a skipped tensor does not advance)
  int ntensors;
};

__global__ void __launch_bounds__(256) mt_sqnorm_kernel(MtTables T, float* __restrict__ partial) {
  const int c = blockIdx.x;
  const int t = T.chunk_tensor[c];
  const float* g = reinterpret_cast<const float*>(T.g[t]);
  float acc = 0.f;
  if (g != nullptr) {
    const int64_t n = T.numel[t], lo = (int64_t)T.chunk_index[c] * CHUNK;
    const int64_t hi = lo + CHUNK < n ? lo + CHUNK : n;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      int64_t i = lo + threadIdx.x * 4;
      for (; i + 3 < hi; i += 1024) {
        const float4 x = *reinterpret_cast<const float4*>(g + i);
        acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc); acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
      }
      for (; i < hi; ++i) acc = fmaf(g[i], g[i], acc);        // at most 3 tail elements, one thread
    } else {
      for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc = fmaf(g[i], g[i], acc);
    }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[c] = s;
  }
}

// sched[0] = total gradient norm, sched[1] = clip coefficient, sched[2] = schedule factor lambda(step); `step` counts the
// completed optimiser steps (the scheduler's epoch), tstep[i] the updates tensor i has received (AdamW's bias correction)
__global__ void __launch_bounds__(256) mt_step_prologue_kernel(MtTables T, const float* __restrict__ partial, int nchunks, float max_norm,
                                                               int64_t warmup_steps, int64_t total_steps,
                                                               int64_t* __restrict__ step, float* __restrict__ sched) {
  __shared__ double red[256];
  for (int i = threadIdx.x; i < T.ntensors; i += 256)
    if (T.g[i] != 0) T.tstep[i] += 1;                                              // the 1-based count this step's update uses
  double acc = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += 256) acc += (double)partial[i];   // fixed assignment, fixed tree: deterministic
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    sched[0] = norm;
    float coef = 1.0f;
    if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);            // torch.nn.utils.clip_grad_norm_
    sched[1] = coef;
    const int64_t s0 = *step;                                                       // scheduler epoch: completed steps
    double lam = 1.0;
    if (total_steps > 0) {                                                          // trainer.py:81-85
      if (s0 < warmup_steps) lam = (double)s0 / (double)(warmup_steps > 1 ? warmup_steps : 1);
      else {
        const int64_t den = total_steps - warmup_steps > 1 ? total_steps - warmup_steps : 1;
        lam = 0.5 * (1.0 + cos(3.14159265358979323846 * (double)(s0 - warmup_steps) / (double)den));
      }
    }
    sched[2] = (float)lam;
    *step = s0 + 1;
  }
}

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, float lr, float wd, float beta1, float beta2, float eps,
                                       float step_size, float bc2_sqrt) {
  // torch.optim.AdamW (single-tensor formula order): decoupled decay, moment updates, bias-corrected step
  p *= 1.0f - lr * wd;
  m = m + (g - m) * (1.0f - beta1);                                                  // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.0f - beta2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) mt_adamw_kernel(MtTables T, const float* __restrict__ sched, float beta1, float beta2, float eps) {
  const int c = blockIdx.x;
  const int t = T.chunk_tensor[c];
  const float* g = reinterpret_cast<const float*>(T.g[t]);
  if (g == nullptr) return;
  float* p = reinterpret_cast<float*>(T.p[t]);
  float* m = reinterpret_cast<float*>(T.m[t]);
  float* v = reinterpret_cast<float*>(T.v[t]);
  __shared__ float bc[2];
  if (threadIdx.x == 0) {
    const double ts = (double)T.tstep[t];
    bc[0] = (float)(1.0 - pow((double)beta1, ts));
    bc[1] = (float)sqrt(1.0 - pow((double)beta2, ts));
  }
  __syncthreads();
  const float coef = sched[1], lr = T.lr[t] * sched[2], wd = T.wd[t];
  const float step_size = lr / bc[0], bc2_sqrt = bc[1];
  const int64_t n = T.numel[t], lo = (int64_t)T.chunk_index[c] * CHUNK;
  const int64_t hi = lo + CHUNK < n ? lo + CHUNK : n;
  const bool al = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  if (al) {
    int64_t i = lo + threadIdx.x * 4;
    for (; i + 3 < hi; i += 1024) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      adamw1(pp.x, gg.x * coef, mm.x, vv.x, lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
      adamw1(pp.y, gg.y * coef, mm.y, vv.y, lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
      adamw1(pp.z, gg.z * coef, mm.z, vv.z, lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
      adamw1(pp.w, gg.w * coef, mm.w, vv.w, lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
    for (; i < hi; ++i) adamw1(p[i], g[i] * coef, m[i], v[i], lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) adamw1(p[i], g[i] * coef, m[i], v[i], lr, wd, beta1, beta2, eps, step_size, bc2_sqrt);
  }
}

}  // namespace

int mt_chunk_elems() { return CHUNK; }

int mt_adamw_step(const int64_t* p, const int64_t* g, const int64_t* m, const int64_t* v, const int64_t* numel, const float* lr,
                  const float* wd, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t* tstep, int ntensors, int nchunks,
                  float max_norm, int64_t warmup_steps, int64_t total_steps, float beta1, float beta2, float eps, int64_t* step,
                  float* sched, float* partial_ws, cudaStream_t st) {
  MtTables T{p, g, m, v, numel, lr, wd, chunk_tensor, chunk_index, tstep, ntensors};
  mt_sqnorm_kernel<<<nchunks, 256, 0, st>>>(T, partial_ws);
  GVIT_CHECK_LAUNCH();
  mt_step_prologue_kernel<<<1, 256, 0, st>>>(T, partial_ws, nchunks, max_norm, warmup_steps, total_steps, step, sched);
  GVIT_CHECK_LAUNCH();
  mt_adamw_kernel<<<nchunks, 256, 0, st>>>(T, sched, beta1, beta2, eps);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

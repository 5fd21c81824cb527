// edges.cu - the bandwidth-bound edges of the block: LayerNorm (nn.LayerNorm at
// /root/reference/src/models/vit.py:103,108,154) and proj_drop + residual add (vit.py:71,117).
// One warp per token row, 16-byte loads/stores, fp32 statistics, no shared-memory staging
// (each element is touched once).
#include "gelu.cuh"
#include "kernels.cuh"
#include "philox.cuh"

namespace gvit {
namespace {

// A lane owns NC chunks of 8 consecutive features: chunk c covers [c*256 + lane*8, +8).  NC = ceil(D / 256) is a
// template parameter so the per-thread arrays are exactly as large as the row needs (D <= 1024 -> NC <= 4).
// Tx is the dtype of x / gamma / beta / dx (the residual stream), Ty the dtype of y / dy (the branch input):
// under autocast the stream is fp32 and the branch bf16, and the cast is folded into this kernel.
template <typename Tx, typename Ty, int NC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const Tx* __restrict__ x, const Tx* __restrict__ gamma,
                                                     const Tx* __restrict__ beta, int64_t rows, int D, float eps,
                                                     Ty* __restrict__ y, float* __restrict__ mean,
                                                     float* __restrict__ rstd) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[NC][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
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
  for (int c = 0; c < NC; ++c) {
    const int d0 = lane * 8 + c * 256;
    if (d0 < D) {
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float dlt = v[c][t] - mu; q = fmaf(dlt, dlt, q); }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
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

// ---- per-warp bulk-copy row pipeline (cp.async.bulk + mbarrier) for the LayerNorm backward ----------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// dx for every row; per-CTA partial dgamma/dbeta (fixed row -> CTA assignment: deterministic, no atomics).
// Each warp owns a private ring of `slots` shared-memory row slots filled by cp.async.bulk (x row + dy row per slot,
// one mbarrier per slot): the rows of the next `slots` iterations are always in flight without costing registers.
// v1 (rows loaded straight into registers, one row in flight per warp, 16 warps/SM) reached 46 % of HBM.
template <typename Tx, typename Ty, int NC, bool HAS_ADD>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const Ty* __restrict__ dy, const Tx* __restrict__ x,
                                                        const Tx* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, int64_t rows, int D, int slots,
                                                        const Tx* __restrict__ dx_add_, Tx* __restrict__ dx,
                                                        float* __restrict__ partial) {
  const Tx* const dx_add = HAS_ADD ? dx_add_ : nullptr;     // compile-time: two clean instantiations
  extern __shared__ __align__(128) uint8_t ring[];
  float (*red)[256] = reinterpret_cast<float (*)[256]>(ring);   // the ring is dead when the final reduction runs
  __shared__ float gsm[NC * 256];
  __shared__ __align__(8) uint64_t bars[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t xbytes = D * sizeof(Tx), ybytes = D * sizeof(Ty);
  const uint32_t xpad = (xbytes + 127u) & ~127u, ypad = (ybytes + 127u) & ~127u;
  const uint32_t slot_bytes = xpad + ypad + (dx_add ? xpad : 0u);     // optional third stream: gradient to add to dx
  uint8_t* wring = ring + (size_t)warp * slots * slot_bytes;
  for (int i = threadIdx.x; i < NC * 256; i += 256) gsm[i] = i < D ? to_f32(gamma[i]) : 0.f;
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) bar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * 8;
  const int64_t row0 = (int64_t)blockIdx.x * 8 + warp;
  auto issue = [&](int64_t r, int s) {      // lane 0 only
    bar_expect_tx(&bars[warp][s], xbytes + ybytes + (dx_add ? xbytes : 0u));
    bulk_copy_g2s(wring + (size_t)s * slot_bytes, x + r * D, xbytes, &bars[warp][s]);
    bulk_copy_g2s(wring + (size_t)s * slot_bytes + xpad, dy + r * D, ybytes, &bars[warp][s]);
    if (dx_add) bulk_copy_g2s(wring + (size_t)s * slot_bytes + xpad + ypad, dx_add + r * D, xbytes, &bars[warp][s]);
  };
  if (lane == 0)
    for (int s = 0; s < slots; ++s)
      if (row0 + s * stride < rows) issue(row0 + s * stride, s);
  float dg[NC][8] = {}, db[NC][8] = {};
  const float invD = 1.0f / D;
  float mu_n = 0.f, rs_n = 0.f;
  if (row0 < rows) { mu_n = mean[row0]; rs_n = rstd[row0]; }
  int it = 0;
  for (int64_t row = row0; row < rows; row += stride, ++it) {
    const int s = it % slots;
    const float mu = mu_n, rs = rs_n;
    if (row + stride < rows) { mu_n = mean[row + stride]; rs_n = rstd[row + stride]; }
    bar_wait(&bars[warp][s], (it / slots) & 1);
    const Tx* xs = reinterpret_cast<const Tx*>(wring + (size_t)s * slot_bytes);
    const Ty* ys = reinterpret_cast<const Ty*>(wring + (size_t)s * slot_bytes + xpad);
    const Tx* as = reinterpret_cast<const Tx*>(wring + (size_t)s * slot_bytes + xpad + ypad);
    // Two passes over the row IN THE SLOT (shared memory), nothing but the 2 x NC x 8 column accumulators lives across
    // them: holding the normalised row and gy in registers through the two warp reductions put the kernel at 128
    // registers with local-memory spills on its hottest instructions (10 % of all stall samples on one STL).
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float xv[8], dyv[8];
        load8(xs + d0, xv);
        load8(ys + d0, dyv);
        const float4 g0 = *reinterpret_cast<const float4*>(&gsm[d0]), g1 = *reinterpret_cast<const float4*>(&gsm[d0 + 4]);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float xh = (xv[t] - mu) * rs;
          dg[c][t] = fmaf(dyv[t], xh, dg[c][t]);
          db[c][t] += dyv[t];
          const float gy = dyv[t] * g[t];
          s1 += gy;
          s2 = fmaf(gy, xh, s2);
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    asm volatile("" ::: "memory");                           // pass 2 re-reads the slot instead of keeping pass 1's values
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int d0 = lane * 8 + c * 256;
      if (d0 < D) {
        float xv[8], dyv[8], o[8];
        load8(xs + d0, xv);
        load8(ys + d0, dyv);
        const float4 g0 = *reinterpret_cast<const float4*>(&gsm[d0]), g1 = *reinterpret_cast<const float4*>(&gsm[d0 + 4]);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = rs * (dyv[t] * g[t] - s1 - ((xv[t] - mu) * rs) * s2);
        if (dx_add) {
          float a[8];
          load8(as + d0, a);
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] += a[t];
        }
        store8(dx + row * D + d0, o);
      }
    }
    __syncwarp();                                            // every lane is done with the slot: refill it
    if (lane == 0 && row + (int64_t)slots * stride < rows) issue(row + (int64_t)slots * stride, s);
  }
  __syncthreads();                                          // all warps out of the ring before it is reused below
  // reduce the 8 warps' partials column by column through shared memory, 256 columns at a time
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 8; ++t) red[warp][lane * 8 + t] = pass == 0 ? dg[c][t] : db[c][t];
      __syncthreads();
      const int col = c * 256 + threadIdx.x;
      if (col < D) {
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += red[w][threadIdx.x];
        partial[((int64_t)pass * gridDim.x + blockIdx.x) * D + col] = sacc;
      }
    }
  }
}

__global__ void ln_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int D, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  // one warp per (column, which): lanes stride over the CTA partials, then a shuffle reduction (fixed order)
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= 2 * D) return;
  const int which = wid / D, col = wid % D;
  const float* src = partial + (int64_t)which * nblk * D + col;
  float a = 0.f;
  for (int i = lane; i < nblk; i += 32) a += src[(int64_t)i * D];
  a = warp_sum(a);
  if (lane == 0) (which == 0 ? dgamma : dbeta)[col] = a;
}

// out = resid + dropout(y): y has the branch dtype Ty, resid / out the stream dtype Tx (Tx == Ty when resid is null).
// The keep mask is stored as one BIT per element (byte i covers elements 8i..8i+7).
template <typename Tx, typename Ty>
__global__ void __launch_bounds__(256) dropout_residual_fwd_kernel(const Ty* __restrict__ y, const Tx* __restrict__ resid,
                                                                   int64_t n, float p, uint64_t seed, uint64_t offset,
                                                                   const uint64_t* __restrict__ offset_dev,
                                                                   Tx* __restrict__ out, uint8_t* __restrict__ mask) {
  if (offset_dev) offset += __ldg(offset_dev);
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t th = dropout_thresh16(p);
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += (int64_t)gridDim.x * blockDim.x * 8) {
    float a[8], r[8] = {};
    load8(y + i8, a);
    if (resid) load8(resid + i8, r);
    if (p > 0.f) {
      const uint32_t bits = keep_bits8(seed, offset + (uint64_t)(i8 >> 3), th);
#pragma unroll
      for (int t = 0; t < 8; ++t) a[t] = (bits >> t) & 1u ? a[t] * scale : 0.f;
      mask[i8 >> 3] = (uint8_t)bits;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] += r[t];
    store8(out + i8, a);
  }
}

template <typename Tx, typename Ty>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const Tx* __restrict__ dout, const uint8_t* __restrict__ mask,
                                                          int64_t n, float p, Ty* __restrict__ dy) {
  const float scale = 1.0f / (1.0f - p);
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += (int64_t)gridDim.x * blockDim.x * 8) {
    float a[8];
    load8(dout + i8, a);
    const uint32_t bits = mask ? mask[i8 >> 3] : 0xffu;          // no mask (p == 0): a pure cast of the stream gradient
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = (bits >> t) & 1u ? a[t] * scale : 0.f;
    store8(dy + i8, a);
  }
}

// Two 8-element groups per thread and iteration (both loads issued before any math): the kernels are ALU-heavy
// (GELU + Philox), so one 16-byte load in flight per thread left the memory pipe idle while the math ran.
template <typename T>
__global__ void __launch_bounds__(256) gelu_dropout_fwd_kernel(const T* __restrict__ u, int64_t n, float p, uint64_t seed,
                                                               uint64_t offset, const uint64_t* __restrict__ offset_dev,
                                                               T* __restrict__ out, uint8_t* __restrict__ mask) {
  if (offset_dev) offset += __ldg(offset_dev);
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t th = dropout_thresh16(p);
  const int64_t step = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += 2 * step) {
    const int64_t j8 = i8 + step;
    const bool two = j8 < n;
    float a[8], b[8];
    load8(u + i8, a);
    if (two) load8(u + j8, b);
    GeluFor<T>::type::fwd8(a, scale);                        // the dropout scale rides in the GELU constants
    if (p > 0.f) {
      const uint32_t bits = keep_bits8(seed, offset + (uint64_t)(i8 >> 3), th);
#pragma unroll
      for (int t = 0; t < 8; ++t) a[t] = (bits >> t) & 1u ? a[t] : 0.f;
      mask[i8 >> 3] = (uint8_t)bits;
    }
    store8(out + i8, a);
    if (two) {
      GeluFor<T>::type::fwd8(b, scale);
      if (p > 0.f) {
        const uint32_t bits = keep_bits8(seed, offset + (uint64_t)(j8 >> 3), th);
#pragma unroll
        for (int t = 0; t < 8; ++t) b[t] = (bits >> t) & 1u ? b[t] : 0.f;
        mask[j8 >> 3] = (uint8_t)bits;
      }
      store8(out + j8, b);
    }
  }
}

// du = dout * keep/(1-p) * gelu'(u); the activation is recomputed from the saved pre-activation
template <typename T>
__global__ void __launch_bounds__(256) gelu_dropout_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ u,
                                                               const uint8_t* __restrict__ mask, int64_t n, float p,
                                                               T* __restrict__ du) {
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i8 < n; i8 += 2 * step) {
    const int64_t j8 = i8 + step;
    const bool two = j8 < n;
    float g0[8], a0[8], g1[8], a1[8];
    load8(dout + i8, g0);
    load8(u + i8, a0);
    uint32_t bits0 = p > 0.f ? mask[i8 >> 3] : 0xffu, bits1 = 0xffu;
    if (two) {
      load8(dout + j8, g1);
      load8(u + j8, a1);
      if (p > 0.f) bits1 = mask[j8 >> 3];
    }
    GeluFor<T>::type::grad8(g0, a0, scale);
#pragma unroll
    for (int t = 0; t < 8; ++t) g0[t] = (bits0 >> t) & 1u ? g0[t] : 0.f;
    store8(du + i8, g0);
    if (two) {
      GeluFor<T>::type::grad8(g1, a1, scale);
#pragma unroll
      for (int t = 0; t < 8; ++t) g1[t] = (bits1 >> t) & 1u ? g1[t] : 0.f;
      store8(du + j8, g1);
    }
  }
}

// ---- column sums: the bias gradient of nn.Linear (vit.py:50,52,83,85), db[c] = sum_r dy[r, c] ------------------------
// grid (column blocks of 256, row chunks); block (32 column groups of 8, 8 row lanes): a warp reads 512 contiguous bytes
// of one row, four rows in flight per thread; per-chunk partials, then a fixed-order reduction over the chunks
// (deterministic, no atomics).  at::sum over dim 0 reached ~3.2 TB/s on these shapes.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int D, int rows_per_chunk,
                                                             int skip_period, float* __restrict__ partial) {
  __shared__ float red[8][256 + 8];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float acc[8] = {};
  if (col < D) {
    int64_t r = r0 + ry;
    for (; r + 56 < r1; r += 64) {                          // eight rows in flight per thread
      float a[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) load8(x + (r + 8 * i) * D + col, a[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool on = !(skip_period && (r + 8 * i) % skip_period == 0);   // e.g. the CLS rows of a (B, 1+Np, D) tensor
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += on ? a[i][t] : 0.f;
      }
    }
    for (; r < r1; r += 8) {
      float a[8];
      load8(x + r * D + col, a);
      const bool on = !(skip_period && r % skip_period == 0);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] += on ? a[t] : 0.f;
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[ry][cx * 8 + t] = acc[t];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < D) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    partial[(int64_t)blockIdx.y * D + c] = sum;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nchunks, int D,
                                                           float* __restrict__ out) {
  // 32 columns x 8 chunk lanes per CTA; fixed summation order (lane-strided, then lanes 0..7)
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < D)
    for (int i = ly; i < nchunks; i += 8) s += partial[(int64_t)i * D + c];
  red[ly][cx] = s;
  __syncthreads();
  if (ly == 0 && c < D) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][cx];
    out[c] = t;
  }
}

// ---- backward edges fused with the bias gradient of the Linear that produced their input -----------------------------
// dy = dropout_bwd(dout) [* gelu'(u)] is exactly the tensor whose column sums are that Linear's bias gradient
// (vit.py:70-71, 90-93), so the same pass accumulates them: same 2-D mapping as colsum_partial_kernel
// (block = 32 column groups x 8 row lanes, grid = column blocks x row chunks), four rows in flight per thread.
template <typename Tx, typename Ty, bool GELU>
__global__ void __launch_bounds__(256, GELU ? 4 : 2) edge_bwd_colsum_kernel(const Tx* __restrict__ dout, const Ty* __restrict__ u,
                                                              const uint8_t* __restrict__ mask, int64_t rows, int D,
                                                              int rows_per_chunk, float p, int skip_period, Ty* __restrict__ dy,
                                                              float* __restrict__ partial) {
  __shared__ float red[8][256 + 8];
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float acc[8] = {};
  if (col < D) {
    // rows in flight per thread.  GELU': ONE, at 32 warps / SM (<= 64 registers) - measured 184 us against 224 us for
    // two rows at 24 warps and 267 us for three (A/B on B200, B = 256): the GELU' arithmetic needs warps to hide behind
    // more than it needs bytes in flight per warp.
    constexpr int RF = GELU ? 1 : 4;
    for (int64_t r = r0 + ry; r < r1; r += 8 * RF) {
      float g[RF][8], a[RF][8];
      uint32_t bits[RF];
#pragma unroll
      for (int i = 0; i < RF; ++i) {
        const int64_t rr = r + 8 * i;
        bits[i] = 0xffu;
        if (rr < r1) {
          const int64_t o = rr * D + col;
          load8(dout + o, g[i]);
          if (GELU) load8(u + o, a[i]);
          if (p > 0.f) bits[i] = mask[o >> 3];
        }
      }
#pragma unroll
      for (int i = 0; i < RF; ++i) {
        const int64_t rr = r + 8 * i;
        if (rr < r1) {
          const bool on = !(skip_period && rr % skip_period == 0);   // rows left out of the column sums (still written to dy)
          if (GELU) {
            GeluFor<Ty>::type::grad8(g[i], a[i], scale);
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) g[i][t] *= scale;
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            float v = (bits[i] >> t) & 1u ? g[i][t] : 0.f;
            v = to_f32(from_f32<Ty>(v));                   // the bias gradient sums the values as stored
            g[i][t] = v;
            acc[t] += on ? v : 0.f;
          }
          store8(dy + rr * D + col, g[i]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[ry][cx * 8 + t] = acc[t];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < D) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    partial[(int64_t)blockIdx.y * D + cc] = sum;
  }
}

inline void colsum_grid(int64_t rows, int D, int per_sm, int* colblocks, int* nchunks, int* rpc) {
  *colblocks = (D + 255) / 256;
  int n = (per_sm * num_sms() + *colblocks - 1) / *colblocks;
  if (n > GVIT_COLSUM_CHUNKS) n = GVIT_COLSUM_CHUNKS;
  if ((int64_t)n * 32 > rows) n = (int)((rows + 31) / 32);
  *rpc = (int)((rows + n - 1) / n);
  *nchunks = (int)((rows + *rpc - 1) / *rpc);
}



template <typename Tx, typename Ty, int NC>
int ln_fwd_launch(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, void* y, float* mean,
                  float* rstd, cudaStream_t st) {
  ln_fwd_kernel<Tx, Ty, NC><<<(int)((rows + 7) / 8), 256, 0, st>>>(static_cast<const Tx*>(x), static_cast<const Tx*>(gamma),
                                                                  static_cast<const Tx*>(beta), rows, D, eps,
                                                                  static_cast<Ty*>(y), mean, rstd);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}
template <typename Tx, typename Ty, int NC>
int ln_bwd_launch(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, int64_t rows,
                  int D, const void* dx_add, void* dx, float* partial, int nblk, cudaStream_t st) {
  // ring: 8 warps x slots x (x row + dy row [+ add row]); aim for two CTAs per SM (<= ~100 KB of dynamic smem each)
  const size_t xpad = ((size_t)D * sizeof(Tx) + 127) & ~size_t(127);
  const size_t slot_bytes = xpad + (((size_t)D * sizeof(Ty) + 127) & ~size_t(127)) + (dx_add ? xpad : 0);
  int slots = (int)((109 * 1024) / (8 * slot_bytes));
  slots = slots > 4 ? 4 : (slots < 2 ? 2 : slots);
  size_t smem = 8 * slots * slot_bytes;
  if (smem < 8 * 256 * sizeof(float)) smem = 8 * 256 * sizeof(float);   // also hosts the final 8 x 256 reduction buffer
  auto kern = dx_add ? ln_bwd_kernel<Tx, Ty, NC, true> : ln_bwd_kernel<Tx, Ty, NC, false>;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nblk, 256, smem, st>>>(static_cast<const Ty*>(dy), static_cast<const Tx*>(x), static_cast<const Tx*>(gamma), mean,
                                rstd, rows, D, slots, static_cast<const Tx*>(dx_add), static_cast<Tx*>(dx), partial);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

#define GVIT_LN_DISPATCH(FN, ...)                                                        \
  do {                                                                                   \
    const int nc = (D + 255) / 256;                                                      \
    if (dtype == GVIT_F32 && y_dtype == GVIT_F32) {                                      \
      using Tx = float; using Ty = float;                                                \
      if (nc == 1) return FN<Tx, Ty, 1>(__VA_ARGS__);                                    \
      if (nc == 2) return FN<Tx, Ty, 2>(__VA_ARGS__);                                    \
      if (nc == 3) return FN<Tx, Ty, 3>(__VA_ARGS__);                                    \
      return FN<Tx, Ty, 4>(__VA_ARGS__);                                                 \
    } else if (dtype == GVIT_F32) {                                                      \
      using Tx = float; using Ty = __nv_bfloat16;                                        \
      if (nc == 1) return FN<Tx, Ty, 1>(__VA_ARGS__);                                    \
      if (nc == 2) return FN<Tx, Ty, 2>(__VA_ARGS__);                                    \
      if (nc == 3) return FN<Tx, Ty, 3>(__VA_ARGS__);                                    \
      return FN<Tx, Ty, 4>(__VA_ARGS__);                                                 \
    } else {                                                                             \
      using Tx = __nv_bfloat16; using Ty = __nv_bfloat16;                                \
      if (nc == 1) return FN<Tx, Ty, 1>(__VA_ARGS__);                                    \
      if (nc == 2) return FN<Tx, Ty, 2>(__VA_ARGS__);                                    \
      if (nc == 3) return FN<Tx, Ty, 3>(__VA_ARGS__);                                    \
      return FN<Tx, Ty, 4>(__VA_ARGS__);                                                 \
    }                                                                                    \
  } while (0)

}  // namespace

int layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, int dtype,
                  int y_dtype, void* y, float* mean, float* rstd, cudaStream_t st) {
  GVIT_LN_DISPATCH(ln_fwd_launch, x, gamma, beta, rows, D, eps, y, mean, rstd, st);
}

static int ln_bwd_main(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                       int64_t rows, int D, int dtype, int y_dtype, const void* dx_add, void* dx, float* partial_ws,
                       int nblk, cudaStream_t st) {
  GVIT_LN_DISPATCH(ln_bwd_launch, dy, x, gamma, mean, rstd, rows, D, dx_add, dx, partial_ws, nblk, st);
}

int layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, int64_t rows,
                  int D, int dtype, int y_dtype, const void* dx_add, void* dx, float* dgamma, float* dbeta,
                  float* partial_ws, cudaStream_t st) {
  int nblk = (int)((rows + 7) / 8);
  if (nblk > GVIT_LN_PARTIALS) nblk = GVIT_LN_PARTIALS;
  int rc = ln_bwd_main(dy, x, gamma, mean, rstd, rows, D, dtype, y_dtype, dx_add, dx, partial_ws, nblk, st);
  if (rc != GVIT_OK) return rc;
  ln_bwd_reduce_kernel<<<(2 * D * 32 + 255) / 256, 256, 0, st>>>(partial_ws, nblk, D, dgamma, dbeta);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int colsum(const void* x, int64_t rows, int D, int dtype, int skip_period, float* out, float* partial_ws, cudaStream_t st) {
  const int colblocks = (D + 255) / 256;
  int nchunks = (6 * num_sms() + colblocks - 1) / colblocks;
  if (nchunks > GVIT_COLSUM_CHUNKS) nchunks = GVIT_COLSUM_CHUNKS;
  if ((int64_t)nchunks * 32 > rows) nchunks = (int)((rows + 31) / 32);
  const int rpc = (int)((rows + nchunks - 1) / nchunks);
  nchunks = (int)((rows + rpc - 1) / rpc);
  dim3 grid(colblocks, nchunks);
  if (dtype == GVIT_F32)
    colsum_partial_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), rows, D, rpc, skip_period, partial_ws);
  else
    colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), rows, D, rpc, skip_period, partial_ws);
  GVIT_CHECK_LAUNCH();
  colsum_final_kernel<<<(D + 31) / 32, 256, 0, st>>>(partial_ws, nchunks, D, out);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dropout_residual_fwd(const void* y, const void* resid, int64_t n, float p, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, int dtype, int y_dtype, void* out, uint8_t* keep_mask, cudaStream_t st) {
  const int64_t n8 = n / 8;
  using bf = __nv_bfloat16;
  if (dtype == GVIT_F32 && y_dtype == GVIT_F32)
    stream_launch(dropout_residual_fwd_kernel<float, float>, n8, st, static_cast<const float*>(y), static_cast<const float*>(resid),
                                                                    n, p, seed, offset, offset_dev, static_cast<float*>(out), keep_mask);
  else if (dtype == GVIT_F32)
    stream_launch(dropout_residual_fwd_kernel<float, bf>, n8, st, static_cast<const bf*>(y), static_cast<const float*>(resid), n, p,
                                                                 seed, offset, offset_dev, static_cast<float*>(out), keep_mask);
  else
    stream_launch(dropout_residual_fwd_kernel<bf, bf>, n8, st, static_cast<const bf*>(y), static_cast<const bf*>(resid), n, p, seed,
                                                              offset, offset_dev, static_cast<bf*>(out), keep_mask);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int dropout_bwd(const void* dout, const uint8_t* keep_mask, int64_t n, float p, int dtype, int y_dtype, void* dy,
                int D, int skip_period, float* colsum_out, float* partial_ws, cudaStream_t st) {
  using bf = __nv_bfloat16;
  if (colsum_out) {
    int cb, nch, rpc;
    const int64_t rows = n / D;
    colsum_grid(rows, D, 6, &cb, &nch, &rpc);
    dim3 g2(cb, nch);
    if (dtype == GVIT_F32 && y_dtype == GVIT_F32)
      edge_bwd_colsum_kernel<float, float, false><<<g2, 256, 0, st>>>(static_cast<const float*>(dout), nullptr, keep_mask, rows, D, rpc, p, skip_period, static_cast<float*>(dy), partial_ws);
    else if (dtype == GVIT_F32)
      edge_bwd_colsum_kernel<float, bf, false><<<g2, 256, 0, st>>>(static_cast<const float*>(dout), nullptr, keep_mask, rows, D, rpc, p, skip_period, static_cast<bf*>(dy), partial_ws);
    else
      edge_bwd_colsum_kernel<bf, bf, false><<<g2, 256, 0, st>>>(static_cast<const bf*>(dout), nullptr, keep_mask, rows, D, rpc, p, skip_period, static_cast<bf*>(dy), partial_ws);
    GVIT_CHECK_LAUNCH();
    colsum_final_kernel<<<(D + 31) / 32, 256, 0, st>>>(partial_ws, nch, D, colsum_out);
    GVIT_CHECK_LAUNCH();
    return GVIT_OK;
  }
  const int64_t n8 = n / 8;
  if (dtype == GVIT_F32 && y_dtype == GVIT_F32)
    stream_launch(dropout_bwd_kernel<float, float>, n8, st, static_cast<const float*>(dout), keep_mask, n, p, static_cast<float*>(dy));
  else if (dtype == GVIT_F32)
    stream_launch(dropout_bwd_kernel<float, bf>, n8, st, static_cast<const float*>(dout), keep_mask, n, p, static_cast<bf*>(dy));
  else
    stream_launch(dropout_bwd_kernel<bf, bf>, n8, st, static_cast<const bf*>(dout), keep_mask, n, p, static_cast<bf*>(dy));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int gelu_dropout_fwd(const void* u, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype,
                     void* out, uint8_t* keep_mask, cudaStream_t st) {
  const int64_t n8 = n / 8;
  if (dtype == GVIT_F32)
    stream_launch(gelu_dropout_fwd_kernel<float>, n8, st, static_cast<const float*>(u), n, p, seed, offset, offset_dev, static_cast<float*>(out), keep_mask);
  else
    stream_launch(gelu_dropout_fwd_kernel<__nv_bfloat16>, n8, st, static_cast<const __nv_bfloat16*>(u), n, p, seed, offset, offset_dev,
                                                                 static_cast<__nv_bfloat16*>(out), keep_mask);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int gelu_dropout_bwd(const void* dout, const void* u, const uint8_t* keep_mask, int64_t n, float p, int dtype, void* du,
                     int D, float* colsum_out, float* partial_ws, cudaStream_t st) {
  if (colsum_out) {
    using bf = __nv_bfloat16;
    int cb, nch, rpc;
    const int64_t rows = n / D;
    colsum_grid(rows, D, 12, &cb, &nch, &rpc);               // 3 waves of 4 resident CTAs: measured best of 3..48
    dim3 g2(cb, nch);
    if (dtype == GVIT_F32)
      edge_bwd_colsum_kernel<float, float, true><<<g2, 256, 0, st>>>(static_cast<const float*>(dout), static_cast<const float*>(u), keep_mask, rows, D, rpc, p, 0, static_cast<float*>(du), partial_ws);
    else
      edge_bwd_colsum_kernel<bf, bf, true><<<g2, 256, 0, st>>>(static_cast<const bf*>(dout), static_cast<const bf*>(u), keep_mask, rows, D, rpc, p, 0, static_cast<bf*>(du), partial_ws);
    GVIT_CHECK_LAUNCH();
    colsum_final_kernel<<<(D + 31) / 32, 256, 0, st>>>(partial_ws, nch, D, colsum_out);
    GVIT_CHECK_LAUNCH();
    return GVIT_OK;
  }
  const int64_t n8 = n / 8;
  if (dtype == GVIT_F32)
    stream_launch(gelu_dropout_bwd_kernel<float>, n8, st, static_cast<const float*>(dout), static_cast<const float*>(u), keep_mask, n, p,
                                                         static_cast<float*>(du));
  else
    stream_launch(gelu_dropout_bwd_kernel<__nv_bfloat16>, n8, st, static_cast<const __nv_bfloat16*>(dout),
                                                                 static_cast<const __nv_bfloat16*>(u), keep_mask, n, p,
                                                                 static_cast<__nv_bfloat16*>(du));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

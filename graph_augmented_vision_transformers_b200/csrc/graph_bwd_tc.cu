// graph_bwd_tc.cu - backward of the graph sub-layer's sparse stages for bf16 on tcgen05.
//
// Given dz = dY Wg (the projection gradient, a library GEMM on the host side) the SIMT path needs four gather kernels
// (reverse adjacency, dvals, dp, kNN backward) that re-read ~16 neighbour rows per token from L2: 450 us at
// B = 256.  Here the same algebra is two dense per-image GEMMs on the tensor cores:
//
//   kernel A  G = dZ P^T (Np x Np, K = D), thread <-> row picks dw_ij = G[i, idx_ij] out of TMEM and forms
//             dvals_ij = w_ij (dw_ij - sum_s w_is dw_is)                    (softmax backward; dS of G1-G3)
//   kernel B  dp = [ A~^T | M3 ] [ dZ ; P ]   (one GEMM, K = 2 Np) with two sparse-but-densely-stored coefficient
//             blocks built in shared memory per 128-row tile:
//               A~^T[j,i] = w_ij                                             (aggregation backward, G5)
//               M3[j,i]   = rn_j rn_i (dS_ji + dS_ij) - [i == j] rn_j^2 t_j  (similarity + L2-normalise backward, G1-G2)
//               t_j       = sum_i (dS_ji + dS_ij) S_ji                        (the radial component p^_j . dp^_j, taken
//                                                                            from the saved similarities instead of a
//                                                                            second pass over D)
// No reverse-adjacency pass is needed: A~^T and dS^T are produced by scattering transposed coordinates; t_j is
// accumulated in 64-bit fixed point so the result does not depend on the order of the shared-memory atomics.
//
// Tile conventions are those of tc.cuh: [rows][64 bf16] 128B-swizzled tiles, K-major or MN-major by descriptor.
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int TILE = 128 * 128;              // [128 rows][64 bf16]

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// =================================================================================================
// kernel A: dvals
// =================================================================================================
constexpr int A_STAGES = 3;
constexpr int A_HALF = 256 * 128;            // one operand slab, up to [256][64]
constexpr int A_STAGE_BYTES = 2 * A_HALF;    // dZ slab | P slab
constexpr int A_THREADS = 320;
struct __align__(8) ACtrl {
  uint64_t full[A_STAGES], empty[A_STAGES], accum_full;
  uint32_t tmem_base;
};
constexpr size_t A_SMEM = (size_t)A_STAGES * A_STAGE_BYTES + sizeof(ACtrl);

template <int KT>
__global__ void __launch_bounds__(A_THREADS, 1) graph_dvals_tc_kernel(const __grid_constant__ CUtensorMap tm_dz,
                                                                      const __grid_constant__ CUtensorMap tm_p, int Np,
                                                                      int D, int k, int NT,
                                                                      const int32_t* __restrict__ idx,
                                                                      const float* __restrict__ w,
                                                                      float* __restrict__ dvals) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stages = smem_raw;
  if ((smem_u32(stages) & 1023u) != 0) __trap();
  ACtrl* ctl = reinterpret_cast<ACtrl*>(stages + (size_t)A_STAGES * A_STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int slabs = D / 64;
  const int mtiles = Np > 128 ? 2 : 1;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_dz);
    prefetch_tmap(&tm_p);
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->accum_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    if (elect_one()) {
      for (int it = 0; it < slabs; ++it) {
        const int s = it % A_STAGES;
        mbar_wait(&ctl->empty[s], ((it / A_STAGES) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[s], (uint32_t)(2 * NT * 128));
        tma_load_3d(stages + (size_t)s * A_STAGE_BYTES, &tm_dz, it * 64, 0, b, &ctl->full[s]);
        tma_load_3d(stages + (size_t)s * A_STAGE_BYTES + A_HALF, &tm_p, it * 64, 0, b, &ctl->full[s]);
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, NT, false, false);
      for (int it = 0; it < slabs; ++it) {
        const int s = it % A_STAGES;
        mbar_wait(&ctl->full[s], (it / A_STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(stages + (size_t)s * A_STAGE_BYTES);
        for (int mt = 0; mt < mtiles; ++mt)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + mt * 256, make_sdesc(base + mt * TILE + kk * 32), make_sdesc(base + A_HALF + kk * 32), idesc,
                    it > 0 || kk > 0);
        umma_commit(&ctl->empty[s]);
      }
      umma_commit(&ctl->accum_full);
    }
  } else {
    const int g = warp >> 2;
    const int wrow0 = g * 128 + (warp & 3) * 32;
    const int row = wrow0 + lane;
    if (g < mtiles && wrow0 < Np) {                      // warp-uniform
      const bool valid = row < Np;
      const int64_t o = ((int64_t)b * Np + (valid ? row : 0)) * k;
      int nb[KT];
      float wj[KT], dw[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const bool on = valid && j < k;
        nb[j] = on ? idx[o + j] : -1;
        wj[j] = on ? w[o + j] : 0.f;
        dw[j] = 0.f;
      }
      mbar_wait(&ctl->accum_full, 0);
      tc_fence_after();
      const uint32_t trow = tmem_lane_base(tmem, warp) + g * 256;
      for (int c0 = 0; c0 < NT; c0 += 32) {
        float v[32];
        if (NT - c0 >= 32) {
          tmem_ld32(trow + c0, v);
        } else {                                          // 16-column tail: do not read columns the MMA never wrote
          float v16[16];
          tmem_ld16(trow + c0, v16);
#pragma unroll
          for (int t = 0; t < 16; ++t) { v[t] = v16[t]; v[16 + t] = 0.f; }
        }
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int col = c0 + t;
#pragma unroll
          for (int j = 0; j < KT; ++j) dw[j] = (nb[j] == col) ? v[t] : dw[j];
        }
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < KT; ++j) s = fmaf(wj[j], dw[j], s);
      if (valid) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (j < k) dvals[o + j] = wj[j] * (dw[j] - s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// =================================================================================================
// kernel B: dp = [A~^T | M3] [dZ ; P]
// =================================================================================================
constexpr int B_THREADS = 192;
constexpr float FIX_SCALE = 68719476736.0f;          // 2^36: fixed-point scale of the t_j accumulator
struct __align__(8) BCtrl {
  float rn[256];
  long long tfix[128];
  uint64_t full[2], empty[2], a_ready, out_full[2], out_free[2];
  uint32_t tmem_base;
};

struct BParams {
  int Np, D, k, NT, nblk, stage_bytes;
  const int32_t* idx;
  const float* w;
  const float* vals;
  const float* dvals;
  const float* rnorm;
};

__device__ __forceinline__ __nv_bfloat16* a2_cell(uint8_t* sA, int row, int col) {
  return reinterpret_cast<__nv_bfloat16*>(sA + (col >> 6) * TILE + swz128(row, col & 63) + (col & 7) * 2);
}

__global__ void __launch_bounds__(B_THREADS, 1) graph_dp_tc_kernel(const __grid_constant__ CUtensorMap tm_dz,
                                                                   const __grid_constant__ CUtensorMap tm_p,
                                                                   const __grid_constant__ CUtensorMap tm_dp,
                                                                   const BParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;                                  // nblk blocks of [128][64]: the coefficient tile, K-major
  if ((smem_u32(sA) & 1023u) != 0) __trap();
  uint8_t* sStage = sA + (size_t)P.nblk * TILE;            // 2 stages of { dZ slab [NT][64] | P slab [NT][64] }
  BCtrl* ctl = reinterpret_cast<BCtrl*>(sStage + 2 * (size_t)P.stage_bytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int mt = blockIdx.x, b = blockIdx.y;
  const int j0 = mt * 128;
  const int slabs = P.D / 64;
  const int NT = P.NT;
  GVIT_TRACE_DECL

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tm_dz);
    prefetch_tmap(&tm_p);
    prefetch_tmap(&tm_dp);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
      mbar_init(&ctl->out_full[s], 1);
      mbar_init(&ctl->out_free[s], 128);
    }
    mbar_init(&ctl->a_ready, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&ctl->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 4) {
    if (elect_one()) {
      for (int s = 0; s < slabs; ++s) {
        const int st = s & 1;
        mbar_wait(&ctl->empty[st], ((s >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[st], (uint32_t)(2 * NT * 128));
        uint8_t* dst = sStage + (size_t)st * P.stage_bytes;
        tma_load_3d(dst, &tm_dz, s * 64, 0, b, &ctl->full[st]);
        tma_load_3d(dst + NT * 128, &tm_p, s * 64, 0, b, &ctl->full[st]);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, 64, false, true);     // coefficients K-major, [dZ;P] slab MN-major
      const uint32_t aA = smem_u32(sA);
      mbar_wait(&ctl->a_ready, 0);
      tc_fence_after();
      for (int s = 0; s < slabs; ++s) {
        const int st = s & 1;
        GVIT_TR(1);
        mbar_wait(&ctl->full[st], (s >> 1) & 1);
        GVIT_TR(2);
        mbar_wait(&ctl->out_free[st], ((s >> 1) & 1) ^ 1);          // accumulator buffer drained by the epilogue
        tc_fence_after();
        GVIT_TR(3);
        const uint32_t aS = smem_u32(sStage + (size_t)st * P.stage_bytes);
        for (int kk = 0; kk < 2 * NT / 16; ++kk)                    // K = [Np rows of dZ | Np rows of P]
          umma_ss(tmem + st * 64, make_sdesc(aA + (kk >> 2) * TILE + (kk & 3) * 32), make_sdesc(aS + kk * 2048), idesc,
                  kk > 0);
        umma_commit(&ctl->out_full[st]);
        GVIT_TR(4);
      }
    }
  } else {
    const int tid = threadIdx.x;                                    // 0..127 == tile row == TMEM lane
    const int jg = j0 + tid;
    const bool valid = jg < P.Np;
    const int E = P.Np * P.k;
    const int32_t* idx_b = P.idx + (int64_t)b * E;
    const float* w_b = P.w + (int64_t)b * E;
    const float* v_b = P.vals + (int64_t)b * E;
    const float* ds_b = P.dvals + (int64_t)b * E;
    // ---- coefficient tile ----------------------------------------------------------------------------------------
    // Every global load of the build is issued up front (edge ids of the whole image, then the payload of the edges
    // that land in this tile): two memory latencies in total.  v1 chased one latency per edge (17k cycles per CTA).
    {
      constexpr int EPT = 16;                                       // edges per thread: 128 * 16 = 2048 per pass
      // t_j in 64-bit fixed point as two native 32-bit shared atomics (lo with carry-out into hi): the sum is the same
      // for every arrival order, and there is no 64-bit CAS loop (ATOMS.CAST.SPIN) under contention
      auto tfix_add = [&](int j, long long v) {
        unsigned int* cell = reinterpret_cast<unsigned int*>(&ctl->tfix[j]);
        const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(v));
        const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(v) >> 32);
        const unsigned int old = atomicAdd(cell, lo);
        atomicAdd(cell + 1, hi + ((old + lo) < old ? 1u : 0u));
      };
      int ii1[8];
      float ds1[8], vv1[8];
      const int kk1 = min(P.k, 8);
      if (valid) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = jg * P.k + min(u, P.k - 1);
          ii1[u] = idx_b[e];
          ds1[u] = ds_b[e];
          vv1[u] = v_b[e];
        }
      }
      int jj[EPT];
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int e = u * 128 + tid;
        jj[u] = e < E ? idx_b[e] - j0 : -1;
      }
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < P.nblk * TILE / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = z4;
      for (int i = tid; i < 256; i += 128) ctl->rn[i] = i < P.Np ? P.rnorm[(int64_t)b * P.Np + i] : 0.f;
      ctl->tfix[tid] = 0;
      float ww[EPT], dsv[EPT], vv[EPT];
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int e = u * 128 + tid;
        if (jj[u] >= 0 && jj[u] < 128) { ww[u] = w_b[e]; dsv[u] = ds_b[e]; vv[u] = v_b[e]; }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      GVIT_TR(10);
      // phase 1: the row's own k entries of dS (forward edges j -> i)
      if (valid) {
        const float rnj = ctl->rn[jg];
        long long tacc = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (u < kk1) {
            *a2_cell(sA, tid, NT + ii1[u]) = __float2bfloat16_rn(rnj * ctl->rn[ii1[u]] * ds1[u]);
            tacc += __float2ll_rn(ds1[u] * vv1[u] * FIX_SCALE);
          }
        }
        for (int s = 8; s < P.k; ++s) {                             // k > 8: the remaining own entries, one by one
          const int e = jg * P.k + s;
          const int i = idx_b[e];
          const float ds = ds_b[e];
          *a2_cell(sA, tid, NT + i) = __float2bfloat16_rn(rnj * ctl->rn[i] * ds);
          tacc += __float2ll_rn(ds * v_b[e] * FIX_SCALE);
        }
        tfix_add(tid, tacc);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // phase 2: every edge i -> j of the image that lands in this tile: A~^T[j,i] = w_e and M3[j,i] += rn_j rn_i dS_e.
      // (j,i) pairs are unique over the edges (a row's k neighbours are distinct), so the 16-bit updates do not race.
      auto apply_edge = [&](int e, int j, float we, float ds, float v) {
        const int i = e / P.k;
        *a2_cell(sA, j, i) = __float2bfloat16_rn(we);
        __nv_bfloat16* c = a2_cell(sA, j, NT + i);
        *c = __float2bfloat16_rn(__bfloat162float(*c) + ctl->rn[j0 + j] * ctl->rn[i] * ds);
        tfix_add(j, __float2ll_rn(ds * v * FIX_SCALE));
      };
#pragma unroll
      for (int u = 0; u < EPT; ++u)
        if (jj[u] >= 0 && jj[u] < 128) apply_edge(u * 128 + tid, jj[u], ww[u], dsv[u], vv[u]);
      for (int e = EPT * 128 + tid; e < E; e += 128) {              // images with more than 2048 edges
        const int j = idx_b[e] - j0;
        if (j >= 0 && j < 128) apply_edge(e, j, w_b[e], ds_b[e], v_b[e]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // phase 3: the radial term on the diagonal
      if (valid) {
        const float rnj = ctl->rn[jg];
        const float t = static_cast<float>(static_cast<double>(ctl->tfix[tid]) * (1.0 / (double)FIX_SCALE));
        __nv_bfloat16* c = a2_cell(sA, tid, NT + jg);
        *c = __float2bfloat16_rn(__bfloat162float(*c) - rnj * rnj * t);
      }
      fence_async_smem();
      mbar_arrive(&ctl->a_ready);
      GVIT_TR(11);
    }
    // ---- per slab: accumulator -> bf16 -> the (consumed) stage -> TMA tile store ------------------------------------
    const uint32_t tO = tmem_lane_base(tmem, warp);
    for (int s = 0; s < slabs; ++s) {
      const int st = s & 1;
      GVIT_TR(12);
      mbar_wait(&ctl->out_full[st], (s >> 1) & 1);
      tc_fence_after();
      GVIT_TR(13);
      float v0[32], v1[32];
      tmem_ld32(tO + st * 64, v0);
      tmem_ld32(tO + st * 64 + 32, v1);
      tc_fence_before();
      mbar_arrive(&ctl->out_free[st]);
      uint8_t* so = sStage + (size_t)st * P.stage_bytes;        // every MMA that read this stage has retired
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o4;
        o4.x = pack2(v0[8 * q + 0], v0[8 * q + 1]); o4.y = pack2(v0[8 * q + 2], v0[8 * q + 3]);
        o4.z = pack2(v0[8 * q + 4], v0[8 * q + 5]); o4.w = pack2(v0[8 * q + 6], v0[8 * q + 7]);
        *reinterpret_cast<uint4*>(so + swz128(tid, 8 * q)) = o4;
        o4.x = pack2(v1[8 * q + 0], v1[8 * q + 1]); o4.y = pack2(v1[8 * q + 2], v1[8 * q + 3]);
        o4.z = pack2(v1[8 * q + 4], v1[8 * q + 5]); o4.w = pack2(v1[8 * q + 6], v1[8 * q + 7]);
        *reinterpret_cast<uint4*>(so + swz128(tid, 32 + 8 * q)) = o4;
      }
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        tma_store_3d(&tm_dp, so, s * 64, j0, b);                 // rows >= Np are clipped by the TMA unit
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&ctl->empty[st]);                            // the producer may refill this stage
      }
      GVIT_TR(14);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 128);
}

template <int KT>
int launch_a(const CUtensorMap& tm_dz, const CUtensorMap& tm_p, const Tokens& t, int k, int NT, const int32_t* idx,
             const float* w, float* dvals, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(graph_dvals_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM));
  graph_dvals_tc_kernel<KT><<<t.B, A_THREADS, A_SMEM, st>>>(tm_dz, tm_p, t.Np, t.D, k, NT, idx, w, dvals);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

inline size_t b_smem(int NT, int* nblk, int* stage_bytes) {
  *nblk = (2 * NT + 63) / 64;
  *stage_bytes = 2 * NT * 128 < TILE ? TILE : 2 * NT * 128;   // also holds the [128][64] output staging tile
  return (size_t)*nblk * TILE + 2 * (size_t)*stage_bytes + sizeof(BCtrl);
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_graph_bwd)

bool graph_bwd_tc_supported(int Np, int D, int k) {
  int nblk, sb;
  const int NT = (Np + 15) & ~15;
  return Np >= 16 && Np <= 256 && D >= 64 && D % 64 == 0 && D <= 1024 && k <= 16 && b_smem(NT, &nblk, &sb) <= 227 * 1024;
}

int graph_bwd_tc(const Tokens& t, int k, const int32_t* idx, const float* vals, const float* w, const float* rnorm,
                 const void* dz, int64_t dz_batch_stride, float* dvals, void* dp, cudaStream_t st) {
  const int NT = (t.Np + 15) & ~15;
  GVIT_REQUIRE(t.B <= 65535, GVIT_ERR_SHAPE, "graph_bwd: batch %d exceeds the grid limit 65535", t.B);
  CUtensorMap tm_dz, tm_p, tm_dp;
  int rc = make_tmap_bf16_3d(&tm_dz, dz, t.D, t.Np, t.B, t.D, (uint64_t)dz_batch_stride, NT);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_p, t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, NT);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_dp, dp, t.D, t.Np, t.B, t.row_stride, t.batch_stride, 128);
  if (rc != GVIT_OK) return rc;
  if (k <= 4) rc = launch_a<4>(tm_dz, tm_p, t, k, NT, idx, w, dvals, st);
  else if (k <= 8) rc = launch_a<8>(tm_dz, tm_p, t, k, NT, idx, w, dvals, st);
  else rc = launch_a<16>(tm_dz, tm_p, t, k, NT, idx, w, dvals, st);
  if (rc != GVIT_OK) return rc;

  BParams P;
  P.Np = t.Np; P.D = t.D; P.k = k; P.NT = NT;
  const size_t smem = b_smem(NT, &P.nblk, &P.stage_bytes);
  P.idx = idx; P.w = w; P.vals = vals; P.dvals = dvals; P.rnorm = rnorm;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(graph_dp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((t.Np + 127) / 128, t.B);
  graph_dp_tc_kernel<<<grid, B_THREADS, smem, st>>>(tm_dz, tm_p, tm_dp, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

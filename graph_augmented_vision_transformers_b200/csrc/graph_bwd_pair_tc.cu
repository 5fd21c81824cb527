// graph_bwd_pair_tc.cu - backward of the graph sub-layer's sparse stages for bf16, 128 < Np <= 256,
// D % 128 == 0, as a 2-SM kernel: ONE CTA PAIR PER IMAGE (`tcgen05.mma.cta_group::2`).  Same algebra as the one-CTA fused kernel
// (graph_bwd_tc.cu: G = dZ P^T -> softmax backward -> dp = [A~^T | M3] [dZ ; P]); what changes is who holds what.
//
// The one-CTA kernel keeps ONE 128-row coefficient tile in shared memory (106 KB at Np = 196), so it walks the image's
// [dZ ; P] slabs once per row tile in phase B (two passes of 6 x 106 KB), after streaming all of dZ and P in phase A: 1.9 MB of
// L2 -> shared-memory operand traffic per image and CTA, and its trace shows phase B at ~4.4k cycles per 128-feature chunk
// against 1.7k of tensor time - the operand stream (~24 B/clk/SM) is the bound.  As a pair, rank r owns row tile r: its
// 128 rows of dZ (A of the Gram product), its coefficient tile (A of the output product), its TMEM lanes, its output
// rows - and every B operand is split between the two CTAs: phase A stages half of P's tokens, phase B half of each
// 128-feature chunk (one 64-feature slab of dZ and one of P).  Both row tiles are served by the SAME pass over the slabs
// (M = 256), so a CTA streams 0.35 + 0.32 MB per image instead of 1.9 MB.
// Measured (profiles/r4n_*): 0.1018 ms against 0.1131 ms for the one-CTA kernel at B = 256 (same box); 0.0923 ms after r5.  Both of its GEMM phases
// now run at the chip-wide rate at which L2 delivers UNIQUE operand bytes (~6.5 TB/s: phase A 12-14k cycles per image while
// all pairs stream, 6.7k once the others have finished; phase B 2.0-2.2k cycles per chunk against 1.7k of tensor time), and
// the extraction / build in between (11-12k cycles per image) streams nothing.  Starting the odd pairs 12k / 20k / 28k cycles
// late so that the phases interleave chip-wide changed nothing (0.1007 / 0.1048 / 0.1108 ms).
//
// Cross-CTA data: the coefficient tile of a row tile needs dvals (dS) of ALL rows of the image.  For k <= 8 (the image's
// 256 x k values fit behind the control block) every row thread sends its k values straight into the peer's shared memory
// with `st.async ... mbarrier::complete_tx::bytes`, the peer's max |dS S| (fixed-point scale of the radial term) follows the
// same way, and the receiving CTA's one arrival is its own expect_tx of those bytes: no fence, no global round trip
// (r5: 0.0964 -> 0.0923 ms; the same exchange as st.shared::cluster + one mbarrier.arrive.release.cluster made the
// arriving thread wait ~6k cycles behind the CTA's global dvals stores).  Larger k: through global memory (dvals is an
// output anyway) behind a release / acquire mbarrier handshake at cluster scope.  The row thread keeps its neighbour list,
// similarities and dS entries in registers from the extraction into the build; the peer-independent build operands are
// fetched and the tile is zeroed while the exchange is in flight (r5: 0.1016 -> 0.0964 ms).
// TMEM per CTA: G [0, NT) | output chunk buffers [256, 384), [384, 512).  Shared memory: coefficient tile (nblk x 16 KB,
// doubles as the extraction scratch) | ring of four NT x 64 slots.
// Warp roles: 0-7 workers (extract: warps 0-3, thread <-> row; build and output: all), 8 TMA producer (both CTAs), 9 MMA
// issuer (leader) + TMEM owner.
#include <float.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int TILE = 128 * 128;              // [128 rows][64 bf16]
constexpr int Q_THREADS = 320;
constexpr int Q_SLOTS = 4;
constexpr int Q_ASLOTS = 6;                  // phase-A slab slots of 32 KB: three in the ring region, three in the (idle) coefficient region
constexpr int Q_ASLOT_BYTES = 32 * 1024;     // [128 rows of dZ][64] | [NT/2 tokens of P][64]

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ __nv_bfloat16* a2_cell(uint8_t* sA, int row, int col) {
  return reinterpret_cast<__nv_bfloat16*>(sA + (col >> 6) * TILE + swz128(row, col & 63) + (col & 7) * 2);
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("gvit: graph_bwd pair handshake timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// asynchronous stores into the peer CTA's shared memory that count their bytes on the peer's mbarrier: the data is visible to
// whoever observes that barrier's phase complete - no release fence on the sending side (an mbarrier.arrive.release.cluster
// behind 128 threads' global dvals stores cost the arriving thread ~6k cycles)
__device__ __forceinline__ void st_async_b32(uint32_t dst_cluster, uint32_t v, uint32_t mbar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster), "r"(v), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t dst_cluster, float4 v, uint32_t mbar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst_cluster), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                 "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

struct __align__(8) QCtrl {
  float rn[256];
  int tfix[128];                               // t_j in fixed point (per-image scale), one fire-and-forget RED.ADD per edge
  int tmax_own[2], tmax_peer[2];               // max |dS_e S_e| over this CTA's / the peer's rows, per image parity (float bits)
  uint64_t full[Q_SLOTS], empty[Q_SLOTS], fullA[Q_ASLOTS], emptyA[Q_ASLOTS], g_full, a_ready, coef_free, out_full[2], out_free[2], peer;
  uint32_t tmem_base;
};
struct QParams {
  int B, Np, D, k, NT, nblk, slot_bytes;
  int dsx;                                     // the image's dvals (256 x k fp32) fit behind QCtrl: the pair exchanges them through
                                               // distributed shared memory instead of global memory + __threadfence
  int kvec;                                    // k in {4, 8, 16} == KT and idx / w / vals / dvals 16-byte aligned: whole rows by vector access
  uint32_t kmagic;                             // ceil(2^32 / k): e / k == __umulhi(e, kmagic) for e < 2^16
  const int32_t* idx;
  const float* w;
  const float* vals;
  float* dvals;                                // written (extract) by the row's CTA, re-read (build) by both CTAs of the pair
  const float* rnorm;
  __nv_bfloat16* dp;
  int64_t dp_bs, dp_rs;                        // batch / row stride of dp in elements
};

template <int KT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q_THREADS, 1) graph_bwd_pair_kernel(
    const __grid_constant__ CUtensorMap tm_dzA, const __grid_constant__ CUtensorMap tm_pB,
    const __grid_constant__ CUtensorMap tm_dzF, const __grid_constant__ CUtensorMap tm_pF, const QParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;                                  // nblk blocks of [128][64]: coefficient tile / extraction scratch
  if ((smem_u32(sA) & 1023u) != 0) __trap();
  uint8_t* ring = sA + (size_t)P.nblk * TILE;              // Q_SLOTS slots of [NT][64]
  QCtrl* ctl = reinterpret_cast<QCtrl*>(ring + (size_t)Q_SLOTS * P.slot_bytes);
  float* ds_img = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ctl + 1) + 15) & ~static_cast<uintptr_t>(15));   // P.dsx: dS of every edge of the image, [256 rows][k], both CTAs hold a copy

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int NT = P.NT, Np = P.Np, k = P.k, NH = P.NT / 2;
  const int slabs = P.D / 64;
  const int fblocks = P.D / 128;
  const int ksteps = NT / 16;
  GVIT_TRACE_DECL
  GVIT_SPAN(0);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_dzA); prefetch_tmap(&tm_pB); prefetch_tmap(&tm_dzF); prefetch_tmap(&tm_pF);
    for (int s = 0; s < Q_SLOTS; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < Q_ASLOTS; ++s) { mbar_init(&ctl->fullA[s], 1); mbar_init(&ctl->emptyA[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->out_full[s], 1); mbar_init(&ctl->out_free[s], 16); }   // leader's: 8 warps x 2 CTAs
    mbar_init(&ctl->g_full, 1);
    mbar_init(&ctl->a_ready, 16);
    mbar_init(&ctl->coef_free, 1);
    mbar_init(&ctl->peer, 1);
    ctl->tmax_own[0] = ctl->tmax_own[1] = 0;
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2sm(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t pi = 0;
      auto load = [&](const CUtensorMap* tm, int c0, int row0, int b, uint32_t pair_bytes) {
        const uint32_t s = pi & (Q_SLOTS - 1);
        mbar_wait(&ctl->empty[s], ((pi / Q_SLOTS) & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(&ctl->full[s], pair_bytes);
        tma_load_3d_2sm(ring + (size_t)s * P.slot_bytes, tm, c0, row0, b, mapa_u32(smem_u32(&ctl->full[s]), 0));
        ++pi;
      };
      // Phase A (the Gram product) wants many slabs in flight - a slab load takes ~2.3k cycles from request to consumption
      // under load, and with the two ring slots per slab of the first version it ran at 1.4-1.6k cycles per slab (17-20k
      // per image against 5k of tensor time, traces v1 / v2).  During phase A the coefficient region is idle (the previous
      // image's output product has retired, this image's extraction has not started), so phase A uses its own geometry: six
      // 32 KB slab slots, three laid over the ring region and three over the coefficient region, with their own barriers.
      // The two geometries never overlap in time: phase A starts once every ring slot of the previous phase B has been
      // consumed, phase B once every phase-A slot has.  The phase-A boxes of the NEXT image are requested into L2 early.
      uint32_t pa = 0, cntB[Q_SLOTS] = {0, 0, 0, 0};
      auto aslot = [&](uint32_t j) { return j < 3 ? ring + (size_t)j * Q_ASLOT_BYTES : sA + (size_t)(j - 3) * Q_ASLOT_BYTES; };
      auto prefetch_a = [&](int b) {
        for (int s = 0; s < slabs; ++s) {
          tma_prefetch_3d(&tm_dzA, s * 64, rank * 128, b);
          tma_prefetch_3d(&tm_pB, s * 64, rank * NH, b);
        }
      };
      Item I, Inext;
      for (int itp = 0; get_item(itp, cid, ncl, P.B, fblocks, I); ++itp) {
        const int b = I.b;
        // every ring slot of the previous image's phase B consumed (its last commit also retires the MMAs that read the
        // coefficient tile)
        for (int q = 0; q < Q_SLOTS; ++q)
          if (cntB[q] > 0) mbar_wait(&ctl->empty[q], (cntB[q] - 1) & 1);
        for (int s = 0; s < slabs; ++s, ++pa) {                               // phase A: own 128 rows of dZ | own half of P's tokens
          const uint32_t j = pa % Q_ASLOTS;
          mbar_wait(&ctl->emptyA[j], ((pa / Q_ASLOTS) & 1) ^ 1);
          const uint32_t fullL = mapa_u32(smem_u32(&ctl->fullA[j]), 0);
          if (rank == 0) mbar_expect_tx(&ctl->fullA[j], 2u * (128u + (uint32_t)NH) * 128u);
          tma_load_3d_2sm(aslot(j), &tm_dzA, s * 64, rank * 128, b, fullL);
          tma_load_3d_2sm(aslot(j) + 128 * 128, &tm_pB, s * 64, rank * NH, b, fullL);
        }
        if (get_item(itp + 1, cid, ncl, P.B, fblocks, Inext)) prefetch_a(Inext.b);
        // every phase-A slot consumed before the ring geometry of phase B is written
        for (uint32_t q = 0; q < (uint32_t)Q_ASLOTS && q < pa; ++q) {
          const uint32_t last = pa - 1 - ((pa - 1 - q) % Q_ASLOTS + Q_ASLOTS) % Q_ASLOTS;   // last load index that used slot q
          mbar_wait(&ctl->emptyA[q], (last / Q_ASLOTS) & 1);
        }
        for (int f = I.c0; f < I.c1; ++f) {                                   // phase B: own 64 features of the chunk, all tokens
          cntB[pi & (Q_SLOTS - 1)]++;
          load(&tm_dzF, (2 * f + rank) * 64, 0, b, 2u * (uint32_t)NT * 128u);
          cntB[pi & (Q_SLOTS - 1)]++;
          load(&tm_pF, (2 * f + rank) * 64, 0, b, 2u * (uint32_t)NT * 128u);
        }
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc_g = make_idesc(256, NT, false, false);
      const uint32_t idesc_o = make_idesc(256, 128, false, true);              // coefficients K-major, slab pair MN-major
      const uint32_t aA = smem_u32(sA), aR = smem_u32(ring);
      uint32_t ci = 0, na = 0, uses[2] = {0, 0}, tcount = 0;
      auto wait_full = [&](uint32_t c) { mbar_wait(&ctl->full[c & (Q_SLOTS - 1)], (c / Q_SLOTS) & 1); };
      auto slot_addr = [&](uint32_t c) { return aR + (c & (Q_SLOTS - 1)) * (uint32_t)P.slot_bytes; };
      uint32_t ca = 0;                                                        // phase-A slab counter
      const uint32_t aC = smem_u32(sA);
      Item I;
      for (int itm = 0; get_item(itm, cid, ncl, P.B, fblocks, I); ++itm) {
        GVIT_TR(1);
        for (int s = 0; s < slabs; ++s, ++ca) {                               // G = dZ P^T (the previous image's G was read out
          const uint32_t j = ca % Q_ASLOTS;                                    // before its a_ready, i.e. before its phase B)
          mbar_wait(&ctl->fullA[j], (ca / Q_ASLOTS) & 1);
          tc_fence_after();
          const uint32_t aZ = (j < 3 ? aR + j * Q_ASLOT_BYTES : aC + (j - 3) * Q_ASLOT_BYTES), aP = aZ + 128 * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_2sm(tmem, make_sdesc(aZ + kk * 32), make_sdesc(aP + kk * 32), idesc_g, s > 0 || kk > 0);
          umma_commit_2sm_mc(&ctl->emptyA[j], 3);
        }
        umma_commit_2sm_mc(&ctl->g_full, 3);
        GVIT_TR(2);
        mbar_wait(&ctl->a_ready, na & 1);                                      // both CTAs' coefficient tiles are in shared memory
        ++na;
        tc_fence_after();
        GVIT_TR(3);
        for (int f = I.c0; f < I.c1; ++f) {
          const uint32_t buf = tcount & 1;
          mbar_wait(&ctl->out_free[buf], (uses[buf] & 1) ^ 1);                 // this buffer's previous chunk drained by both CTAs
          ++uses[buf];
          tc_fence_after();
          for (int half = 0; half < 2; ++half) {                               // K = [NT rows of dZ | NT rows of P]
            wait_full(ci);
            tc_fence_after();
            const uint32_t aS = slot_addr(ci);
            for (int kk = 0; kk < ksteps; ++kk) {
              const int kg = half * ksteps + kk;
              umma_ss_2sm(tmem + 256 + buf * 128, make_sdesc(aA + (kg >> 2) * TILE + (kg & 3) * 32), make_sdesc(aS + kk * 2048), idesc_o,
                          half > 0 || kk > 0);
            }
            umma_commit_2sm_mc(&ctl->empty[ci & (Q_SLOTS - 1)], 3);
            ++ci;
          }
          umma_commit_2sm_mc(&ctl->out_full[buf], 3);
          ++tcount;
          GVIT_TR(4);
        }
        umma_commit_2sm_mc(&ctl->coef_free, 3);
      }
    }
  } else {
    // ---------------------------------------------------------------- workers (256 threads per CTA)
    const int tid = threadIdx.x;
    const int E = Np * k;
    const int j0 = rank * 128;                                           // this CTA's row tile
    uint32_t it = 0, tcount = 0;
    const int wrow0 = j0 + (warp & 3) * 32;
    float* scr = reinterpret_cast<float*>(sA) + warp * (32 * 36) + lane * 36;
    const uint32_t a_readyL = mapa_u32(smem_u32(&ctl->a_ready), 0);
    const uint32_t out_freeL0 = mapa_u32(smem_u32(&ctl->out_free[0]), 0), out_freeL1 = mapa_u32(smem_u32(&ctl->out_free[1]), 0);
    const uint32_t peer_bar = mapa_u32(smem_u32(&ctl->peer), rank ^ 1);
    Item I;
    for (; get_item((int)it, cid, ncl, P.B, fblocks, I); ++it) {
      const int b = I.b;
      const int32_t* idx_b = P.idx + (int64_t)b * E;
      const float* w_b = P.w + (int64_t)b * E;
      const float* v_b = P.vals + (int64_t)b * E;
      float* ds_b = P.dvals + (int64_t)b * E;
      // thread tid < 128 owns row j0 + tid in the extraction AND in the build: its neighbour list, similarities and dS entries
      // stay in registers across the two; the edge-parallel build operands that do not depend on the peer (jj, ww, vv) are
      // fetched, and the coefficient tile is zeroed, while the dvals handshake with the peer CTA is in flight
      constexpr int EPT = 8;                                          // edges per thread: 256 * 8 = 2048 per pass
      int nb[KT], jj[EPT];
      float vj[KT], dsown[KT], ww[EPT], vv[EPT];
      // ---- extract: dvals of row (wrow0 + lane), warps 0-3 --------------------------------------------------------
      {
        const int row = wrow0 + lane;
        const bool active = warp < 4 && wrow0 < Np;                    // warp-uniform
        const bool valid = row < Np;
        const int64_t o = (int64_t)(valid ? row : 0) * k;
        float wj[KT], dw[KT];                                          // vj: the row's similarities, for the fixed-point scale below -
        if (active) {                                                  // loaded here so that the latency hides behind the Gram product
          if (P.kvec && valid) {
#pragma unroll
            for (int j = 0; j < KT; j += 4) {
              const int4 i4 = *reinterpret_cast<const int4*>(idx_b + o + j);
              const float4 w4 = *reinterpret_cast<const float4*>(w_b + o + j);
              const float4 v4 = *reinterpret_cast<const float4*>(v_b + o + j);
              nb[j] = i4.x; nb[j + 1] = i4.y; nb[j + 2] = i4.z; nb[j + 3] = i4.w;
              wj[j] = w4.x; wj[j + 1] = w4.y; wj[j + 2] = w4.z; wj[j + 3] = w4.w;
              vj[j] = v4.x; vj[j + 1] = v4.y; vj[j + 2] = v4.z; vj[j + 3] = v4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
              const bool on = valid && j < k;
              nb[j] = on ? idx_b[o + j] : -1;
              wj[j] = on ? w_b[o + j] : 0.f;
              vj[j] = on ? v_b[o + j] : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < KT; ++j) dw[j] = 0.f;
        }
        for (int i = tid; i < 256; i += 256) ctl->rn[i] = i < Np ? P.rnorm[(int64_t)b * Np + i] : 0.f;
        if (it > 0) mbar_wait(&ctl->coef_free, (it - 1) & 1);          // the scratch is the (retired) coefficient tile
        mbar_wait(&ctl->g_full, it & 1);
        tc_fence_after();
        GVIT_TR(10);
        if (active) {
          const uint32_t trow = tmem_lane_base(tmem, warp);
          for (int c0 = 0; c0 < NT; c0 += 32) {
            float v[32];
            if (NT - c0 >= 32) {
              tmem_ld32(trow + c0, v);
            } else {                                                    // 16-column tail: do not read columns the MMA never wrote
              float v16[16];
              tmem_ld16(trow + c0, v16);
#pragma unroll
              for (int t = 0; t < 16; ++t) { v[t] = v16[t]; v[16 + t] = 0.f; }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(scr + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
            for (int j = 0; j < KT; ++j)
              if (nb[j] >= c0 && nb[j] < c0 + 32) dw[j] = scr[nb[j] - c0];
          }
          GVIT_TR(20);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < KT; ++j) s = fmaf(wj[j], dw[j], s);
          if (valid) {
            float m = 0.f;
            float (&d)[KT] = dsown;
#pragma unroll
            for (int j = 0; j < KT; ++j) {
              d[j] = j < k ? wj[j] * (dw[j] - s) : 0.f;
              m = fmaxf(m, fabsf(d[j] * vj[j]));
            }
            // dvals is an output: written once per image (the pair that holds the other half of a split image computes the same
            // values); without the shared-memory exchange it is also how the two CTAs of a pair see each other's rows
            if (P.kvec && (I.primary || !P.dsx)) {
#pragma unroll
              for (int j = 0; j < KT; j += 4) *reinterpret_cast<float4*>(ds_b + o + j) = make_float4(d[j], d[j + 1], d[j + 2], d[j + 3]);
            } else if (I.primary || !P.dsx) {
#pragma unroll
              for (int j = 0; j < KT; ++j)
                if (j < k) ds_b[o + j] = d[j];
            }
            if (P.dsx) {                                               // my rows into my own copy and into the peer's
              const uint32_t peer_ds = mapa_u32(smem_u32(ds_img + o), rank ^ 1);
              if (P.kvec) {
#pragma unroll
                for (int j = 0; j < KT; j += 4) {
                  const float4 d4 = make_float4(d[j], d[j + 1], d[j + 2], d[j + 3]);
                  *reinterpret_cast<float4*>(ds_img + o + j) = d4;
                  st_async_f32x4(peer_ds + 4 * j, d4, peer_bar);
                }
              } else {
#pragma unroll
                for (int j = 0; j < KT; ++j)
                  if (j < k) { ds_img[o + j] = d[j]; st_async_b32(peer_ds + 4 * j, __float_as_uint(d[j]), peer_bar); }
              }
            }
            atomicMax(&ctl->tmax_own[it & 1], __float_as_int(m));       // non-negative floats order like their bit patterns
          }
        }
        GVIT_TR(21);
        tc_fence_before();
        if (!P.dsx) __threadfence();                                   // this CTA's dvals rows in global memory: visible to the peer
        asm volatile("bar.sync 1, 256;" ::: "memory");                 // G read out, scratch free, tmax_own final
        GVIT_TR(22);
        if (tid == 0) {
          if (P.dsx) {
            // my maximum follows my rows as one more counted store; the one arrival of MY barrier's phase is my own expect_tx
            // of everything the peer sends: its valid rows x k dS values + its maximum
            st_async_b32(mapa_u32(smem_u32(&ctl->tmax_peer[it & 1]), rank ^ 1), (uint32_t)ctl->tmax_own[it & 1], peer_bar);
            const int peer_rows = rank == 0 ? Np - 128 : 128;           // the pair kernel runs for 128 < Np <= 256 only
            mbar_expect_tx(&ctl->peer, (uint32_t)(peer_rows * k * 4 + 4));
          } else {
            st_cluster_u32(mapa_u32(smem_u32(&ctl->tmax_peer[it & 1]), rank ^ 1), (uint32_t)ctl->tmax_own[it & 1]);
            mbar_arrive_release_cluster(peer_bar);                      // "my rows of dvals and my maximum are published"
          }
          ctl->tmax_own[(it + 1) & 1] = 0;                              // nobody touches the other image parity right now
        }
        // while the handshake is in flight: everything of the build that does not need the peer's dvals
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
          const int e = u * 256 + tid;
          jj[u] = e < E ? idx_b[e] - j0 : -1;
        }
        {
          const uint4 z4 = make_uint4(0, 0, 0, 0);
          for (int i = tid; i < P.nblk * TILE / 16; i += 256) reinterpret_cast<uint4*>(sA)[i] = z4;   // scratch is free (barrier above)
          if (tid < 128) ctl->tfix[tid] = 0;
        }
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
          const int e = u * 256 + tid;
          if (jj[u] >= 0 && jj[u] < 128) { ww[u] = w_b[e]; vv[u] = v_b[e]; }
        }
        if (P.dsx) mbar_wait(&ctl->peer, it & 1);                      // the peer's rows and maximum have landed in my shared memory
        else mbar_wait_acquire_cluster(&ctl->peer, it & 1);            // ... the peer's rows are published in global memory
        GVIT_TR(11);
      }
      // fixed-point scale of t_j: every term is <= M = max |dS_e S_e| < 2^(e+1), a row sums at most 2^9 of them, so terms
      // scaled by 2^(20-e) keep |sum| < 2^31; the sum is exact integer arithmetic: identical for every arrival order
      float tscale, tinv;
      {
        const int bits = max(ctl->tmax_own[it & 1], ctl->tmax_peer[it & 1]);
        const int ex = (bits >> 23) & 0xff;
        int se = ex == 0 ? 0 : 127 + 20 - (ex - 127);
        se = se > 254 ? 254 : se;
        tscale = __int_as_float(se << 23);
        tinv = se == 0 ? 0.f : __int_as_float((254 - se) << 23);
      }
      // ---- this CTA's coefficient tile [A~^T | M3] (rows j0 .. j0+127), from the edges of the WHOLE image ----------------
      GVIT_TR(12);
      {
        auto tfix_add = [&](int j, float x) { atomicAdd(&ctl->tfix[j], __float2int_rn(x * tscale)); };
        const int jg = j0 + tid;
        const bool valid = tid < 128 && jg < Np;
        float dsv[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
          const int e = u * 256 + tid;
          if (jj[u] >= 0 && jj[u] < 128) dsv[u] = P.dsx ? ds_img[e] : __ldcg(ds_b + e);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        GVIT_TR(23);
        // phase 1: the row's own k entries of dS (forward edges j -> i)
        if (valid) {
          const float rnj = ctl->rn[jg];
          int tacc = 0;
#pragma unroll
          for (int u = 0; u < KT; ++u) {
            if (u < k) {
              if (nb_ok(nb[u], Np)) *a2_cell(sA, tid, NT + nb[u]) = __float2bfloat16_rn(rnj * ctl->rn[nb[u]] * dsown[u]);
              tacc += __float2int_rn(dsown[u] * vj[u] * tscale);
            }
          }
          atomicAdd(&ctl->tfix[tid], tacc);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        GVIT_TR(24);
        // phase 2: every edge i -> j of the image that lands in this tile: A~^T[j,i] = w_e, M3[j,i] += rn_j rn_i dS_e.
        // (j,i) pairs are unique over the edges (a row's k neighbours are distinct), so the 16-bit updates do not race.
        auto apply_edge = [&](int e, int j, float we, float ds, float v) {
          const int i = (int)__umulhi((unsigned)e, P.kmagic);        // e / k
          *a2_cell(sA, j, i) = __float2bfloat16_rn(we);
          __nv_bfloat16* c = a2_cell(sA, j, NT + i);
          *c = __float2bfloat16_rn(__bfloat162float(*c) + ctl->rn[j0 + j] * ctl->rn[i] * ds);
          tfix_add(j, ds * v);
        };
#pragma unroll
        for (int u = 0; u < EPT; ++u)
          if (jj[u] >= 0 && jj[u] < 128) apply_edge(u * 256 + tid, jj[u], ww[u], dsv[u], vv[u]);
        for (int e = EPT * 256 + tid; e < E; e += 256) {              // images with more than 2048 edges
          const int j = idx_b[e] - j0;
          if (j >= 0 && j < 128) apply_edge(e, j, w_b[e], P.dsx ? ds_img[e] : __ldcg(ds_b + e), v_b[e]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        GVIT_TR(25);
        // phase 3: the radial term on the diagonal
        if (valid) {
          const float rnj = ctl->rn[jg];
          const float t = static_cast<float>(ctl->tfix[tid]) * tinv;
          __nv_bfloat16* c = a2_cell(sA, tid, NT + jg);
          *c = __float2bfloat16_rn(__bfloat162float(*c) - rnj * rnj * t);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_readyL);            // leader: both coefficient tiles are built
        GVIT_TR(13);
      }
      // ---- output chunks of this row tile: TMEM -> bf16 -> global (warp & 3 = lane quadrant, warp >> 2 = 64-feature half) ----
      const int hsel = warp >> 2;
      for (int f = I.c0; f < I.c1; ++f, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(&ctl->out_full[buf], (tcount >> 1) & 1);
        tc_fence_after();
        GVIT_TR(14);
        const uint32_t tO = tmem_lane_base(tmem, warp) + 256 + buf * 128 + hsel * 64;
        float v0[32], v1[32];
        tmem_ld32(tO, v0);
        tmem_ld32(tO + 32, v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(buf ? out_freeL1 : out_freeL0);
        {
          // thread <-> row holds 64 features = eight 16-byte chunks.  A store instruction with one row per lane touches 32
          // lines (16 bytes each): 2048 line-tag lookups per tile paced the whole kernel (4.5k cycles per tile against
          // 2.4k of MMA).  An 8 x 8 chunk transpose inside each 8-lane group (three butterfly rounds, 48 shuffles) makes
          // store s write rows 8g + s with lanes 8g .. 8g+7 covering one full 128-byte line each: 4 lines per instruction.
          uint32_t c[8][4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            c[q][0] = pack2(v0[8 * q + 0], v0[8 * q + 1]); c[q][1] = pack2(v0[8 * q + 2], v0[8 * q + 3]);
            c[q][2] = pack2(v0[8 * q + 4], v0[8 * q + 5]); c[q][3] = pack2(v0[8 * q + 6], v0[8 * q + 7]);
            c[4 + q][0] = pack2(v1[8 * q + 0], v1[8 * q + 1]); c[4 + q][1] = pack2(v1[8 * q + 2], v1[8 * q + 3]);
            c[4 + q][2] = pack2(v1[8 * q + 4], v1[8 * q + 5]); c[4 + q][3] = pack2(v1[8 * q + 6], v1[8 * q + 7]);
          }
#pragma unroll
          for (int m = 1; m <= 4; m <<= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              if (a & m) continue;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t send = up ? c[a][e] : c[a | m][e];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
                if (up) c[a][e] = recv; else c[a | m][e] = recv;
              }
            }
          }
          // slot s of lane 8g + j now holds chunk j of row 8g + s
          const int rbase = j0 + (warp & 3) * 32 + (lane & ~7);
          __nv_bfloat16* dst = P.dp + (int64_t)b * P.dp_bs + (int64_t)rbase * P.dp_rs + f * 128 + hsel * 64 + (lane & 7) * 8;
#pragma unroll
          for (int s8 = 0; s8 < 8; ++s8)
            if (rbase + s8 < Np)
              st_global_hint(dst + (int64_t)s8 * P.dp_rs, make_uint4(c[s8][0], c[s8][1], c[s8][2], c[s8][3]), L2_EVICT_FIRST);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  GVIT_SPAN(1);
  if (warp == 9) tmem_dealloc_2sm(tmem, 512);
}

inline size_t q_smem(int NT, int* nblk, int* slot_bytes) {
  *nblk = (2 * NT + 63) / 64;
  if (*nblk < 6) *nblk = 6;                                             // three 32 KB phase-A slots (and the extraction strips) live here
  *slot_bytes = NT * 128 < 24 * 1024 ? 24 * 1024 : NT * 128;           // four ring slots also hold three 32 KB phase-A slots
  return (size_t)*nblk * TILE + (size_t)Q_SLOTS * *slot_bytes + sizeof(QCtrl);
}

template <int KT>
int launch_pair(const CUtensorMap (&tm)[4], const QParams& P, size_t smem, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(graph_bwd_pair_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int pairs = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((num_sms() / 2) * 2);
    cfg.blockDim = dim3(Q_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&pairs, graph_bwd_pair_kernel<KT>, &cfg) != cudaSuccess || pairs < 1) {
      (void)cudaGetLastError();
      pairs = num_sms() / 2;
    }
  }
  // fewer images than pairs: two pairs per image when they fit (get_item splits the output chunks between them)
  const int used = P.B >= pairs ? pairs : (2 * P.B <= pairs && (P.D / 128) % 2 == 0 ? 2 * P.B : P.B);
  const int grid = 2 * used;
  graph_bwd_pair_kernel<KT><<<grid, Q_THREADS, smem, st>>>(tm[0], tm[1], tm[2], tm[3], P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_graph_bwd_pair)

bool graph_bwd_pair_supported(int Np, int D, int k) {
  int nblk, sb;
  const int NT = (Np + 15) & ~15;
  return Np > 128 && Np <= 256 && D >= 128 && D % 128 == 0 && D <= 1024 && k <= 16 && q_smem(NT, &nblk, &sb) <= 227 * 1024;
}

int graph_bwd_pair_tc(const Tokens& t, int k, const int32_t* idx, const float* vals, const float* w, const float* rnorm,
                      const void* dz, int64_t dz_batch_stride, float* dvals, void* dp, cudaStream_t st) {
  const int NT = (t.Np + 15) & ~15;
  CUtensorMap tm[4];
  int rc = make_tmap_bf16_3d(&tm[0], dz, t.D, t.Np, t.B, t.D, (uint64_t)dz_batch_stride, 128);                 // row-tile boxes of dZ
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm[1], t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, NT / 2);               // token halves of P
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm[2], dz, t.D, t.Np, t.B, t.D, (uint64_t)dz_batch_stride, NT);                    // whole 64-feature slabs
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm[3], t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, NT);
  if (rc != GVIT_OK) return rc;
  QParams P;
  P.B = t.B; P.Np = t.Np; P.D = t.D; P.k = k; P.NT = NT;
  size_t smem = q_smem(NT, &P.nblk, &P.slot_bytes);
  P.dsx = smem + (size_t)256 * k * 4 + 16 <= 227 * 1024;
  if (P.dsx) smem += (size_t)256 * k * 4 + 16;
  P.idx = idx; P.w = w; P.vals = vals; P.dvals = dvals; P.rnorm = rnorm;
  P.dp = static_cast<__nv_bfloat16*>(dp); P.dp_bs = t.batch_stride; P.dp_rs = t.row_stride;
  P.kmagic = (uint32_t)((0x100000000ull + (uint64_t)k - 1) / (uint64_t)k);
  P.kvec = (k == 4 || k == 8 || k == 16) && aligned16(idx) && aligned16(w) && aligned16(vals) && aligned16(dvals);
  if (k <= 4) return launch_pair<4>(tm, P, smem, st);
  if (k <= 8) return launch_pair<8>(tm, P, smem, st);
  return launch_pair<16>(tm, P, smem, st);
}

}  // namespace gvit

// knn_tc.cu - graph construction G1-G3 for bf16 tokens on tcgen05 (north_star:
// "a tiled similarity GEMM fused with a top-k select that emits the adjacency").
//
// Persistent CTAs, one image (Np <= 256 patch tokens) per iteration.  The Gram matrix G = P P^T is accumulated in TMEM by
// tcgen05.mma from TMA-staged 64-feature slabs of the image's tokens: the SAME shared-memory slab is the
// A operand (its rows [128*mt, 128*mt+128)) and the B operand (all rows), so each token is read from HBM
// exactly once (Np*D*2 bytes per image) and the Np x Np similarity matrix never leaves the SM.
// bf16 x bf16 products are exact in fp32 and accumulate in fp32, so G matches an fp32 evaluation of the
// same bf16 tokens to ~1e-6; the norms come from G's diagonal and S_ij = (G_ij * rn_i) * rn_j.
// Epilogue: thread <-> similarity row (TMEM lane); it streams its row out of TMEM 32 columns at a time and keeps a
// sorted top-k of 32-bit KEYS in registers: key = fixed-point similarity (24 bits: round(s * 2^22) + 2^22, s in [-1, 1])
// above (255 - column) in the low byte.  Keys are distinct, larger key == larger similarity or, at equal similarity,
// lower column, so inserting a candidate into the sorted list is one max/min pair per slot (2 integer instructions; the
// fp32 (value, index) insertion needed a compare, a predicate-or and four selects per slot and made the top-k 70 % of
// the kernel).  The 2^-22 quantum (2.4e-7) is below the ~1e-6 accumulation-order noise of the bf16 Gram matrix; the
// emitted similarities are the de-quantised keys.
//
// The TMA producer only depends on the slab ring, so while the eight epilogue warps select the neighbours of image n it
// already streams the first slabs of image n+1 (the tensor pipe has to wait for the accumulators: 2 x NT columns do not
// fit TMEM twice).  One CTA per image paid the launch / TMEM-allocation / cold-ring prologue 1.73 times per SM.
// Warp roles: 0-3 epilogue of rows 0-127, 4-7 epilogue of rows 128-255, 8 TMA producer, 9 MMA issuer.
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 256 * 128;   // 256 token rows x 64 bf16
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;           // two 128 x (<=256) fp32 accumulators
// similarities are quantised to round(s * KEY_SCALE) for the top-k keys: 2^22 - 8 leaves room for |s| up to 1 + 1.9e-6
// (rounding can push a cosine that far past 1) inside the 23-bit window of the float -> integer trick
constexpr float KEY_SCALE = 4194296.0f;

struct __align__(8) Ctrl {
  float rn[256];
  uint64_t full[STAGES], empty[STAGES], accum_full, accum_free;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(Ctrl);

// compare-exchange on keys, larger first
__device__ __forceinline__ void cex(uint32_t& a, uint32_t& b) {
  const uint32_t hi = max(a, b);
  b = min(a, b);
  a = hi;
}
// k = 8: eight candidates are sorted with the 19-exchange network, merged with the running list (element-wise max against
// the reversed candidates leaves the top 8 as a bitonic sequence) and re-sorted with the 12-exchange bitonic network:
// 70 max/min per 8 candidates against 128 for eight sorted insertions.  The integer pipe (half rate) is what the top-k
// is bound by, so the instruction count is the run time.
__device__ __forceinline__ void topk8_merge(uint32_t (&key)[8], uint32_t (&c)[8]) {
  cex(c[0], c[1]); cex(c[2], c[3]); cex(c[4], c[5]); cex(c[6], c[7]);
  cex(c[0], c[2]); cex(c[1], c[3]); cex(c[4], c[6]); cex(c[5], c[7]);
  cex(c[1], c[2]); cex(c[5], c[6]); cex(c[0], c[4]); cex(c[3], c[7]);
  cex(c[1], c[5]); cex(c[2], c[6]);
  cex(c[1], c[4]); cex(c[3], c[6]);
  cex(c[2], c[4]); cex(c[3], c[5]);
  cex(c[3], c[4]);
#pragma unroll
  for (int i = 0; i < 8; ++i) key[i] = max(key[i], c[7 - i]);
  cex(key[0], key[4]); cex(key[1], key[5]); cex(key[2], key[6]); cex(key[3], key[7]);
  cex(key[0], key[2]); cex(key[1], key[3]); cex(key[4], key[6]); cex(key[5], key[7]);
  cex(key[0], key[1]); cex(key[2], key[3]); cex(key[4], key[5]); cex(key[6], key[7]);
}

// One similarity row (this thread's TMEM lane, accumulator at `trow`) -> its KT largest keys, sorted.  rn = the column
// norms in shared memory; key = fixed-point similarity above (255 - column), see the file header.
// Blocks FIRST, FIRST + STEP, ... of 32 columns are scanned (two threads can share a row).  The block loop is unrolled so
// that (255 - column) is an immediate: the key is ONE integer multiply-add on the FMA pipe, which leaves the half-rate
// integer pipe - what this selection is bound by - to the max / min network alone (8.75 instead of 10.75 per candidate).
template <int KT, int FIRST = 0, int STEP = 1>
__device__ __forceinline__ void select_row(uint32_t trow, int Np, const float* rn, float rn_i, uint32_t (&key)[KT]) {
#pragma unroll
  for (int s = 0; s < KT; ++s) key[s] = 0u;               // below every real key
  const float ci = rn_i * KEY_SCALE;                      // fixed-point scale folded into the row factor
  // one 32-column block of this thread's similarity row; MASK only for the block that straddles Np
  auto block = [&](const int c0, auto mask_tag) {
    constexpr bool MASK = decltype(mask_tag)::value;
    float v[32];
    tmem_ld32(trow + c0, v);
#pragma unroll
    for (int t8 = 0; t8 < 32; t8 += 8) {
      if (MASK && c0 + t8 >= Np) break;                   // whole group beyond the last token (warp-uniform)
      const float4 rna = *reinterpret_cast<const float4*>(&rn[c0 + t8]);
      const float4 rnb = *reinterpret_cast<const float4*>(&rn[c0 + t8 + 4]);
      const float rnv[8] = {rna.x, rna.y, rna.z, rna.w, rnb.x, rnb.y, rnb.z, rnb.w};
      uint32_t cand[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        // round(s * KEY_SCALE) lands in the low mantissa bits of (x + 1.5 * 2^23); * 256 drops the exponent byte and
        // leaves 2^22 + round(s * KEY_SCALE) in bits 8..31; the addend is the column tag
        const float x = fmaf(v[t8 + u] * ci, rnv[u], 12582912.0f);
        asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(cand[u]) : "r"(__float_as_uint(x)), "r"(255u - (uint32_t)(c0 + t8 + u)));
        if (MASK && c0 + t8 + u >= Np) cand[u] = 0u;      // columns >= Np can never be selected
      }
      if constexpr (KT == 8) {
        topk8_merge(key, cand);
      } else {
        // sorted insertion: the candidate sinks through the list, every slot keeps the larger key
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t kx = cand[u];
#pragma unroll
          for (int s = 0; s < KT; ++s) cex(key[s], kx);
        }
      }
    }
  };
#pragma unroll
  for (int bi = FIRST; bi < 8; bi += STEP) {              // Np <= 256: at most eight blocks
    const int c0 = 32 * bi;
    if (c0 >= Np) break;
    if (c0 + 32 <= Np) block(c0, std::false_type{});
    else block(c0, std::true_type{});
  }
}

// this thread's neighbours out of its sorted keys.  k == KT (4, 8, 16, 32: the shipped configurations) writes each row's
// k indices and k similarities as 16-byte vectors: sixteen scalar stores per thread kept the selection warps ~3k cycles per
// image in the store queue (profiles/r4g_trace_knn_pair_v3.txt).
template <int KT>
__device__ __forceinline__ void emit_row(const uint32_t (&key)[KT], int k, int32_t* idx, float* vals, int64_t o) {
  if (k == KT && ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(vals)) & 15u) == 0) {   // o = row * k: a multiple of 4
#pragma unroll
    for (int s = 0; s < KT; s += 4) {
      int4 i4;
      float4 v4;
      i4.x = 255 - (int)(key[s] & 0xffu);     v4.x = (float)((int)(key[s] >> 8) - 4194304) * (1.0f / KEY_SCALE);
      i4.y = 255 - (int)(key[s + 1] & 0xffu); v4.y = (float)((int)(key[s + 1] >> 8) - 4194304) * (1.0f / KEY_SCALE);
      i4.z = 255 - (int)(key[s + 2] & 0xffu); v4.z = (float)((int)(key[s + 2] >> 8) - 4194304) * (1.0f / KEY_SCALE);
      i4.w = 255 - (int)(key[s + 3] & 0xffu); v4.w = (float)((int)(key[s + 3] >> 8) - 4194304) * (1.0f / KEY_SCALE);
      *reinterpret_cast<int4*>(idx + o + s) = i4;
      *reinterpret_cast<float4*>(vals + o + s) = v4;
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < KT; ++s)
    if (s < k) {
      idx[o + s] = 255 - (int)(key[s] & 0xffu);
      vals[o + s] = (float)((int)(key[s] >> 8) - 4194304) * (1.0f / KEY_SCALE);
    }
}

template <int KT>
__global__ void __launch_bounds__(THREADS, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmap, int B, int Np, int D,
                                                            int k, int NT, int32_t* __restrict__ idx,
                                                            float* __restrict__ vals, float* __restrict__ rnorm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // no static __shared__ in these kernels: base is 1024-aligned
  uint8_t* stages = smem_raw;   // NOT rounded through an integer: that made every access a generic LD/ST instead of LDS/STS
  if ((smem_u32(stages) & 1023u) != 0) __trap();
  Ctrl* ctl = reinterpret_cast<Ctrl*>(stages + (size_t)STAGES * STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role id
  const int slabs = D / 64;
  const int mtiles = Np > 128 ? 2 : 1;
  GVIT_TRACE_DECL

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->accum_full, 1);
    mbar_init(&ctl->accum_free, 256);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      int it = 0;                                          // ring fill counter, runs across images
      for (int b = blockIdx.x; b < B; b += gridDim.x)
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const int s = it % STAGES;
          mbar_wait(&ctl->empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&ctl->full[s], (uint32_t)NT * 128u);
          tma_load_3d(stages + (size_t)s * STAGE_BYTES, &tmap, sl * 64, 0, b, &ctl->full[s]);
        }
    }
  } else if (warp == 9) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      const uint32_t idesc = make_idesc(128, NT, false, false);
      int it = 0, n = 0;
      for (int b = blockIdx.x; b < B; b += gridDim.x, ++n) {
        if (n > 0) {                    // the epilogue warps have read every accumulator column of the previous image
          mbar_wait(&ctl->accum_free, (n - 1) & 1);
          tc_fence_after();
        }
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const int s = it % STAGES;
          mbar_wait(&ctl->full[s], (it / STAGES) & 1);
          tc_fence_after();
          GVIT_TR(1);
          const uint32_t base = smem_u32(stages + (size_t)s * STAGE_BYTES);
          for (int mt = 0; mt < mtiles; ++mt)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + mt * 256, make_sdesc(base + mt * (128 * 128) + kk * 32), make_sdesc(base + kk * 32), idesc,
                      sl > 0 || kk > 0);
          umma_commit(&ctl->empty[s]);    // slab consumed -> producer may refill it
        }
        umma_commit(&ctl->accum_full);    // all MMAs retired -> accumulators readable
        GVIT_TR(2);
      }
    }
  } else {
    const int g = warp >> 2;                              // accumulator tile
    const int wrow0 = g * 128 + (warp & 3) * 32;          // first similarity row of this warp
    const int row = wrow0 + lane;
    const bool active = g < mtiles && wrow0 < Np;         // warp-uniform
    const uint32_t trow = tmem_lane_base(tmem, warp) + g * 256;
    int n = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++n) {
    float rn_i = 0.f;
    if (active) {
      GVIT_TR(10);
      mbar_wait(&ctl->accum_full, n & 1);
      tc_fence_after();
      GVIT_TR(11);
      float v[32];
      tmem_ld32(trow + wrow0, v);                         // the 32x32 block on the diagonal
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) d = (i == lane) ? v[i] : d;
      rn_i = 1.0f / fmaxf(sqrtf(d), 1e-12f);
      if (row < Np) { ctl->rn[row] = rn_i; rnorm[(int64_t)b * Np + row] = rn_i; }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 epilogue warps: norms visible
    GVIT_TR(12);
    if (active) {
      uint32_t key[KT];
      select_row<KT>(trow, Np, ctl->rn, rn_i, key);
      GVIT_TR(13);
      tc_fence_before();
      mbar_arrive(&ctl->accum_free);                        // this thread's accumulator row is in registers (keys)
      if (row < Np) emit_row<KT>(key, k, idx, vals, ((int64_t)b * Np + row) * k);
    } else {
      mbar_arrive(&ctl->accum_free);                        // idle warps keep the arrival count fixed at 256
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");          // rn[] of this image is dead before the next one overwrites it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// knn_pair_kernel (128 < Np <= 256): ONE CTA PAIR PER IMAGE, `tcgen05.mma.cta_group::2`.
//
// The trace of knn_tc_kernel (profiles/r4g_trace_knn_two_pass.txt) shows what bounds the Gram GEMM: shared-memory
// bandwidth.  An un-paired CTA issues, per 64-feature slab, eight 128 x NT x 16 MMAs that each read 4 KB of A and 6.6 KB of
// B from shared memory while TMA writes the 26.6 KB slab: 111 KB per slab at 128 B/clk = 870 cycles (measured 860) against
// 832 of tensor time - and then the tensor pipe idles while the eight warps select (2 x NT accumulator columns do not fit
// TMEM twice).  Splitting the passes per row tile so that GEMM and selection overlap re-reads the slabs and ended 10 % faster
// only (0.0352 -> 0.0317 ms, same trace file).
// As a pair, the two row tiles of an image are the two halves of ONE M = 256 MMA: rank r stages its 128 rows (A) and its
// half of the tokens (B: rows [r NT/2, (r + 1) NT/2)), so an MMA reads 7.3 KB per CTA and a slab costs 29 KB of fill: ~456
// shared-memory cycles against 416 tensor cycles per slab.  Each CTA's accumulator is 128 lanes x NT columns, so TWO fit
// its TMEM: the Gram GEMM of image n + 1 runs while the eight selection warps take the neighbours of image n out of the
// other buffer (TWO threads per row, each scanning every other 32-column block and keeping its own sorted list; the odd
// thread hands its keys over through shared memory and the even one merges: one warpgroup per image had the same ALU
// work but twice the buffer turnaround, and the tensor pipe waited for it, profiles/r4g_trace_knn_pair_v2.txt).
// Column norms: every CTA finds the norms of its own rows on its accumulator's diagonal and sends them to the peer with
// st.async (remote store + complete_tx on the peer's mbarrier: 512 bytes per image and direction, no cluster-scope fence;
// a release.cluster arrival per thread cost 3-6k cycles per image, profiles/r4g_trace_knn_pair_v1.txt).
// Roles per CTA: warps 0-3 / 4-7 selection (even / odd column blocks of the same 128 rows), 8 TMA producer, 9 MMA issuer
// (leader CTA only) and TMEM owner.
constexpr int P_STAGES = 5;
constexpr int P_STAGE_BYTES = 2 * 128 * 128;   // A: this CTA's 128 token rows x 64 features | B: its <= 128 tokens of the N extent

struct __align__(8) CtrlP {
  float rn[2][256];                        // column norms of the image in flight, per accumulator buffer
  uint32_t xch[GVIT_MAX_K][128];           // the odd-block thread of a row hands its keys to the even-block thread
  uint64_t full[P_STAGES], empty[P_STAGES], acc_full[2], acc_free[2], norm_bar[2];
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES_P = (size_t)P_STAGES * P_STAGE_BYTES + sizeof(CtrlP);

// one fp32 into the peer CTA's shared memory; its arrival is counted (4 bytes) on the mbarrier `mbar_cluster` of that CTA, and
// a thread that sees the barrier's phase complete sees the value: no fence on either side
__device__ __forceinline__ void st_async_f32(uint32_t dst_cluster, float v, uint32_t mbar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(dst_cluster), "r"(__float_as_uint(v)), "r"(mbar_cluster) : "memory");
}

template <int KT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) knn_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                        const __grid_constant__ CUtensorMap tmB,
                                                                                        int B, int Np, int D, int k, int NT,
                                                                                        int32_t* __restrict__ idx, float* __restrict__ vals,
                                                                                        float* __restrict__ rnorm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stages = smem_raw;
  if ((smem_u32(stages) & 1023u) != 0) __trap();
  CtrlP* ctl = reinterpret_cast<CtrlP*>(stages + (size_t)P_STAGES * P_STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int slabs = D / 64, NH = NT / 2;                   // NH tokens of the N extent per CTA
  GVIT_TRACE_DECL
  GVIT_SPAN(0);
  GVIT_TR(30);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_free[s], 16);                    // leader's: the 8 selection warps of each CTA
      mbar_init(&ctl->norm_bar[s], 1);                     // one expect_tx arrival per image; the peer's 128 norms are the bytes
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2sm(&ctl->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
  GVIT_TR(33);

  if (warp == 8) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int b = cid; b < B; b += ncl)
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const uint32_t s = it % P_STAGES;
          mbar_wait(&ctl->empty[s], ((it / P_STAGES) & 1) ^ 1);
          const uint32_t fullL = mapa_u32(smem_u32(&ctl->full[s]), 0);
          if (rank == 0) mbar_expect_tx(&ctl->full[s], (uint32_t)(2 * (128 + NH) * 128));
          uint8_t* st = stages + (size_t)s * P_STAGE_BYTES;
          tma_load_3d_2sm(st, &tmA, sl * 64, rank * 128, b, fullL);              // this CTA's row tile (rows >= Np: zeros)
          tma_load_3d_2sm(st + 128 * 128, &tmB, sl * 64, rank * NH, b, fullL);   // its tokens of the N extent
        }
    }
  } else if (warp == 9) {
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(256, NT, false, false);
      uint32_t it = 0, n = 0;
      for (int b = cid; b < B; b += ncl, ++n) {
        const uint32_t buf = n & 1;
        mbar_wait(&ctl->acc_free[buf], ((n >> 1) & 1) ^ 1);                       // both CTAs selected image n - 2 out of it
        tc_fence_after();
        GVIT_TR(1);
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const uint32_t s = it % P_STAGES;
          mbar_wait(&ctl->full[s], (it / P_STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(stages + (size_t)s * P_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_2sm(tmem + buf * 256, make_sdesc(base + kk * 32), make_sdesc(base + 128 * 128 + kk * 32), idesc, sl > 0 || kk > 0);
          umma_commit_2sm_mc(&ctl->empty[s], 3);
        }
        umma_commit_2sm_mc(&ctl->acc_full[buf], 3);
        GVIT_TR(2);
      }
    }
  } else {
    const int p = warp >> 2;                              // 0: even 32-column blocks of the row, 1: odd blocks
    const int wt0 = (warp & 3) * 32;                      // first tile row of this warp
    const int r = wt0 + lane;                             // tile row == TMEM lane
    const int row = rank * 128 + r;                       // token row
    const bool active = rank * 128 + wt0 < Np;            // warp-uniform
    const uint32_t freeL0 = mapa_u32(smem_u32(&ctl->acc_free[0]), 0), freeL1 = mapa_u32(smem_u32(&ctl->acc_free[1]), 0);
    uint32_t n = 0;
    for (int b = cid; b < B; b += ncl, ++n) {
      const uint32_t buf = n & 1, ph = (n >> 1) & 1;
      const uint32_t tl = tmem_lane_base(tmem, warp) + buf * 256;
      float* rn = ctl->rn[buf];
      GVIT_TR(10);
      mbar_wait(&ctl->acc_full[buf], ph);
      tc_fence_after();
      GVIT_TR(11);
      if (p == 0) {
        if (warp == 0 && lane == 0) mbar_expect_tx(&ctl->norm_bar[buf], 128u * 4u);   // the peer's 128 norms of this image
        float rn_i = 0.f;
        if (active) {
          float v[32];
          tmem_ld32(tl + rank * 128 + wt0, v);            // the 32x32 block on the diagonal
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) d = (i == lane) ? v[i] : d;
          rn_i = row < Np ? 1.0f / fmaxf(sqrtf(d), 1e-12f) : 0.f;
          if (row < Np) rnorm[(int64_t)b * Np + row] = rn_i;
        }
        rn[row] = rn_i;                                   // rows >= Np: 0 (their columns are never selected)
        st_async_f32(mapa_u32(smem_u32(&rn[row]), rank ^ 1), rn_i, mapa_u32(smem_u32(&ctl->norm_bar[buf]), rank ^ 1));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // this CTA's 128 norms are in its shared memory
      mbar_wait(&ctl->norm_bar[buf], ph);                 // ... and the peer's 128 have landed
      GVIT_TR(12);
      uint32_t key[KT];
      if (active) {
        if (p == 0) select_row<KT, 0, 2>(tl, Np, rn, rn[row], key);
        else select_row<KT, 1, 2>(tl, Np, rn, rn[row], key);
        if (p == 1) {
#pragma unroll
          for (int s2 = 0; s2 < KT; ++s2) ctl->xch[s2][r] = key[s2];
        }
      }
      GVIT_TR(13);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(buf ? freeL1 : freeL0);   // this warp's accumulator rows are in registers (keys)
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the odd-block keys are in shared memory
      if (p == 0 && active && row < Np) {
        uint32_t other[KT];
#pragma unroll
        for (int s2 = 0; s2 < KT; ++s2) other[s2] = ctl->xch[s2][r];
        if constexpr (KT == 8) {
          topk8_merge(key, other);
        } else {
#pragma unroll
          for (int u = 0; u < KT; ++u) {
            uint32_t kx = other[u];
#pragma unroll
            for (int s2 = 0; s2 < KT; ++s2) cex(key[s2], kx);
          }
        }
        emit_row<KT>(key, k, idx, vals, ((int64_t)b * Np + row) * k);
      }
    }
  }
  GVIT_TR(31);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  GVIT_TR(32);
  GVIT_SPAN(1);
  if (warp == 9) tmem_dealloc_2sm(tmem, TMEM_COLS);
}

template <int KT>
int launch_pair(const Tokens& t, int k, int NT, int32_t* idx, float* vals, float* rnorm, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_3d(&tmA, t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tmB, t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, NT / 2);
  if (rc != GVIT_OK) return rc;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(knn_pair_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_P));
  int pairs = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((num_sms() / 2) * 2);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES_P;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&pairs, knn_pair_kernel<KT>, &cfg) != cudaSuccess || pairs < 1) {
      (void)cudaGetLastError();
      pairs = num_sms() / 2;
    }
  }
  const int grid = 2 * (t.B < pairs ? t.B : pairs);
  knn_pair_kernel<KT><<<grid, THREADS, SMEM_BYTES_P, st>>>(tmA, tmB, t.B, t.Np, t.D, k, NT, idx, vals, rnorm);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

template <int KT>
int launch(const CUtensorMap& tmap, const Tokens& t, int k, int NT, int32_t* idx, float* vals, float* rnorm, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int grid = t.B < num_sms() ? t.B : num_sms();
  knn_tc_kernel<KT><<<grid, THREADS, SMEM_BYTES, st>>>(tmap, t.B, t.Np, t.D, k, NT, idx, vals, rnorm);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_knn)

bool knn_tc_supported(int Np, int D, int k) { return Np >= 16 && Np <= 256 && D >= 64 && D % 64 == 0 && k <= 32; }

int knn_fwd_tc(const Tokens& t, int k, int32_t* idx, float* vals, float* rnorm, cudaStream_t st) {
  const int NT = (t.Np + 15) & ~15;
  CUtensorMap tmap;
  int rc = make_tmap_bf16_3d(&tmap, t.ptr, t.D, t.Np, t.B, t.row_stride, t.batch_stride, NT);
  if (rc != GVIT_OK) return rc;
  // 128 < Np <= 256: one CTA pair per image (cta_group::2); GVIT_KNN_NOPAIR=1 keeps the one-CTA kernel (A/B switch)
  static const bool pair = getenv("GVIT_KNN_NOPAIR") == nullptr;
  if (pair && t.Np > 128) {
    if (k <= 4) return launch_pair<4>(t, k, NT, idx, vals, rnorm, st);
    if (k <= 8) return launch_pair<8>(t, k, NT, idx, vals, rnorm, st);
    if (k <= 16) return launch_pair<16>(t, k, NT, idx, vals, rnorm, st);
    return launch_pair<32>(t, k, NT, idx, vals, rnorm, st);
  }
  if (k <= 4) return launch<4>(tmap, t, k, NT, idx, vals, rnorm, st);
  if (k <= 8) return launch<8>(tmap, t, k, NT, idx, vals, rnorm, st);
  if (k <= 16) return launch<16>(tmap, t, k, NT, idx, vals, rnorm, st);
  return launch<32>(tmap, t, k, NT, idx, vals, rnorm, st);
}

}  // namespace gvit

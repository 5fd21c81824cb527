// agg4_tc.cu - fused aggregation G4+G5+G6 + residual for bf16, 128 < Np <= 256, D <= 768, as a 2-SM kernel: ONE CTA PAIR PER
// IMAGE (north_star: "gather / normalise, then a tcgen05 GEMM with TMA-staged tiles that keeps A.X in
// shared memory or TMEM with no HBM round-trip").
//
//   out[b,1+i,:] = resid[b,1+i,:] + ( sum_j softmax_k(vals)_ij * p[b, idx_ij, :] ) Wg^T + bias
//
// agg3_tc.cu gives each 128-row tile of an image its own CTA, so both CTAs of an image stream ALL token slabs (301 KB) and
// ALL of Wg (1.18 MB) through L2 -> shared memory.  Here the two row tiles of an image are the two halves of `cta_group::2`
// MMAs (M = 256): every B operand - token slab or W piece - is split between the pair, each CTA stages HALF of it, so the
// operand bytes per CTA and per image halve (0.74 MB).  The pair is persistent over its work items (get_item below: whole
// images in full rounds; the images left over after the last full round are split between two pairs, half of the output
// chunks each).
//
// Per item (rank r = 0 / 1 owns token rows [128 r, 128 r + 128) = its TMEM lanes, its A~ tile, its output rows):
//   A~ tile    : 128 x NT dense bf16 adjacency rows in shared memory (zeroed, then each row's k softmax weights scattered);
//                the tile of item n + 1 is built while item n's projection runs (the row warps would otherwise wait there).
//   Z phase    : Z = A~ . P, 128 features per step: MMA M = 256, N = 128, K = tokens; rank r stages the 64-feature token
//                slab 2 t + r (MN-major B) as two half-slabs of R0 and NT - R0 tokens, one 16 KB ring slot each.  fp32 result in a TMEM staging area, converted IN TENSOR MEMORY to packed bf16
//                by BOTH row warpgroups (64 staging columns each), so a step's conversion overlaps the next step's MMAs:
//                the whole aggregated tile Z [128 x D] of a CTA ends up in TMEM columns [0, D/2) (training also sends each
//                [32 rows x 64 features] piece out through a TMA store for the weight gradient).
//   projection : per 128-feature output chunk OUT = Z . W_chunk^T, A from TMEM (TS form), N = 128 per instruction (a 2-SM
//                MMA takes ~75-80 cycles whatever its N: 64-wide chunks ran the tensor pipe at 40 %); rank r stages W rows
//                [128 n + 64 r, + 64) in pieces of 128 reduction columns (16 KB).  ONE fp32 chunk buffer: TMEM is full, so the
//                MMAs of chunk n + 1 wait until both warpgroups have read chunk n (~1.8k of the ~5.3k cycles per chunk: the
//                bound of this design, see profiles/README.md).  Epilogue: warpgroup g takes the 64-feature half g of the
//                chunk in units of [32 rows x 128 bytes] (one per chunk on a bf16 stream, two on the fp32 stream); every row
//                warp owns two staging tiles used in turn: the unit's residual tile lands by TMA load while the previous unit
//                is worked on, bias + branch value are added in place, one TMA store sends the tile out.
// TMEM (512 columns per CTA): Z bf16 [0, D/2) | Z fp32 staging: step t even -> [64 t, 64 t + 128) (in place), odd ->
//                [384, 512) | OUT chunk buffer [384, 512).
// Shared memory per CTA: A~ (4 x 16 KB) | ring of 6 x 16 KB slots | 8 warps x 2 x 4 KB staging tiles.
// Warp roles: 0-3 row warpgroup 0, 4-7 row warpgroup 1 (thread <-> token row = TMEM lane), 8 TMA producer (both CTAs),
// 9 MMA issuer (leader) and TMEM owner.
// Barriers.  In the leader, arrived at by both CTAs: full[slot] (TMA bytes), a_ready, conv_done[step parity], out_free.
// In each CTA, released by multicast tcgen05.commit: empty[slot], a_free, zs_full[step parity], out_full.  The two warps
// that share a TMEM lane quarter meet at a 64-thread named barrier inside every in-place conversion.  Every waiter counts
// the completions it has consumed and waits for them one by one, so no barrier can run two phases ahead of a waiter.
// Measured at B = 256, Np = 196, D = 768, k = 8 (training: w and Z saved), same box: fp32 residual stream 0.1037 ms against
// 0.1401 ms for agg3; bf16 stream 0.089-0.091 ms against 0.1210 ms.
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int THREADS = 320;
constexpr int TILE = 128 * 128;             // [128 rows][64 bf16]
constexpr int A_BYTES = 4 * TILE;           // A~ [128][256] as four 64-column blocks
constexpr int SLOT = 16 * 1024;             // ring slot: HALF a token slab ([<=128 tokens][64]) or one W piece (2 boxes of [64][64])
constexpr int NSLOT = 6;
constexpr int WSTAGE = 4 * 1024;            // staging tile ([32 rows][128 B]): residual in (TMA load), result out (TMA store)
constexpr int NSTG = 2;                     // two staging tiles per row warp: one loading / storing while the other is worked on
constexpr int T_OUT = 384;                  // first OUT chunk buffer / odd Z staging

struct __align__(8) Ctrl {
  uint64_t full[NSLOT], empty[NSLOT], a_ready, a_free, zs_full[2], conv_done[2], out_full, out_free;
  uint64_t rbar[8][NSTG];                   // per row warp and staging tile: the residual tile has landed
  uint32_t tmem_base;
  __align__(16) __nv_bfloat16 bias[768];
};
constexpr size_t SMEM_BYTES = A_BYTES + NSLOT * SLOT + 8 * NSTG * WSTAGE + sizeof(Ctrl);

struct Params {
  int B, Np, D, k, NT;
  int R0;                                   // tokens in the first half-slab: a multiple of 16, the second holds NT - R0 (<= R0)
  int kvec;                                 // k == KT and idx / vals / w_save 16-byte aligned: whole adjacency rows by vector access
  const int32_t* idx;
  const float* vals;
  const __nv_bfloat16* bias;
  const void* resid;                        // bf16, or fp32 when the kernel is instantiated with RES32 (fp32 residual stream)
  void* out;                                // same type as resid
  float* w_save;
  __nv_bfloat16* z_save;
  int64_t zbs;                              // batch stride of z_save in elements (rows are D apart)
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ int wrow0_of(int rank, int warp) { return rank * 128 + (warp & 3) * 32; }

// consume completions of `bar` until `seen` reaches `need` (one parity wait per completion: never skips a phase)
__device__ __forceinline__ void wait_upto(uint64_t* bar, uint32_t& seen, uint32_t need) {
  while (seen < need) { mbar_wait(bar, seen & 1); ++seen; }
}

template <int KT, bool RES32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) agg4_tc_kernel(const __grid_constant__ CUtensorMap tm_tok,
                                                                                       const __grid_constant__ CUtensorMap tm_w,
                                                                                       const __grid_constant__ CUtensorMap tm_z,
                                                                                       const __grid_constant__ CUtensorMap tm_out,
                                                                                       const __grid_constant__ CUtensorMap tm_res, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;
  if ((smem_u32(sA) & 1023u) != 0) __trap();
  uint8_t* sRing = sA + A_BYTES;
  uint8_t* sStg = sRing + NSLOT * SLOT;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(sStg + 8 * NSTG * WSTAGE);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  GVIT_TRACE_DECL
  GVIT_SPAN(0);
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int D = P.D, NT = P.NT;
  const int nchunk = D / 128;               // 128-feature output chunks (64 W rows per CTA)
  const int nstep = D / 128;                // Z steps of 128 features (64 per CTA)
  const int npiece = D / 128;               // W pieces of 128 reduction columns per output chunk: 2 boxes of [64][64]

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_tok);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_z);
    prefetch_tmap(&tm_out);
    prefetch_tmap(&tm_res);
    for (int w8 = 0; w8 < 8; ++w8)
      for (int s = 0; s < NSTG; ++s) mbar_init(&ctl->rbar[w8][s], 1);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->zs_full[s], 1);
      mbar_init(&ctl->conv_done[s], 16);    // the 8 row warps of each CTA
    }
    mbar_init(&ctl->out_full, 1);
    mbar_init(&ctl->out_free, 16);          // leader's: the 8 row warps of each CTA have read the chunk out of TMEM
    mbar_init(&ctl->a_ready, 8);            // warpgroup 0 of each CTA
    mbar_init(&ctl->a_free, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2sm(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t c = 0;                                                  // ring fill counter
      Item I;
      for (int it = 0; get_item(it, cid, ncl, P.B, nchunk, I); ++it) {
        const int b = I.b;
        for (int t = 0; t < nstep; ++t) {                              // this CTA's 64-feature token slab of step t, as two
          for (int hf = 0; hf < 2; ++hf, ++c) {                        // half-slabs of R0 tokens (rows >= Np arrive as zeros)
            const uint32_t sl = c % NSLOT;
            mbar_wait(&ctl->empty[sl], ((c / NSLOT) & 1) ^ 1);
            if (rank == 0) mbar_expect_tx(&ctl->full[sl], (uint32_t)(2 * P.R0 * 128));
            tma_load_3d_2sm(sRing + sl * SLOT, &tm_tok, (2 * t + rank) * 64, hf * P.R0, b, mapa_u32(smem_u32(&ctl->full[sl]), 0));
          }
        }
        for (int n = I.c0; n < I.c1; ++n) {                            // W rows [128 n + 64 rank, + 64), pieces of 128 columns
          for (int p = 0; p < npiece; ++p, ++c) {
            const uint32_t sl = c % NSLOT;
            mbar_wait(&ctl->empty[sl], ((c / NSLOT) & 1) ^ 1);
            if (rank == 0) mbar_expect_tx(&ctl->full[sl], (uint32_t)(2 * 2 * 8192));
            const uint32_t fullL = mapa_u32(smem_u32(&ctl->full[sl]), 0);
            for (int j = 0; j < 2; ++j)
              tma_load_3d_2sm(sRing + sl * SLOT + j * 8192, &tm_w, p * 128 + j * 64, n * 128 + rank * 64, 0, fullL);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t aA = smem_u32(sA), aR = smem_u32(sRing);
      const uint32_t idesc_z = make_idesc(256, 128, false, true);      // A~ K-major, token slabs MN-major
      const uint32_t idesc_w = make_idesc(256, 128, false, false);     // Z from TMEM, W piece K-major (64 rows per CTA)
      uint32_t c = 0;                                                  // ring consume counter
      uint32_t a_seen = 0, conv_seen0 = 0, conv_seen1 = 0, conv_iss0 = 0, conv_iss1 = 0, free_seen = 0, out_iss = 0;
      Item I;
      for (int it = 0; get_item(it, cid, ncl, P.B, nchunk, I); ++it) {
        wait_upto(&ctl->a_ready, a_seen, a_seen + 1);                  // both CTAs' adjacency tiles are built
        tc_fence_after();
        GVIT_TR(1);
        for (int t = 0; t < nstep; ++t) {                              // ---- Z phase
          const int par = t & 1;
          const uint32_t stg = par ? T_OUT : 64 * t;
          if (par) {                                                   // earlier steps of this parity converted; odd staging =
            wait_upto(&ctl->conv_done[1], conv_seen1, conv_iss1);      // the OUT buffer of the previous image
            wait_upto(&ctl->out_free, free_seen, out_iss);
          } else {
            wait_upto(&ctl->conv_done[0], conv_seen0, conv_iss0);
          }
          const uint32_t sl0 = c % NSLOT, sl1 = (c + 1) % NSLOT;
          GVIT_TR(2);
          mbar_wait(&ctl->full[sl0], (c / NSLOT) & 1);
          mbar_wait(&ctl->full[sl1], ((c + 1) / NSLOT) & 1);
          tc_fence_after();
          GVIT_TR(3);
          const uint32_t aTok0 = aR + sl0 * SLOT, aTok1 = aR + sl1 * SLOT;
          const int K0 = P.R0 / 16;
          for (int ks = 0; ks < NT / 16; ++ks)
            umma_ss_2sm(tmem + stg, make_sdesc(aA + (ks >> 2) * TILE + (ks & 3) * 32),
                        make_sdesc(ks < K0 ? aTok0 + ks * 2048 : aTok1 + (ks - K0) * 2048), idesc_z, ks > 0);
          umma_commit_2sm_mc(&ctl->zs_full[par], 3);
          umma_commit_2sm_mc(&ctl->empty[sl0], 3);
          umma_commit_2sm_mc(&ctl->empty[sl1], 3);
          if (par) ++conv_iss1; else ++conv_iss0;
          c += 2;
        }
        umma_commit_2sm_mc(&ctl->a_free, 3);                           // the A~ tiles may be rebuilt for the next image
        wait_upto(&ctl->conv_done[0], conv_seen0, conv_iss0);          // every Z step is packed bf16 in TMEM
        wait_upto(&ctl->conv_done[1], conv_seen1, conv_iss1);
        tc_fence_after();
        GVIT_TR(4);
        for (int n = I.c0; n < I.c1; ++n) {                            // ---- projection: N = 128 per instruction (a 2-SM MMA
          wait_upto(&ctl->out_free, free_seen, out_iss);               // takes ~80 cycles whatever its N: 64-wide chunks ran at 40 %)
          tc_fence_after();
          GVIT_TR(5);
          for (int p = 0; p < npiece; ++p, ++c) {
            const uint32_t sl = c % NSLOT;
            mbar_wait(&ctl->full[sl], (c / NSLOT) & 1);
            tc_fence_after();
            GVIT_TR(6);
            const uint32_t aW = aR + sl * SLOT;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ts_2sm(tmem + T_OUT, tmem + (p * 128 + j * 64 + kk * 16) / 2, make_sdesc(aW + j * 8192 + kk * 32), idesc_w,
                            p > 0 || j > 0 || kk > 0);
            umma_commit_2sm_mc(&ctl->empty[sl], 3);
          }
          umma_commit_2sm_mc(&ctl->out_full, 3);
          GVIT_TR(7);
          ++out_iss;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ row warpgroups (both CTAs)
    const int g = warp >> 2;                                           // warpgroup: parity of the steps / chunks it owns
    const int row = threadIdx.x & 127;                                 // == TMEM lane
    const int rowg = rank * 128 + row;                                 // token row inside the image
    const bool valid = rowg < P.Np;
    uint8_t* stg0 = sStg + warp * (NSTG * WSTAGE);                     // this warp's two private staging tiles (4 KB each)
    const bool wact = wrow0_of(rank, warp) < P.Np;                     // warp-uniform: the warp owns at least one token row
    uint32_t rph0 = 0, rph1 = 0;                                       // residual-tile arrivals consumed per staging tile
    const int wrow0 = rank * 128 + (warp & 3) * 32;                    // first token row of this warp
    const uint32_t tl = tmem_lane_base(tmem, warp);
    const uint32_t a_readyL = mapa_u32(smem_u32(&ctl->a_ready), 0);
    const uint32_t conv_doneL0 = mapa_u32(smem_u32(&ctl->conv_done[0]), 0), conv_doneL1 = mapa_u32(smem_u32(&ctl->conv_done[1]), 0);   // in the leader
    const uint32_t out_freeL = mapa_u32(smem_u32(&ctl->out_free), 0);
    uint32_t afree_seen = 0, zs_seen0 = 0, zs_seen1 = 0, full_seen = 0;        // completions consumed of a_free / zs_full[parity] / out_full
    {
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < 768 / 8; i += 256)
        reinterpret_cast<uint4*>(ctl->bias)[i] = (P.bias && i < D / 8) ? reinterpret_cast<const uint4*>(P.bias)[i] : z4;
    }
    // ---- G4 + adjacency tile of image bb (the it-th of this pair): zero A~, then scatter each row's k softmax weights (bf16)
    //      at its neighbour columns.  Runs in the slot where the row warps would otherwise wait for the first projection
    //      chunk of the PREVIOUS image, so the MMA issuer finds a_ready complete when it gets to the next image.
    auto build_adj = [&](int bb, int it, bool primary) {
      GVIT_TR(10);
      if (it > 0) wait_upto(&ctl->a_free, afree_seen, (uint32_t)it);     // the previous image's Z MMAs have read A~
      GVIT_TR(11);
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sA)[i] = z4;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      GVIT_TR(20);
      if (g == 0) {
        if (valid) {
          float w[KT];
          int nb[KT];
          float mx = -FLT_MAX, sum = 0.f;
          const int64_t o = ((int64_t)bb * P.Np + rowg) * P.k;
          if (P.kvec) {                                                  // whole rows of 16 / 32 / 64 bytes: vector loads
#pragma unroll
            for (int j = 0; j < KT; j += 4) {
              const int4 i4 = *reinterpret_cast<const int4*>(P.idx + o + j);
              const float4 v4 = *reinterpret_cast<const float4*>(P.vals + o + j);
              nb[j] = i4.x; nb[j + 1] = i4.y; nb[j + 2] = i4.z; nb[j + 3] = i4.w;
              w[j] = v4.x; w[j + 1] = v4.y; w[j + 2] = v4.z; w[j + 3] = v4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
              const bool on = j < P.k;
              nb[j] = on ? P.idx[o + j] : 0;
              w[j] = on ? P.vals[o + j] : -FLT_MAX;
            }
          }
#pragma unroll
          for (int j = 0; j < KT; ++j) mx = fmaxf(mx, w[j]);
          GVIT_TR(21);
#pragma unroll
          for (int j = 0; j < KT; ++j) { w[j] = j < P.k ? expf(w[j] - mx) : 0.f; sum += w[j]; }
          const float inv = 1.0f / sum;
#pragma unroll
          for (int j = 0; j < KT; ++j) {
            if (j < P.k) {
              const float wj = w[j] * inv;
              w[j] = wj;
              if (P.w_save && primary && !P.kvec) P.w_save[o + j] = wj;
              const int cidx = nb[j];
              if (nb_ok(cidx, P.Np))
                *reinterpret_cast<__nv_bfloat16*>(sA + (cidx >> 6) * TILE + swz128(row, cidx & 63) + (cidx & 7) * 2) = __float2bfloat16_rn(wj);
            }
          }
          if (P.w_save && primary && P.kvec) {
#pragma unroll
            for (int j = 0; j < KT; j += 4) *reinterpret_cast<float4*>(P.w_save + o + j) = make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_readyL);
        GVIT_TR(12);
      } else if (rank == 0 && primary) {
        // CLS row: out[bb,0,:] = resid[bb,0,:] (the graph leaves CLS untouched, section 9 G0), 16 bytes per thread and trip
        constexpr int EPV = RES32 ? 4 : 8;                               // elements per 16-byte vector
        constexpr int ESZ = RES32 ? 4 : 2;
        for (int c = row; c < D / EPV; c += 128) {
          const int64_t o = ((int64_t)bb * (P.Np + 1) * D + c * EPV) * ESZ;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (P.resid) v = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(P.resid) + o);
          *reinterpret_cast<uint4*>(static_cast<uint8_t*>(P.out) + o) = v;
        }
      }
    };
    Item I, Inext;
    if (get_item(0, cid, ncl, P.B, nchunk, I)) build_adj(I.b, 0, I.primary);
    for (int iter = 0; get_item(iter, cid, ncl, P.B, nchunk, I); ++iter) {
      const int b = I.b;
      const bool zsave = P.z_save != nullptr && I.primary;
      // ---- Z phase: BOTH warpgroups convert every step, warpgroup g the 64 staging columns [64 g, 64 g + 64) -> packed bf16
      //      at TMEM columns [64 t + 32 g, + 32), so a step's conversion takes half as long and overlaps the next step's MMAs;
      //      optional copy out for the backward (64 features per warpgroup and step) through the warp staging.
      //      Even steps convert in place: warpgroup 1 writes columns that warpgroup 0 reads, and the next (odd) step's
      //      destination is the upper half of this staging area, so the two warps that share a TMEM lane quarter meet at a
      //      64-thread named barrier between their loads and their stores.
      for (int t = 0; t < nstep; ++t) {
        const int par = t & 1;
        const uint32_t src = tl + (par ? T_OUT : 64 * t) + 64 * g, dst = tl + 64 * t + 32 * g;
        GVIT_TR(13);
        mbar_wait(&ctl->zs_full[par], (par ? zs_seen1 : zs_seen0) & 1);
        GVIT_TR(14);
        if (par) ++zs_seen1; else ++zs_seen0;
        tc_fence_after();
        uint32_t pk0[16], pk1[16];
        {
          float va[32], vb[32];
          tmem_ld32x2(src, src + 32, va, vb);
#pragma unroll
          for (int e = 0; e < 16; ++e) { pk0[e] = pack2(va[2 * e], va[2 * e + 1]); pk1[e] = pack2(vb[2 * e], vb[2 * e + 1]); }
        }
        if (!par) {
          tc_fence_before();
          asm volatile("bar.sync %0, 64;" ::"r"(2 + (warp & 3)) : "memory");
          tc_fence_after();
        }
        GVIT_TR(15);
        tmem_st16(dst, pk0);
        tmem_st16(dst + 16, pk1);
        uint8_t* stgz = stg0 + par * WSTAGE;                             // steps alternate between the two staging tiles
        if (zsave) {                                                     // [32 rows][64 features] of this warp -> staging -> one TMA tile
          if (lane == 0) {                                               // store (asynchronous); this tile's previous store (two steps
            if (t < 2) tma_store_wait_read();                            // back, or the previous item's last ones) has read it - the
            else tma_store_wait_read1();                                 // other tile's may still be pending
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(stgz + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(pk0[4 * q], pk0[4 * q + 1], pk0[4 * q + 2], pk0[4 * q + 3]);
            *reinterpret_cast<uint4*>(stgz + lane * 128 + (((4 + q) ^ (lane & 7)) << 4)) = make_uint4(pk1[4 * q], pk1[4 * q + 1], pk1[4 * q + 2], pk1[4 * q + 3]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(par ? conv_doneL1 : conv_doneL0);
        GVIT_TR(16);
        if (zsave) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && wrow0 < P.Np) {                               // rows >= Np are clipped by the TMA unit
            tma_store_3d(&tm_z, stgz, t * 128 + g * 64, wrow0, b);
            tma_store_commit();
          }
        }
      }
      // ---- projection epilogue.  Unit = one [32 rows x 128 bytes] tile of this warp: the 64-feature half g of a chunk on a
      //      bf16 stream, one 32-feature half of it on the fp32 stream.  Units alternate between the two staging tiles: while
      //      unit u is worked on in tile u & 1 (residual tile landed by TMA -> + branch value + bias in place -> TMA store), the
      //      residual tile of unit u + 1 is already loading into the other one.
      constexpr int UPC = RES32 ? 2 : 1;                                 // units per chunk
      const int nunit = (I.c1 - I.c0) * UPC;                             // this warpgroup's units of the item
      const bool has_res = P.resid != nullptr;
      auto load_unit = [&](int u, int tile) {                            // lane 0 only
        if (u < nunit && has_res && wact) {
          const int n = 2 * I.c0 + g + 2 * (u / UPC);                    // 64-feature index
          uint64_t* bar = &ctl->rbar[warp][tile];
          mbar_expect_tx(bar, (uint32_t)WSTAGE);
          tma_load_3d(stg0 + tile * WSTAGE, &tm_res, n * 64 + (RES32 ? (u & 1) * 32 : 0), wrow0, b, bar);
        }
      };
      if (lane == 0) { tma_store_wait_read(); load_unit(0, 0); }         // first residual tile: in flight during the build
      __syncwarp();
      if (get_item(iter + 1, cid, ncl, P.B, nchunk, Inext)) build_adj(Inext.b, iter + 1, Inext.primary);   // next adjacency tile, under this projection
      float v0[32], v1[32];
      for (int u = 0; u < nunit; ++u) {
        const int tile = u & 1;
        const int n = 2 * I.c0 + g + 2 * (u / UPC);
        uint8_t* stg = stg0 + tile * WSTAGE;
        GVIT_TR(22);
        if (lane == 0) { tma_store_wait_read(); load_unit(u + 1, tile ^ 1); }   // the other tile's store has read it: refill it
        __syncwarp();
        GVIT_TR(23);
        if (!RES32 || (u & 1) == 0) {                                    // a new chunk: take it out of TMEM and hand the buffer back
          GVIT_TR(17);
          mbar_wait(&ctl->out_full, full_seen & 1);
          GVIT_TR(18);
          ++full_seen;
          tc_fence_after();
          tmem_ld32x2(tl + T_OUT + g * 64, tl + T_OUT + g * 64 + 32, v0, v1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(out_freeL);
          GVIT_TR(19);
        }
        if (!wact) continue;                                             // no token rows: nothing to add, nothing to store
        if (has_res) {
          mbar_wait(&ctl->rbar[warp][tile], (tile ? rph1 : rph0) & 1);   // the residual tile has landed
          if (tile) ++rph1; else ++rph0;
        }
        GVIT_TR(24);
        if constexpr (!RES32) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t off = lane * 128 + ((q ^ (lane & 7)) << 4);
            uint4 r4 = make_uint4(0, 0, 0, 0);
            if (has_res) r4 = *reinterpret_cast<const uint4*>(stg + off);
            const uint4 b4 = *reinterpret_cast<const uint4*>(&ctl->bias[n * 64 + q * 8]);
            const float* vv = q < 4 ? &v0[8 * q] : &v1[8 * (q - 4)];
            const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
            uint32_t oo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              oo[e] = pack2(vv[2 * e] + bf_lo(bb[e]) + bf_lo(rr[e]), vv[2 * e + 1] + bf_hi(bb[e]) + bf_hi(rr[e]));
            *reinterpret_cast<uint4*>(stg + off) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
          }
        } else {
          const int hh = u & 1;                                          // 32 features = 128 bytes of fp32 per row and half
          const float* vv = hh == 0 ? v0 : v1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {                                  // 4 features per 16-byte chunk
            const uint32_t off = lane * 128 + ((q ^ (lane & 7)) << 4);
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res) r4 = *reinterpret_cast<const float4*>(stg + off);
            const uint2 b2 = *reinterpret_cast<const uint2*>(&ctl->bias[n * 64 + hh * 32 + q * 4]);
            // the branch value as the bf16 projection would store it, then the fp32 add (autocast semantics)
            const uint32_t y01 = pack2(vv[4 * q] + bf_lo(b2.x), vv[4 * q + 1] + bf_hi(b2.x));
            const uint32_t y23 = pack2(vv[4 * q + 2] + bf_lo(b2.y), vv[4 * q + 3] + bf_hi(b2.y));
            r4.x += bf_lo(y01); r4.y += bf_hi(y01); r4.z += bf_lo(y23); r4.w += bf_hi(y23);
            *reinterpret_cast<float4*>(stg + off) = r4;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {                                                 // tm_out starts at token row 1: the CLS row is skipped
          tma_store_3d(&tm_out, stg, n * 64 + (RES32 ? (u & 1) * 32 : 0), wrow0, b);
          tma_store_commit();
        }
      }
    }
  }
  if (warp < 8 && lane == 0) tma_store_wait_all();                       // every tile store of this warp has landed
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  GVIT_SPAN(1);
  if (warp == 9) tmem_dealloc_2sm(tmem, 512);
}

template <int KT, bool RES32>
int launch2(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const CUtensorMap& tm_z, const CUtensorMap& tm_out,
            const CUtensorMap& tm_res, const Params& P, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(agg4_tc_kernel<KT, RES32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  int pairs = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((num_sms() / 2) * 2);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&pairs, agg4_tc_kernel<KT, RES32>, &cfg) != cudaSuccess || pairs < 1) {
      (void)cudaGetLastError();
      pairs = num_sms() / 2;
    }
  }
  // experiment switch (profiles/r5b_batch_sweep.txt): fewer pairs -> is the kernel bound per SM or chip-wide?
  static const int max_pairs = [] { const char* e = getenv("GVIT_AGG_MAXPAIRS"); return e ? atoi(e) : 0; }();
  if (max_pairs >= 1 && max_pairs < pairs) pairs = max_pairs;
  // fewer images than pairs: two pairs per image when they fit (get_item splits the output chunks between them)
  const int used = P.B >= pairs ? pairs : (2 * P.B <= pairs && (P.D / 128) % 2 == 0 ? 2 * P.B : P.B);
  const int grid = 2 * used;
  agg4_tc_kernel<KT, RES32><<<grid, THREADS, SMEM_BYTES, st>>>(tm_tok, tm_w, tm_z, tm_out, tm_res, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}
template <int KT>
int launch(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const CUtensorMap& tm_z, const CUtensorMap& tm_out, const CUtensorMap& tm_res,
           const Params& P, bool res32, cudaStream_t st) {
  return res32 ? launch2<KT, true>(tm_tok, tm_w, tm_z, tm_out, tm_res, P, st) : launch2<KT, false>(tm_tok, tm_w, tm_z, tm_out, tm_res, P, st);
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_agg4)

bool agg4_tc_supported(int Np, int D, int k) {
  return Np > 128 && Np <= 256 && D >= 128 && D % 128 == 0 && D <= 768 && k <= 16;   // two row tiles; both half-slabs non-empty
}

int agg4_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
                const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
                cudaStream_t st) {
  Params P;
  P.B = B; P.Np = Np; P.D = D; P.k = k;
  P.NT = (Np + 15) & ~15;
  P.R0 = ((P.NT + 31) / 32) * 16;                                      // 144 <= NT <= 256: 80 <= R0 <= 128 tokens, NT - R0 in [64, R0]
  P.idx = idx; P.vals = vals;
  P.bias = static_cast<const __nv_bfloat16*>(bias);
  P.resid = resid;
  P.out = out;
  const bool res32 = resid_dtype == GVIT_F32;
  P.w_save = w_save;
  P.kvec = (k == 4 || k == 8 || k == 16) && aligned16(idx) && aligned16(vals) && (!w_save || aligned16(w_save));
  P.z_save = static_cast<__nv_bfloat16*>(z_save);
  P.zbs = z_batch_stride;

  CUtensorMap tm_tok, tm_w, tm_z, tm_out, tm_res;
  const __nv_bfloat16* tok = static_cast<const __nv_bfloat16*>(h) + D;   // skip the CLS row (section 9, G0)
  int rc = make_tmap_bf16_3d(&tm_tok, tok, D, Np, B, D, (uint64_t)(Np + 1) * D, P.R0);   // half-slab boxes
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_w, Wg, D, D, 1, D, (uint64_t)D * D, 64);
  if (rc != GVIT_OK) return rc;
  // output / saved-Z tiles of [32 rows][128 bytes] leave through TMA stores: out starts at token row 1 (the CLS row is copied
  // by ordinary stores), rows >= Np of a warp's tile are clipped by the TMA unit
  if (res32) rc = make_tmap_f32_3d(&tm_out, static_cast<float*>(out) + D, D, Np, B, D, (uint64_t)(Np + 1) * D, 32);
  else rc = make_tmap_bf16_3d(&tm_out, static_cast<__nv_bfloat16*>(out) + D, D, Np, B, D, (uint64_t)(Np + 1) * D, 32);
  if (rc != GVIT_OK) return rc;
  if (z_save) {
    rc = make_tmap_bf16_3d(&tm_z, z_save, D, Np, B, D, (uint64_t)z_batch_stride, 32);
    if (rc != GVIT_OK) return rc;
  } else {
    tm_z = tm_out;                                                       // never dereferenced
  }
  if (resid) {                                                         // residual tiles arrive through TMA loads (token row 1 onwards)
    if (res32) rc = make_tmap_f32_3d(&tm_res, static_cast<const float*>(resid) + D, D, Np, B, D, (uint64_t)(Np + 1) * D, 32);
    else rc = make_tmap_bf16_3d(&tm_res, static_cast<const __nv_bfloat16*>(resid) + D, D, Np, B, D, (uint64_t)(Np + 1) * D, 32);
    if (rc != GVIT_OK) return rc;
  } else {
    tm_res = tm_out;                                                     // never dereferenced
  }
  if (k <= 4) return launch<4>(tm_tok, tm_w, tm_z, tm_out, tm_res, P, res32, st);
  if (k <= 8) return launch<8>(tm_tok, tm_w, tm_z, tm_out, tm_res, P, res32, st);
  return launch<16>(tm_tok, tm_w, tm_z, tm_out, tm_res, P, res32, st);
}

}  // namespace gvit

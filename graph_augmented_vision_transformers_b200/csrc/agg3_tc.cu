// agg3_tc.cu - fused aggregation G4+G5+G6 + residual for bf16, D <= 768 (north_star: "gather /
// normalise, then a tcgen05 GEMM with TMA-staged tiles that keeps A.X in shared memory or TMEM with no HBM round-trip").
//
//   out[b,1+i,:] = resid[b,1+i,:] + ( sum_j softmax_k(vals)_ij * p[b, idx_ij, :] ) Wg^T + bias
//
// One CTA owns 128 token rows of one image and ALL D output features, in two phases:
//   Z phase    : Z = A~ . P for the whole feature dimension, 128 features per step: tcgen05.mma (A~ dense bf16 in shared
//                memory, K = tokens; B = two TMA-staged 64-feature token slabs read MN-major, N = 128) into an fp32
//                staging area of TMEM; the row warps convert it IN TENSOR MEMORY to packed bf16.  After the phase the
//                whole aggregated tile Z [128 x D] sits in TMEM columns [0, D/2) - it never touches shared memory or HBM
//                (training additionally streams a copy out for the weight gradient).
//   projection : for each 64-feature output chunk: OUT = Z . W_chunk^T with A read from TMEM (tcgen05.mma TS form,
//                N/2 + 10 cycles per instruction on B200 against N/2 + 41 for an A operand in shared memory) and W_chunk
//                streamed through a 4-slot TMA ring; the fp32 chunk is double-buffered in TMEM so that bias + residual +
//                store of chunk n overlap the MMAs of chunk n+1.
// v2 (agg_tc.cu, kept for D = 1024) split the output features over CTAs, recomputed Z per split and interleaved one
// Z slab with one projection slab: 21 short dependent phases per CTA, 0.30 ms at B = 256.
//
// TMEM (512 columns): Z bf16 [0, D/2) | Z fp32 staging: step t even -> [64t, 64t+128) (in place), odd -> [384, 512)
//                     | OUT chunk buffers [384, 448), [448, 512) (projection phase).
// Shared memory: A~ (4 x 16 KB) | ring of 4 x 32 KB slots (token slabs, then W pieces) | 8 x 4 KB per-warp staging.
// Warp roles: 0-3 row warpgroup 0, 4-7 row warpgroup 1 (thread <-> token row = TMEM lane; warpgroup g owns the Z steps
// and output chunks of parity g, i.e. one of the two TMEM staging / OUT buffers), 8 TMA producer, 9 MMA issuer.
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int THREADS = 320;
constexpr int TILE = 128 * 128;             // [128 rows][64 bf16]
constexpr int A_BYTES = 4 * TILE;           // A~ [128][256] as four 64-column blocks
constexpr int SLOT = 32 * 1024;             // ring slot: one token slab [<=256][64] or one W piece [64][<=256]
constexpr int NSLOT = 4;
constexpr int WSTAGE = 4 * 1024;            // per-warp staging ([32 rows][128 B]) for coalesced global access
constexpr int T_OUT = 384;                  // first OUT chunk buffer / odd Z staging

struct __align__(8) Ctrl {
  uint64_t full[NSLOT], empty[NSLOT], a_ready, zs_full[2], conv_done[2], out_full[2], out_free[2];
  uint32_t tmem_base;
  __align__(16) __nv_bfloat16 bias[768];
};
constexpr size_t SMEM_BYTES = A_BYTES + NSLOT * SLOT + 8 * WSTAGE + sizeof(Ctrl);

struct Params {
  int Np, D, k, NT;
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

// RES32: resid / out are fp32 - the residual stream torch.autocast keeps in fp32 (vit.py:117-118 add onto an fp32 x); the
// branch value (projection + bias) is rounded to bf16 first, exactly what `x + graph(norm_g(x))` computes under autocast.
template <int KT, bool RES32>
__global__ void __launch_bounds__(THREADS, 1) agg3_tc_kernel(const __grid_constant__ CUtensorMap tm_tok,
                                                             const __grid_constant__ CUtensorMap tm_w, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;
  if ((smem_u32(sA) & 1023u) != 0) __trap();
  uint8_t* sRing = sA + A_BYTES;
  uint8_t* sStg = sRing + NSLOT * SLOT;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(sStg + 8 * WSTAGE);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int mt = blockIdx.x, b = blockIdx.y;
  const int D = P.D, NT = P.NT;
  const int nslab = D / 64;                 // 64-feature token slabs == 64-feature output chunks
  const int nstep = (D + 127) / 128;        // Z steps of (up to) 128 features
  const int npiece = (D + 255) / 256;       // W pieces of (up to) 256 reduction columns per output chunk
  GVIT_TRACE_DECL

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_tok);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->zs_full[s], 1);
      mbar_init(&ctl->conv_done[s], 128);
      mbar_init(&ctl->out_full[s], 1);
      mbar_init(&ctl->out_free[s], 128);
    }
    mbar_init(&ctl->a_ready, 128);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int c = 0;                                                       // ring fill counter
      for (int s = 0; s < nslab; ++s, ++c) {                           // token slabs (whole image: K of A~ . P)
        const int sl = c % NSLOT;
        mbar_wait(&ctl->empty[sl], ((c / NSLOT) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[sl], (uint32_t)(NT * 128));
        tma_load_3d(sRing + sl * SLOT, &tm_tok, s * 64, 0, b, &ctl->full[sl]);
      }
      for (int n = 0; n < nslab; ++n) {                                // W: 64 output features x D, in pieces of 256
        for (int p = 0; p < npiece; ++p, ++c) {
          const int sl = c % NSLOT;
          const int nbox = min(4, (D - p * 256) / 64);
          mbar_wait(&ctl->empty[sl], ((c / NSLOT) & 1) ^ 1);
          mbar_expect_tx(&ctl->full[sl], (uint32_t)(nbox * 8192));
          for (int j = 0; j < nbox; ++j)
            tma_load_3d(sRing + sl * SLOT + j * 8192, &tm_w, p * 256 + j * 64, n * 64, 0, &ctl->full[sl]);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t aA = smem_u32(sA), aR = smem_u32(sRing);
      const uint32_t idesc_w = make_idesc(128, 64, false, false);      // Z from TMEM, W piece K-major
      int c = 0;                                                       // ring consume counter
      mbar_wait(&ctl->a_ready, 0);
      tc_fence_after();
      for (int t = 0; t < nstep; ++t) {                                // ---- Z phase
        const int width = min(128, D - t * 128);
        const uint32_t stg = (t & 1) ? T_OUT : 64 * t;
        if (t >= 2) mbar_wait(&ctl->conv_done[t & 1], ((t - 2) >> 1) & 1);   // staging (odd) / neighbours converted
        const int sl0 = c % NSLOT;
        GVIT_TR(1);
        mbar_wait(&ctl->full[sl0], (c / NSLOT) & 1);
        if (width == 128) mbar_wait(&ctl->full[sl0 + 1], ((c + 1) / NSLOT) & 1);   // slabs 2t, 2t+1: slots (0,1) or (2,3)
        tc_fence_after();
        const uint32_t idesc_z = make_idesc(128, width, false, true);  // A~ K-major, token slabs MN-major
        const uint32_t aTok = aR + sl0 * SLOT;
        for (int ks = 0; ks < NT / 16; ++ks)
          umma_ss(tmem + stg, make_sdesc(aA + (ks >> 2) * TILE + (ks & 3) * 32),
                  width == 128 ? make_sdesc_lbo(aTok + ks * 2048, SLOT) : make_sdesc(aTok + ks * 2048), idesc_z, ks > 0);
        umma_commit(&ctl->zs_full[t & 1]);
        GVIT_TR(2);
        umma_commit(&ctl->empty[sl0]);
        if (width == 128) umma_commit(&ctl->empty[sl0 + 1]);
        c += width == 128 ? 2 : 1;
      }
      // every Z step converted to bf16 (in order, so the last one or two arrivals cover all of them)
      if (nstep >= 2) mbar_wait(&ctl->conv_done[(nstep - 2) & 1], ((nstep - 2) >> 1) & 1);
      mbar_wait(&ctl->conv_done[(nstep - 1) & 1], ((nstep - 1) >> 1) & 1);
      tc_fence_after();
      GVIT_TR(3);
      for (int n = 0; n < nslab; ++n) {                                // ---- projection
        const int buf = n & 1;
        mbar_wait(&ctl->out_free[buf], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        GVIT_TR(4);
        for (int p = 0; p < npiece; ++p, ++c) {
          const int sl = c % NSLOT;
          const int nbox = min(4, (D - p * 256) / 64);
          mbar_wait(&ctl->full[sl], (c / NSLOT) & 1);
          tc_fence_after();
          GVIT_TR(5);
          const uint32_t aW = aR + sl * SLOT;
          for (int j = 0; j < nbox; ++j)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ts(tmem + T_OUT + buf * 64, tmem + (p * 256 + j * 64 + kk * 16) / 2, make_sdesc(aW + j * 8192 + kk * 32),
                      idesc_w, p > 0 || j > 0 || kk > 0);
          umma_commit(&ctl->empty[sl]);
        }
        umma_commit(&ctl->out_full[buf]);
        GVIT_TR(6);
      }
    }
  } else {
    // ------------------------------------------------------------------ row warpgroups
    const int g = warp >> 2;                                           // warpgroup: parity of the steps / chunks it owns
    const int row = threadIdx.x & 127;                                 // == TMEM lane
    const int rowg = mt * 128 + row;
    const bool valid = rowg < P.Np;
    uint8_t* stg = sStg + warp * WSTAGE;                               // this warp's private staging (4 KB)
    const int wrow0 = mt * 128 + (warp & 3) * 32;                      // first token row of this warp
    // ---- G4 + adjacency tile: zero A~, then scatter each row's k softmax weights (bf16) at its neighbour columns
    {
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sA)[i] = z4;
      for (int i = threadIdx.x; i < 768 / 8; i += 256)
        reinterpret_cast<uint4*>(ctl->bias)[i] = (P.bias && i < D / 8) ? reinterpret_cast<const uint4*>(P.bias)[i] : z4;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 0) {
        if (valid) {
          float w[KT];
          int nb[KT];
          float mx = -FLT_MAX, sum = 0.f;
          const int64_t o = ((int64_t)b * P.Np + rowg) * P.k;
#pragma unroll
          for (int j = 0; j < KT; ++j) {
            const bool on = j < P.k;
            nb[j] = on ? P.idx[o + j] : 0;
            w[j] = on ? P.vals[o + j] : -FLT_MAX;
            mx = fmaxf(mx, w[j]);
          }
#pragma unroll
          for (int j = 0; j < KT; ++j) { w[j] = j < P.k ? expf(w[j] - mx) : 0.f; sum += w[j]; }
          const float inv = 1.0f / sum;
#pragma unroll
          for (int j = 0; j < KT; ++j) {
            if (j < P.k) {
              const float wj = w[j] * inv;
              if (P.w_save) P.w_save[o + j] = wj;
              const int cidx = nb[j];
              if (nb_ok(cidx, P.Np))
                *reinterpret_cast<__nv_bfloat16*>(sA + (cidx >> 6) * TILE + swz128(row, cidx & 63) + (cidx & 7) * 2) = __float2bfloat16_rn(wj);
            }
          }
        }
        fence_async_smem();
        mbar_arrive(&ctl->a_ready);
        GVIT_TR(10);
      } else if (mt == 0) {
        // CLS row: out[b,0,:] = resid[b,0,:] (the graph leaves CLS untouched, section 9 G0), 16 bytes per thread and trip
        constexpr int EPV = RES32 ? 4 : 8;                               // elements per 16-byte vector
        constexpr int ESZ = RES32 ? 4 : 2;
        for (int c = row; c < D / EPV; c += 128) {
          const int64_t o = ((int64_t)b * (P.Np + 1) * D + c * EPV) * ESZ;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (P.resid) v = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(P.resid) + o);
          *reinterpret_cast<uint4*>(static_cast<uint8_t*>(P.out) + o) = v;
        }
      }
    }
    const int ch8 = lane & 7, r8 = lane >> 3;                           // coalesced pattern: 8 lanes per 128-byte row segment
    // residual rows of this warp for output chunk n, coalesced (lane -> rows r8 + 4i, 16-byte chunk ch8)
    // RES32: a 64-feature chunk is 256 bytes per row = two 128-byte halves (32 features each)
    constexpr int NH = RES32 ? 2 : 1;
    auto load_resid = [&](int n, uint4 (&rr)[8 * NH]) {
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r8 + 4 * i;
          rr[hh * 8 + i] = make_uint4(0, 0, 0, 0);
          if (P.resid && n < nslab && wrow0 + r < P.Np) {
            const int64_t e = ((int64_t)b * (P.Np + 1) + 1 + wrow0 + r) * D + n * 64;
            if constexpr (RES32) rr[hh * 8 + i] = *reinterpret_cast<const uint4*>(static_cast<const float*>(P.resid) + e + hh * 32 + ch8 * 4);
            else rr[i] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(P.resid) + e + ch8 * 8);
          }
        }
    };
    uint4 rnext[8 * NH];
    load_resid(g, rnext);                                              // first chunk of this warpgroup: in flight during Z
    // ---- Z phase (steps of parity g): fp32 staging -> packed bf16 at TMEM columns [64t, 64t + width/2);
    //      optional copy out for the backward, 64 features at a time through the warp staging
    const uint32_t tl = tmem_lane_base(tmem, warp);
    for (int t = g; t < nstep; t += 2) {
      const int width = min(128, D - t * 128);
      const uint32_t src = tl + ((t & 1) ? T_OUT : 64 * t), dst = tl + 64 * t;
      mbar_wait(&ctl->zs_full[t & 1], (t >> 1) & 1);
      // an odd step's bf16 destination is the upper half of the previous (even) step's in-place staging, which the
      // OTHER warpgroup converts: wait until it has read it
      if (t & 1) mbar_wait(&ctl->conv_done[0], ((t - 1) >> 1) & 1);
      // ... and an even step waits for the other warpgroup's previous (odd) step.  Not a data dependency - a PHASE guard:
      // the MMA warp's step t only needs conv_done[t & 1], so warpgroup 0 could finish steps 0 AND 2 (two completions of
      // conv_done[0]) before a delayed warpgroup 1 (it starts with the CLS-row copy and the residual prefetch, both global
      // round trips) had begun its parity wait for step 1 - which then named a phase two completions back and never
      // returned.  One CTA in ~10^5 hung that way (r2y: mbarrier timeout in a second-wave mt = 0 CTA at B = 128).
      else if (t >= 2) mbar_wait(&ctl->conv_done[1], ((t - 2) >> 1) & 1);
      tc_fence_after();
      GVIT_TR(11);
      for (int h0 = 0; h0 < width; h0 += 64) {
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          const int c0 = h0 + cc;
          float v[32];
          tmem_ld32(src + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack2(v[2 * e], v[2 * e + 1]);
          tmem_st16(dst + (c0 >> 1), pk);                             // in place for even t: columns already read
          if (P.z_save) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((((cc >> 3) + q) ^ (lane & 7)) << 4)) =
                  make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        if (h0 + 64 >= width) {                                        // whole step converted: release it to the MMA warp
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&ctl->conv_done[t & 1]);
          GVIT_TR(12);
        }
        if (P.z_save) {                                                // coalesced: 4 whole 128-byte row segments per instr
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r8 + 4 * i;
            if (wrow0 + r < P.Np) {
              const uint4 v4 = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
              *reinterpret_cast<uint4*>(P.z_save + (int64_t)b * P.zbs + (int64_t)(wrow0 + r) * D + t * 128 + h0 + ch8 * 8) = v4;
            }
          }
          __syncwarp();
        }
      }
    }
    // ---- projection epilogue (chunks of parity g): + bias + residual, coalesced through the warp staging
    for (int n = g; n < nslab; n += 2) {
      const int buf = n & 1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r8 + 4 * i;
        *reinterpret_cast<uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4)) = rnext[i];
      }
      if constexpr (!RES32) load_resid(n + 2, rnext);                  // next chunk of this warpgroup: in flight meanwhile
      __syncwarp();
      GVIT_TR(13);
      mbar_wait(&ctl->out_full[buf], (n >> 1) & 1);
      tc_fence_after();
      GVIT_TR(14);
      float v0[32], v1[32];
      tmem_ld32(tl + T_OUT + buf * 64, v0);
      tmem_ld32(tl + T_OUT + buf * 64 + 32, v1);
      tc_fence_before();
      mbar_arrive(&ctl->out_free[buf]);
      if constexpr (!RES32) {
        __nv_bfloat16* outp = static_cast<__nv_bfloat16*>(P.out);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t off = lane * 128 + ((q ^ (lane & 7)) << 4);
          const uint4 r4 = *reinterpret_cast<const uint4*>(stg + off);
          const uint4 b4 = *reinterpret_cast<const uint4*>(&ctl->bias[n * 64 + q * 8]);
          const float* vv = q < 4 ? &v0[8 * q] : &v1[8 * (q - 4)];
          const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
          uint32_t oo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            oo[e] = pack2(vv[2 * e] + bf_lo(bb[e]) + bf_lo(rr[e]), vv[2 * e + 1] + bf_hi(bb[e]) + bf_hi(rr[e]));
          *reinterpret_cast<uint4*>(stg + off) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r8 + 4 * i;
          if (wrow0 + r < P.Np) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(outp + ((int64_t)b * (P.Np + 1) + 1 + wrow0 + r) * D + n * 64 + ch8 * 8) = v4;
          }
        }
        __syncwarp();
      } else {
        float* outp = static_cast<float*>(P.out);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                               // 32 features = 128 bytes of fp32 per row and half
          if (hh == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = r8 + 4 * i;
              *reinterpret_cast<uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4)) = rnext[8 + i];
            }
            load_resid(n + 2, rnext);                                  // both halves consumed: next chunk in flight
            __syncwarp();
          }
          const float* vv = hh == 0 ? v0 : v1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {                                // 4 features per 16-byte chunk
            const uint32_t off = lane * 128 + ((q ^ (lane & 7)) << 4);
            float4 r4 = *reinterpret_cast<const float4*>(stg + off);
            const uint2 b2 = *reinterpret_cast<const uint2*>(&ctl->bias[n * 64 + hh * 32 + q * 4]);
            // the branch value as the bf16 projection would store it, then the fp32 add (autocast semantics)
            const uint32_t y01 = pack2(vv[4 * q] + bf_lo(b2.x), vv[4 * q + 1] + bf_hi(b2.x));
            const uint32_t y23 = pack2(vv[4 * q + 2] + bf_lo(b2.y), vv[4 * q + 3] + bf_hi(b2.y));
            r4.x += bf_lo(y01); r4.y += bf_hi(y01); r4.z += bf_lo(y23); r4.w += bf_hi(y23);
            *reinterpret_cast<float4*>(stg + off) = r4;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r8 + 4 * i;
            if (wrow0 + r < P.Np) {
              const uint4 v4 = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
              *reinterpret_cast<uint4*>(outp + ((int64_t)b * (P.Np + 1) + 1 + wrow0 + r) * D + n * 64 + hh * 32 + ch8 * 4) = v4;
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

template <int KT, bool RES32>
int launch2(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const Params& P, int B, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(agg3_tc_kernel<KT, RES32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  dim3 grid((P.Np + 127) / 128, B);
  agg3_tc_kernel<KT, RES32><<<grid, THREADS, SMEM_BYTES, st>>>(tm_tok, tm_w, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}
template <int KT>
int launch(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const Params& P, int B, bool res32, cudaStream_t st) {
  return res32 ? launch2<KT, true>(tm_tok, tm_w, P, B, st) : launch2<KT, false>(tm_tok, tm_w, P, B, st);
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_agg)

bool agg3_tc_supported(int Np, int D, int k) {
  return Np >= 16 && Np <= 256 && D >= 64 && D % 64 == 0 && D <= 768 && k <= 16;
}

int agg3_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
                const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
                cudaStream_t st) {
  Params P;
  P.Np = Np; P.D = D; P.k = k;
  P.NT = (Np + 15) & ~15;
  GVIT_REQUIRE(B <= 65535, GVIT_ERR_SHAPE, "agg_fwd: batch %d exceeds the grid limit 65535", B);
  P.idx = idx; P.vals = vals;
  P.bias = static_cast<const __nv_bfloat16*>(bias);
  P.resid = resid;
  P.out = out;
  const bool res32 = resid_dtype == GVIT_F32;
  P.w_save = w_save;
  P.z_save = static_cast<__nv_bfloat16*>(z_save);
  P.zbs = z_batch_stride;

  CUtensorMap tm_tok, tm_w;
  const __nv_bfloat16* tok = static_cast<const __nv_bfloat16*>(h) + D;   // skip the CLS row (section 9, G0)
  int rc = make_tmap_bf16_3d(&tm_tok, tok, D, Np, B, D, (uint64_t)(Np + 1) * D, P.NT);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_w, Wg, D, D, 1, D, (uint64_t)D * D, 64);
  if (rc != GVIT_OK) return rc;
  if (k <= 4) return launch<4>(tm_tok, tm_w, P, B, res32, st);
  if (k <= 8) return launch<8>(tm_tok, tm_w, P, B, res32, st);
  return launch<16>(tm_tok, tm_w, P, B, res32, st);
}

}  // namespace gvit

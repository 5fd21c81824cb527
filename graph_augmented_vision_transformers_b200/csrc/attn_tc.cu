// attn_tc.cu - fused softmax(QK^T * scale)V on tcgen05 for bf16, head_dim 64; replaces
// /root/reference/src/models/vit.py:64-69 (three ATen kernels + a copy, with the (B,H,N,N) score
// tensor materialised four times in HBM) by one kernel that reads q,k,v once and writes o once.
//
// Data layout: the packed projection output of vit.py:59, (B,N,3,H,64), is addressed in place through
// ONE 3-D tensor map {3*H*64, N, B} with box {64, 128, 1}: the Q tile of head h is the box at column
// h*64, the K tile at H*64 + h*64, the V tile at 2*H*64 + h*64.  Rows past N are zero-filled by TMA.
// A [128][64] bf16 tile (128-byte rows, 128B swizzle) is used K-major for Q and K (K = head dim) and
// MN-major for V (K = keys), so no transpose is ever materialised.
//
// Forward: N <= 256 below (attn_fwd_tc2_kernel: a whole head per work item, no online rescale); N > 256 in
//   attn_fwd_long_tc.cu (key blocks + online softmax).
// Backward (one CTA per head x image, N <= 256), key-major so every transposed product is a plain
//   K-major/MN-major operand:  S^T = K Q^T, dP^T = V dO^T  -> threads form P^T, dS^T (and dS, transposed
//   through shared memory) -> dV += P^T dO, dK += dS^T Q, dQ += dS K, all accumulated in TMEM.
//
// Warp roles: 0-3 softmax / epilogue, 4 TMA producer, 5 MMA issuer (+ TMEM allocation).
#include <float.h>
#include <stdlib.h>

#include "gelu.cuh"
#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr int TILE_BYTES = 128 * 128;   // [128 rows][64 bf16]

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error
// 7.5e-5 - far below the bf16 resolution of the probabilities it feeds).  The softmax loops are MUFU-bound
// (one MUFU.EX2 per score, 4 lanes/clk/SMSP), so every fourth exponential is evaluated this way to balance the pipes.
// x <= 0 is expected; x < -125 is clamped (result ~2^-125 instead of 0 / denormal).
__device__ __forceinline__ float ex2_emul(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;                 // 1.5 * 2^23: round to nearest integer in the low mantissa bits
  const float f = x - (xr - 12582912.0f);           // in [-0.5, 0.5]
  float p = fmaf(0.05517164617776871f, f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 64 fp32 -> bf16 into row `row` of a swizzled [128][64] staging tile (the layout a SWIZZLE_128B TMA store reads)
__device__ __forceinline__ void stage_out64(uint8_t* tile, int row, const float (&a)[32], const float (&b)[32], float mul) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack_bf16(a[8 * q + 0] * mul, a[8 * q + 1] * mul);
    w.y = pack_bf16(a[8 * q + 2] * mul, a[8 * q + 3] * mul);
    w.z = pack_bf16(a[8 * q + 4] * mul, a[8 * q + 5] * mul);
    w.w = pack_bf16(a[8 * q + 6] * mul, a[8 * q + 7] * mul);
    *reinterpret_cast<uint4*>(tile + swz128(row, 8 * q)) = w;
    w.x = pack_bf16(b[8 * q + 0] * mul, b[8 * q + 1] * mul);
    w.y = pack_bf16(b[8 * q + 2] * mul, b[8 * q + 3] * mul);
    w.z = pack_bf16(b[8 * q + 4] * mul, b[8 * q + 5] * mul);
    w.w = pack_bf16(b[8 * q + 6] * mul, b[8 * q + 7] * mul);
    *reinterpret_cast<uint4*>(tile + swz128(row, 32 + 8 * q)) = w;
  }
}

// =================================================================================================
// forward, N <= 256 keys (ViT-B/16 at 224: N = 197): persistent and pipelined.
//
// One CTA per SM loops over (image, head) work items.  An item is the whole attention problem of one head: all its
// keys fit ONE tcgen05.mma N extent (NT = ceil16(N) <= 256), so there is no online-softmax rescaling - the score
// row is complete after a single S = Q K^T.  Per item:
//   TMA producer (warp 8)  : Q, K, V tiles of the NEXT item into the other shared-memory stage (2 x 96 KB stages)
//   MMA issuer   (warp 9)  : S_g = Q_g K^T  (g = 0,1: query rows [128g, 128g+128)) -> TMEM region g (256 columns);
//                            O_g = P_g V with P_g read straight from TMEM (tcgen05.mma A-from-TMEM) -> region g
//   softmax WG g (warps 4g..4g+3, thread <-> query row = TMEM lane): row max, p = exp2(..), row sum; P_g is written
//                            back IN PLACE over S_g as packed bf16 (tcgen05.st) - it never touches shared memory;
//                            then O_g / l -> bf16 -> the (dead) Q_g tile of the stage -> one TMA tile store.
// TMEM region g (256 of the 512 columns): S_g at [0, NT) -> P_g at [0, NT/2), O_g at [128, 192).
// Regions, stages and warpgroups are decoupled by mbarriers, so S of item i+1 overlaps the epilogue of item i.
// =================================================================================================
constexpr int F2_THREADS = 320;
constexpr int F2_STAGE_BYTES = 6 * TILE_BYTES;     // Q[256][64] | K[256][64] | V[256][64]
struct __align__(8) F2Ctrl {
  uint64_t full[2], empty[2], s_full[2], p_full[2], o_full[2], t_free[2];
  uint32_t tmem_base;
};
constexpr size_t F2_SMEM = 1024 + 2 * F2_STAGE_BYTES + sizeof(F2Ctrl);

__global__ void __launch_bounds__(F2_THREADS, 1) attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                     const __grid_constant__ CUtensorMap tm_out, int N,
                                                                     int H, int items, float scale,
                                                                     float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // no static __shared__ in these kernels: base is 1024-aligned
  uint8_t* sm = smem_raw;   // NOT rounded through an integer: that made every access a generic LD/ST instead of LDS/STS
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  F2Ctrl* ctl = reinterpret_cast<F2Ctrl*>(sm + 2 * F2_STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role id
  const int MT = (N + 127) >> 7;                 // 128-row query tiles (1 or 2) == 128-row key/value tiles
  const int NT = (N + 15) & ~15;                 // key extent of the MMAs
  GVIT_TRACE_DECL
  GVIT_SPAN(0);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_out);
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], MT); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&ctl->s_full[g], 1);
      mbar_init(&ctl->p_full[g], 128);
      mbar_init(&ctl->o_full[g], 1);
      mbar_init(&ctl->t_free[g], 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int s = it & 1, b = w / H, h = w - b * H;
        mbar_wait(&ctl->empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[s], (uint32_t)(3 * MT * TILE_BYTES));
        uint8_t* st = sm + s * F2_STAGE_BYTES;
        for (int j = 0; j < MT; ++j) {
          tma_load_3d(st + j * TILE_BYTES, &tm_qkv, h * 64, j * 128, b, &ctl->full[s]);
          tma_load_3d(st + (2 + j) * TILE_BYTES, &tm_qkv, (H + h) * 64, j * 128, b, &ctl->full[s]);
          tma_load_3d(st + (4 + j) * TILE_BYTES, &tm_qkv, (2 * H + h) * 64, j * 128, b, &ctl->full[s]);
        }
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      const uint32_t idesc_s = make_idesc(128, NT, false, false);
      const uint32_t idesc_o = make_idesc(128, 64, false, true);
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t aQ = smem_u32(sm + s * F2_STAGE_BYTES), aK = aQ + 2 * TILE_BYTES, aV = aQ + 4 * TILE_BYTES;
        mbar_wait(&ctl->full[s], (it >> 1) & 1);
        GVIT_TR(1);
        for (int g = 0; g < MT; ++g) {
          mbar_wait(&ctl->t_free[g], (it & 1) ^ 1);          // region g drained by the previous item's epilogue
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + g * 256, make_sdesc(aQ + g * TILE_BYTES + kk * 32), make_sdesc(aK + kk * 32), idesc_s, kk > 0);
          umma_commit(&ctl->s_full[g]);
        }
        for (int g = 0; g < MT; ++g) {
          mbar_wait(&ctl->p_full[g], it & 1);                // P_g is in TMEM, S_g fully consumed
          tc_fence_after();
          for (int ks = 0; ks < NT / 16; ++ks)
            umma_ts(tmem + g * 256 + 128, tmem + g * 256 + ks * 8, make_sdesc(aV + ks * 2048), idesc_o, ks > 0);
          umma_commit(&ctl->o_full[g]);
          GVIT_TR(2);
        }
      }
    }
  } else {
    const int g = warp >> 2;
    if (g < MT) {
      const int r = (warp & 3) * 32 + lane;                  // row inside the tile == TMEM lane
      const uint32_t tR = tmem_lane_base(tmem + g * 256, warp);
      const float sl2 = scale * LOG2E;
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int s = it & 1, b = w / H, h = w - b * H;
        GVIT_TR(10);
        mbar_wait(&ctl->s_full[g], it & 1);
        tc_fence_after();
        GVIT_TR(11);
        float mx = -3.0e38f;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld32(tR + c0, v);
          if (c0 + 32 <= N) {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = fmaxf(mx, v[t]);
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = (c0 + t < N) ? fmaxf(mx, v[t]) : mx;
          }
        }
        const float msc = mx * sl2;
        float l = 0.f;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld32(tR + c0, v);
          uint32_t pk[16];
          if (c0 + 32 <= N) {                                // full chunk: no masks, every 4th exp2 off the MUFU pipe
#pragma unroll
            for (int t = 0; t < 32; t += 4) {
              const float p0 = ex2(fmaf(v[t], sl2, -msc));
              const float p1 = ex2(fmaf(v[t + 1], sl2, -msc));
              const float p2 = ex2(fmaf(v[t + 2], sl2, -msc));
              const float p3 = ex2_emul(fmaf(v[t + 3], sl2, -msc));
              l += (p0 + p1) + (p2 + p3);
              pk[t >> 1] = pack_bf16(p0, p1);
              pk[(t >> 1) + 1] = pack_bf16(p2, p3);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
              const float p0 = (c0 + t < N) ? ex2(fmaf(v[t], sl2, -msc)) : 0.f;
              const float p1 = (c0 + t + 1 < N) ? ex2(fmaf(v[t + 1], sl2, -msc)) : 0.f;
              l += p0 + p1;
              pk[t >> 1] = pack_bf16(p0, p1);
            }
          }
          tmem_st16(tR + (c0 >> 1), pk);                     // in place: columns [c0/2, c0/2+16) were read already
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ctl->p_full[g]);
        GVIT_TR(12);

        mbar_wait(&ctl->o_full[g], it & 1);
        tc_fence_after();
        GVIT_TR(13);
        const float inv = 1.0f / l;
        uint8_t* so = sm + s * F2_STAGE_BYTES + g * TILE_BYTES;   // the Q_g tile: dead once S_g has been issued
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float v[32];
          tmem_ld32(tR + 128 + hlf * 32, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o4;
            o4.x = pack_bf16(v[8 * q + 0] * inv, v[8 * q + 1] * inv);
            o4.y = pack_bf16(v[8 * q + 2] * inv, v[8 * q + 3] * inv);
            o4.z = pack_bf16(v[8 * q + 4] * inv, v[8 * q + 5] * inv);
            o4.w = pack_bf16(v[8 * q + 6] * inv, v[8 * q + 7] * inv);
            *reinterpret_cast<uint4*>(so + swz128(r, hlf * 32 + 8 * q)) = o4;
          }
        }
        tc_fence_before();
        mbar_arrive(&ctl->t_free[g]);
        GVIT_TR(14);                        // TMEM region g may be overwritten by the next S_g
        const int q = g * 128 + r;
        if (q < N) lse[((int64_t)b * H + h) * N + q] = (msc + log2f(l)) * LN2;
        fence_async_smem();                                  // generic-proxy tile writes -> visible to the TMA store
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if ((warp & 3) == 0 && lane == 0) {
          tma_store_3d(&tm_out, so, h * 64, g * 128, b);     // rows >= N are clipped by the TMA unit
          tma_store_commit();
          tma_store_wait_read();
          mbar_arrive(&ctl->empty[s]);                       // stage s (Q/K/V and the O staging) may be refilled
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  GVIT_SPAN(1);
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// =================================================================================================
// backward (N <= 256): persistent, key-major, pipelined over 64-query sub-tiles.
//
// One CTA per SM loops over (image, head) items; all of a head's Q, K, V, dO tiles (<= 8 x 16 KB) are resident.
// For key tile kt and query tile qt the 128 x 128 score block is processed as TWO 64-query sub-tiles g = 0,1, each
// owned by one softmax warpgroup (warps 4g..4g+3, thread <-> key row = TMEM lane) and one 64-column TMEM sub-buffer:
//   MMA1(g) : S^T_g = K Q_g^T,  dP^T_g = V dO_g^T                     (tcgen05.mma, M = keys, N = <=64 queries)
//   WG g    : P^T_g = exp2(S^T_g c - lse), dS^T_g = P^T_g (dP^T_g - delta)   -> bf16, written back IN PLACE into the
//             TMEM columns they were computed from (two bf16 per 32-bit column); dS^T_g also into an smem staging tile
//   MMA2(g) : dV += P^T_g dO_g,  dK += dS^T_g Q_g     (A operand read from TMEM, accumulate across all query tiles)
//   dQ[qt] += dS K  once per (kt, qt): the [keys][128 queries] dS^T staging pair is read as an MN-major A operand.
// Every MMA here is 128 x 64 x 16, which reads 6 KB of operands from shared memory per 32 tensor-pipe cycles - more than
// the 128 B/clk the SM delivers - so shared-memory bandwidth, not the tensor pipe, paces the kernel: taking the A
// operands of dV and dK from TMEM (and not staging P at all) removes 30 % of that traffic.
// MMA1 of the NEXT (kt, qt) is issued into sub-buffer g right behind MMA2(g) (the pipe executes in order, so the
// in-place operands are consumed first), so the tensor pipe, the two warpgroups and the epilogue stores overlap
// instead of taking turns (v1 ran them strictly in sequence: 470 us).
// TMEM: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ [384,512).
// The dV / dK / dQ epilogues (wait for the step's MMAs, TMEM -> bf16 -> staging tile -> TMA store) belong to a warpgroup of
// their own: on the softmax warpgroups they were a third of every item's critical path (three ~2000-cycle epilogues plus
// the waits for the accumulators to become final), with the tensor pipe idle behind them.
// Warp roles: 0-3 WG0, 4-7 WG1, 8 TMA producer, 9 MMA issuer (+ TMEM allocation), 10-11 delta / lse helpers,
// 12-15 epilogue warpgroup (thread <-> accumulator row = TMEM lane).
// =================================================================================================
constexpr int B2_THREADS = 512;   // 8 softmax warps + TMA + MMA + 2 delta helpers + 4 epilogue warps
struct __align__(16) BwdCtrl {
  float lse2[2][256], delta[2][256];            // double-buffered by item parity
  uint64_t kv_full[2], q_full[2], kv_empty[2], q_empty[2], s_full[2], p_full[2], st_free, dvk_free, dq_free, delta_ready[2], wg_done;
  uint32_t tmem_base;
};
// Q[2], dO[2], K[2], V[2] tiles + output staging (one tile per warpgroup) + dS^T staging (2 sub-tiles)
constexpr size_t BWD_SMEM = 1024 + 12 * TILE_BYTES + sizeof(BwdCtrl);

// two 32-column TMEM reads of this thread's lane, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); }
}

// P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - delta) for W (16 or 32) consecutive queries of this thread's key row,
// written as bf16 into the two staging tiles.  No masks: the (negated) lse2 is -inf for padded queries (p = 0 exactly), and padded
// KEY rows only feed accumulator rows that are never stored (dV, dK) or multiply zero-filled K rows (dQ).  The
// softmax scale is applied to dK / dQ when they are stored.
template <int W>
__device__ __forceinline__ void bwd_chunk(float (&s)[W], float (&dp)[W], const float* lse2, const float* delta,
                                          float sl2, uint32_t tP, uint32_t tdS, uint8_t* dst_tile, int row, int c0) {
  // written in phases over the whole chunk (all exponent arguments, then all MUFU.EX2, then dS) so that W
  // independent transcendental ops are in flight: with two warps per scheduler the loop is latency-, not rate-bound
  // lse2[] and delta[] hold the NEGATED values (the helper warps write them that way), so both affine steps are one packed
  // fp32x2 instruction per two queries (FFMA2 / FADD2 / FMUL2): the loop is issue-bound, and these are a third of it
  const f32x2 sl22 = pk2(sl2, sl2);
#pragma unroll
  for (int q4 = 0; q4 < W; q4 += 4) {
    const ulonglong2 l = *reinterpret_cast<const ulonglong2*>(lse2 + q4);
    up2(fma2(pk2(s[q4 + 0], s[q4 + 1]), sl22, f32x2{l.x}), s[q4 + 0], s[q4 + 1]);
    up2(fma2(pk2(s[q4 + 2], s[q4 + 3]), sl22, f32x2{l.y}), s[q4 + 2], s[q4 + 3]);
  }
#pragma unroll
  for (int e = 0; e < W; ++e) s[e] = (e & 3) == 3 ? ex2_emul(s[e]) : ex2(s[e]);   // balance the MUFU and FMA pipes
#pragma unroll
  for (int q4 = 0; q4 < W; q4 += 4) {
    const ulonglong2 d = *reinterpret_cast<const ulonglong2*>(delta + q4);
    up2(mul2(pk2(s[q4 + 0], s[q4 + 1]), add2(pk2(dp[q4 + 0], dp[q4 + 1]), f32x2{d.x})), dp[q4 + 0], dp[q4 + 1]);
    up2(mul2(pk2(s[q4 + 2], s[q4 + 3]), add2(pk2(dp[q4 + 2], dp[q4 + 3]), f32x2{d.y})), dp[q4 + 2], dp[q4 + 3]);
  }
  uint32_t pk[W / 2], dk[W / 2];
#pragma unroll
  for (int e = 0; e < W; e += 2) {
    pk[e >> 1] = pack_bf16(s[e], s[e + 1]);
    dk[e >> 1] = pack_bf16(dp[e], dp[e + 1]);
  }
  // in place: the 32-bit columns [c0/2, c0/2 + W/2) of both regions were read by this or an earlier chunk
  if constexpr (W == 32) {
    tmem_st16(tP + (c0 >> 1), pk);
    tmem_st16(tdS + (c0 >> 1), dk);
  } else {
    tmem_st8(tP + (c0 >> 1), pk);
    tmem_st8(tdS + (c0 >> 1), dk);
  }
#pragma unroll
  for (int q8 = 0; q8 < W; q8 += 8)
    *reinterpret_cast<uint4*>(dst_tile + swz128(row, c0 + q8)) = make_uint4(dk[q8 / 2], dk[q8 / 2 + 1], dk[q8 / 2 + 2], dk[q8 / 2 + 3]);
}

// width (multiple of 16) of 64-query sub-tile g of a query tile with nq (multiple of 16) padded queries; an
// all-padding second sub-tile still runs, 16 wide, so that every (kt, qt) step has the same barrier traffic
__device__ __forceinline__ int sub_width(int nq, int g) { return g == 0 ? min(64, nq) : max(16, nq - 64); }

__global__ void __launch_bounds__(B2_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                    const __grid_constant__ CUtensorMap tm_do,
                                                                    const __grid_constant__ CUtensorMap tm_dqkv,
                                                                    const __grid_constant__ CUtensorMap tm_o, int N,
                                                                    int H, int items, float scale,
                                                                    const __nv_bfloat16* __restrict__ out,
                                                                    const __nv_bfloat16* __restrict__ dout,
                                                                    const float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // no static __shared__ in these kernels: base is 1024-aligned
  uint8_t* sm = smem_raw;   // NOT rounded through an integer: that made every access a generic LD/ST instead of LDS/STS
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  uint8_t* sQ = sm;                        // 2 tiles
  uint8_t* sdO = sm + 2 * TILE_BYTES;      // 2 tiles
  uint8_t* sK = sm + 4 * TILE_BYTES;       // 2 tiles
  uint8_t* sV = sm + 6 * TILE_BYTES;       // 2 tiles
  uint8_t* sOut = sm + 8 * TILE_BYTES;     // warpgroup g's output staging tile at + g*TILE_BYTES (dV / dK / dQ -> TMA store)
  uint8_t* sdST = sm + 10 * TILE_BYTES;    // sub-tile g at + g*TILE_BYTES: [128 keys][64 queries]; the pair = MN-major A of dQ
  BwdCtrl* ctl = reinterpret_cast<BwdCtrl*>(sm + 12 * TILE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role id
  const int T = (N + 127) / 128;           // tiles along queries == tiles along keys (1 or 2)
  GVIT_TRACE_DECL
  GVIT_SPAN(0);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dqkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->kv_full[i], 1);
      mbar_init(&ctl->q_full[i], 1);
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->p_full[i], 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->kv_empty[i], 1);
      mbar_init(&ctl->q_empty[i], 1);
    }
    mbar_init(&ctl->st_free, 1);
    mbar_init(&ctl->dvk_free, 128);
    mbar_init(&ctl->dq_free, 128);
    mbar_init(&ctl->wg_done, 256);
    mbar_init(&ctl->delta_ready[0], 2);
    mbar_init(&ctl->delta_ready[1], 2);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
  const uint32_t tST = tmem, tdPT = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      int ic = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++ic) {
        const int b = w / H, h = w - b * H;
        const uint32_t prev = (ic - 1) & 1;
        // Each tile pair is refilled as soon as the LAST MMA of the previous item that reads it has retired (per-slot
        // empty barriers, in the order the slots are released), so the inputs of the next item's first steps land while
        // the current item is still computing: with one barrier for all eight tiles every item began with ~5000 idle cycles.
        if (ic > 0) mbar_wait(&ctl->kv_empty[0], prev);        // released after step (kt = 0, qt = T-1)
        GVIT_TR(1);
        mbar_expect_tx(&ctl->kv_full[0], 2 * TILE_BYTES);
        tma_load_3d(sK, &tm_qkv, (H + h) * 64, 0, b, &ctl->kv_full[0]);
        tma_load_3d(sV, &tm_qkv, (2 * H + h) * 64, 0, b, &ctl->kv_full[0]);
        for (int t = 0; t < T; ++t) {
          if (ic > 0) mbar_wait(&ctl->q_empty[t], prev);       // released after step (kt = T-1, qt = t)
          mbar_expect_tx(&ctl->q_full[t], 2 * TILE_BYTES);
          tma_load_3d(sQ + t * TILE_BYTES, &tm_qkv, h * 64, t * 128, b, &ctl->q_full[t]);
          tma_load_3d(sdO + t * TILE_BYTES, &tm_do, h * 64, t * 128, b, &ctl->q_full[t]);
        }
        if (T > 1) {
          if (ic > 0) mbar_wait(&ctl->kv_empty[1], prev);      // released after the last step
          mbar_expect_tx(&ctl->kv_full[1], 2 * TILE_BYTES);
          tma_load_3d(sK + TILE_BYTES, &tm_qkv, (H + h) * 64, 128, b, &ctl->kv_full[1]);
          tma_load_3d(sV + TILE_BYTES, &tm_qkv, (2 * H + h) * 64, 128, b, &ctl->kv_full[1]);
        }
        // the NEXT item's tiles (and its O rows for the delta prologue) into L2 now: there is no shared memory to
        // double-buffer 128 KB of inputs, and all CTAs start their loads together, so an un-prefetched item start
        // waited ~5000 cycles on DRAM with the tensor pipe idle
        const int wn = w + gridDim.x;
        if (wn < items) {
          const int bn = wn / H, hn = wn - bn * H;
          for (int t = 0; t < T; ++t) {
            tma_prefetch_3d(&tm_qkv, (H + hn) * 64, t * 128, bn);
            tma_prefetch_3d(&tm_qkv, (2 * H + hn) * 64, t * 128, bn);
            tma_prefetch_3d(&tm_qkv, hn * 64, t * 128, bn);
            tma_prefetch_3d(&tm_do, hn * 64, t * 128, bn);
            tma_prefetch_3d(&tm_o, hn * 64, t * 128, bn);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t adST = smem_u32(sdST);
      const uint32_t idesc_mn = make_idesc(128, 64, false, true);    // A K-major (staging), B MN-major (dO / Q rows)
      const uint32_t idesc_tt = make_idesc(128, 64, true, true);     // A MN-major (dS), B MN-major (K rows)
      int ic = 0, itc = 0, kc = 0;                                   // items, (kt,qt) steps, key tiles done so far
      // S^T_g = K_kt Q_{qt,g}^T and dP^T_g = V_kt dO_{qt,g}^T into TMEM sub-buffer g
      auto issue_mma1 = [&](int kt, int qt, int g) {
        const int nq = (min(128, N - qt * 128) + 15) & ~15;
        const uint32_t idesc = make_idesc(128, sub_width(nq, g), false, false);
        const uint32_t boff = qt * TILE_BYTES + g * 8192;            // query rows [64g, 64g+64) of the tile
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tST + g * 64, make_sdesc(aK + kt * TILE_BYTES + kk * 32), make_sdesc(aQ + boff + kk * 32), idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tdPT + g * 64, make_sdesc(aV + kt * TILE_BYTES + kk * 32), make_sdesc(adO + boff + kk * 32), idesc, kk > 0);
        umma_commit(&ctl->s_full[g]);
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++ic) {
        const int par = ic & 1;
        mbar_wait(&ctl->kv_full[0], par);
        mbar_wait(&ctl->q_full[0], par);
        tc_fence_after();
        GVIT_TR(2);
        // the TMEM sub-buffers were drained when p_full of the previous step completed (waited below)
        issue_mma1(0, 0, 0);
        issue_mma1(0, 0, 1);
        for (int kt = 0; kt < T; ++kt) {
          const int nk = (min(128, N - kt * 128) + 15) & ~15;
          for (int qt = 0; qt < T; ++qt, ++itc) {
            const int nq = (min(128, N - qt * 128) + 15) & ~15;
            const bool last = (kt == T - 1 && qt == T - 1);
            const int nkt = (qt == T - 1) ? kt + 1 : kt, nqt = (qt == T - 1) ? 0 : qt + 1;
            for (int g = 0; g < 2; ++g) {
              mbar_wait(&ctl->p_full[g], itc & 1);       // P^T_g / dS^T_g are in TMEM (and dS^T_g staged)
              tc_fence_after();
              GVIT_TR(3 + g);
              if (g == 0 && qt == 0 && kc > 0) {          // dV / dK are about to be overwritten: epilogue has read them
                mbar_wait(&ctl->dvk_free, (kc - 1) & 1);
                tc_fence_after();
              }
              const int wg = sub_width(nq, g);
              for (int ks = 0; ks < wg / 16; ++ks) {      // K = queries of the sub-tile; A = 8 packed TMEM columns per slice
                const uint32_t boff = qt * TILE_BYTES + g * 8192 + ks * 2048;
                const bool acc = qt > 0 || g > 0 || ks > 0;
                umma_ts(tdV, tST + g * 64 + ks * 8, make_sdesc(adO + boff), idesc_mn, acc);
                umma_ts(tdK, tdPT + g * 64 + ks * 8, make_sdesc(aQ + boff), idesc_mn, acc);
              }
              if (!last) {                                // behind MMA2(g) in the (in-order) pipe: it overwrites its A operands
                if (g == 0) {
                  if (nqt == 0) mbar_wait(&ctl->kv_full[nkt], par);   // first use of the next key tile
                  else if (kt == 0) mbar_wait(&ctl->q_full[nqt], par); // first use of the next query tile
                  tc_fence_after();
                }
                issue_mma1(nkt, nqt, g);                  // next scores overlap this step's dQ and the other sub-tile
              }
            }
            if (kt == 0 && qt == 0 && ic > 0) {           // dQ is about to be overwritten: previous item stored it
              mbar_wait(&ctl->dq_free, (ic - 1) & 1);
              tc_fence_after();
            }
            for (int ks = 0; ks < nk / 16; ++ks)          // dQ[qt] += dS K_kt, K = keys of this tile
              umma_ss(tdQ + qt * 64, make_sdesc_lbo(adST + ks * 2048, TILE_BYTES), make_sdesc(aK + kt * TILE_BYTES + ks * 2048),
                      idesc_tt, kt > 0 || ks > 0);
            umma_commit(&ctl->st_free);                   // staging consumed; accumulators of this step final
            if (qt == T - 1) umma_commit(&ctl->kv_empty[kt]);   // K / V tile kt: no later MMA of this item reads it
            if (kt == T - 1) umma_commit(&ctl->q_empty[qt]);    // Q / dO tile qt: likewise
            GVIT_TR(5);
            if (qt == T - 1) ++kc;
          }
        }
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ epilogue warpgroup
    const int t = (warp & 3) * 32 + lane;                 // accumulator row == TMEM lane
    uint8_t* tile0 = sOut;                                // dV, then dQ tile 0
    uint8_t* tile1 = sOut + TILE_BYTES;                   // dK, then dQ tile 1
    bool store_pending = false;                           // thread 0: TMA stores may still be reading the staging tiles
    // wait until the previous stores have read the staging tiles, let every thread see it
    auto tiles_reusable = [&]() {
      if (t == 0 && store_pending) tma_store_wait_read();
      asm volatile("bar.sync 4, 128;" ::: "memory");
    };
    auto tiles_written = [&]() {
      fence_async_smem();
      asm volatile("bar.sync 4, 128;" ::: "memory");
    };
    int ic = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++ic) {
      const int b = w / H, h = w - b * H;
      for (int kt = 0; kt < T; ++kt) {
        // every step's phase is waited in order (a parity wait may only name the current or the preceding phase); the
        // last one, step (kt, T-1), completes key tile kt
        for (int qt = 0; qt < T; ++qt) mbar_wait(&ctl->st_free, (ic * T * T + kt * T + qt) & 1);
        tc_fence_after();
        GVIT_TR(16);
        float v0[32], v1[32];
        tiles_reusable();
        tmem_ld32x2(tmem_lane_base(tdV, warp), tmem_lane_base(tdV, warp) + 32, v0, v1);
        stage_out64(tile0, t, v0, v1, 1.0f);
        tmem_ld32x2(tmem_lane_base(tdK, warp), tmem_lane_base(tdK, warp) + 32, v0, v1);
        tc_fence_before();
        mbar_arrive(&ctl->dvk_free);                      // both accumulators are in registers / staged
        stage_out64(tile1, t, v0, v1, scale);             // dS was formed without the softmax scale
        tiles_written();
        if (t == 0) {
          tma_store_3d(&tm_dqkv, tile0, (2 * H + h) * 64, kt * 128, b);   // dV; rows >= N are clipped
          tma_store_3d(&tm_dqkv, tile1, (H + h) * 64, kt * 128, b);       // dK
          tma_store_commit();
          store_pending = true;
        }
      }
      // dQ (final with the same step as the last key tile)
      {
        float v0[32], v1[32];
        tiles_reusable();
        tmem_ld32x2(tmem_lane_base(tdQ, warp), tmem_lane_base(tdQ, warp) + 32, v0, v1);
        stage_out64(tile0, t, v0, v1, scale);
        if (T > 1) {
          tmem_ld32x2(tmem_lane_base(tdQ + 64, warp), tmem_lane_base(tdQ + 64, warp) + 32, v0, v1);
          stage_out64(tile1, t, v0, v1, scale);
        }
        tc_fence_before();
        mbar_arrive(&ctl->dq_free);
        GVIT_TR(17);
        tiles_written();
        if (t == 0) {
          tma_store_3d(&tm_dqkv, tile0, h * 64, 0, b);
          if (T > 1) tma_store_3d(&tm_dqkv, tile1, h * 64, 128, b);
          tma_store_commit();
          store_pending = true;
        }
      }
    }
  } else if (warp >= 10) {
    // ------------------------------------------------------------------ helper warps: delta / lse ONE ITEM AHEAD
    // delta = rowsum(dO * O) and lse in log2 units for the next item, into the item-parity double buffer, while the
    // warpgroups are busy with the current item (as a warpgroup prologue it cost ~5000 idle cycles per item).
    // 8 lanes share one 128-byte row (coalesced); helper h owns query rows [128h, 128h+128).
    const int hw = warp - 10, sub = lane >> 3, ch = lane & 7;
    const int64_t rstride = (int64_t)H * 64;
    int ic = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++ic) {
      const int b = w / H, h = w - b * H;
      const int par = ic & 1;
      if (ic >= 2) mbar_wait(&ctl->wg_done, (ic - 2) & 1);  // the warpgroups are done reading this buffer (item ic-2)
      const __nv_bfloat16* obase = out + ((int64_t)b * N * H + h) * 64 + ch * 8;
      const __nv_bfloat16* dbase = dout + ((int64_t)b * N * H + h) * 64 + ch * 8;
      for (int rb = 0; rb < 4; ++rb) {
        uint4 ov[8], dv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = hw * 128 + rb * 32 + i * 4 + sub;
          ov[i] = dv[i] = make_uint4(0, 0, 0, 0);
          if (row < N) {
            ov[i] = *reinterpret_cast<const uint4*>(obase + row * rstride);
            dv[i] = *reinterpret_cast<const uint4*>(dbase + row * rstride);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162* oa = reinterpret_cast<const __nv_bfloat162*>(&ov[i]);
          const __nv_bfloat162* da = reinterpret_cast<const __nv_bfloat162*>(&dv[i]);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = __bfloat1622float2(oa[e]), y = __bfloat1622float2(da[e]);
            d = fmaf(x.x, y.x, d);
            d = fmaf(x.y, y.y, d);
          }
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (ch == 0) ctl->delta[par][hw * 128 + rb * 32 + i * 4 + sub] = -d;     // negated: see bwd_chunk
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = hw * 128 + j * 32 + lane;
        // negated (see bwd_chunk); -inf for padded queries: exp2(x - inf) = 0 masks them without a select
        ctl->lse2[par][q] = q < N ? -lse[((int64_t)b * H + h) * N + q] * LOG2E : __int_as_float(0xff800000);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->delta_ready[par]);
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = warp >> 2;                              // warpgroup == sub-tile == TMEM sub-buffer
    const int t = (warp & 3) * 32 + lane;                 // key row of the tile == TMEM lane
    const int tid = g * 128 + t;                          // 0..255
    const float sl2 = scale * LOG2E;
    const uint32_t lST = tmem_lane_base(tST + g * 64, warp), ldPT = tmem_lane_base(tdPT + g * 64, warp);
    uint8_t* dst_g = sdST + g * TILE_BYTES;
    int ic = 0, itc = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++ic) {
      const int par = ic & 1;
      mbar_wait(&ctl->delta_ready[par], (ic >> 1) & 1);     // delta / lse of this item: written by the helper warps
      GVIT_TR(10);
      for (int kt = 0; kt < T; ++kt) {
        const int key = kt * 128 + t;
        const bool kvalid = key < N;
        for (int qt = 0; qt < T; ++qt, ++itc) {
          const int nq = (min(128, N - qt * 128) + 15) & ~15;
          const int wg = sub_width(nq, g);
          const int q0 = qt * 128 + g * 64;
          mbar_wait(&ctl->s_full[g], itc & 1);
          tc_fence_after();
          GVIT_TR(11);
          for (int c0 = 0; c0 < wg; c0 += 32) {
            const float* l2p = &ctl->lse2[par][q0 + c0];
            const float* dlp = &ctl->delta[par][q0 + c0];
            if (wg - c0 >= 32) {
              float s[32], dp[32];
              tmem_ld32x2(lST + c0, ldPT + c0, s, dp);
              GVIT_TR(12);
              if (c0 == 0 && itc > 0) mbar_wait(&ctl->st_free, (itc - 1) & 1);   // staging of the previous step consumed
              GVIT_TR(14);
              bwd_chunk<32>(s, dp, l2p, dlp, sl2, lST, ldPT, dst_g, t, c0);
            } else {                                      // 16-column tail: never touch stale TMEM columns
              float s[16], dp[16];
              tmem_ld16x2(lST + c0, ldPT + c0, s, dp);
              if (c0 == 0 && itc > 0) mbar_wait(&ctl->st_free, (itc - 1) & 1);
              bwd_chunk<16>(s, dp, l2p, dlp, sl2, lST, ldPT, dst_g, t, c0);
            }
            GVIT_TR(13);
          }
          tmem_st_wait();
          fence_async_smem();
          tc_fence_before();
          mbar_arrive(&ctl->p_full[g]);
          GVIT_TR(15);
        }
      }
      mbar_arrive(&ctl->wg_done);                         // lse2 / delta of this item are no longer read
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding TMA stores of this thread (no-op for most)
  tc_fence_before();
  __syncthreads();
  GVIT_SPAN(1);
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace

GVIT_TRACE_SETTER(gvit_debug_set_trace_attn)

bool attn_fwd_tc_supported(int N, int dh) { return dh == 64 && N >= 1; }
bool attn_bwd_tc_supported(int N, int dh) { return dh == 64 && N >= 1; }   // N > 256: the two-pass kernel of attn_long_tc.cu

int attn_fwd_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, cudaStream_t st) {
  CUtensorMap tm;
  int rc = make_tmap_bf16_3d(&tm, qkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  if (N <= 256) {                                            // whole head per work item, persistent CTAs
    CUtensorMap tm_out;
    rc = make_tmap_bf16_3d(&tm_out, out, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
    if (rc != GVIT_OK) return rc;
    GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM));
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attn_fwd_tc2_kernel<<<grid, F2_THREADS, F2_SMEM, st>>>(tm, tm_out, N, H, items, scale, lse);
    GVIT_CHECK_LAUNCH();
    return GVIT_OK;
  }
  return attn_fwd_long_tc(qkv, B, N, H, scale, out, lse, st);   // N > 256: the persistent key-block kernel (attn_fwd_long_tc.cu)
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, float scale,
                float* delta_ws, void* dqkv, cudaStream_t st) {
  if (N > 256) return attn_bwd_long_tc(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st);
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_bf16_3d(&tm_qkv, qkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_do, dout, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  CUtensorMap tm_dqkv, tm_o;
  rc = make_tmap_bf16_3d(&tm_o, out, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_dqkv, dqkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  const int items = B * H;
  const int grid = items < num_sms() ? items : num_sms();
  (void)delta_ws;                                            // only the fp32-FMA path needs the global delta workspace
  attn_bwd_tc_kernel<<<grid, B2_THREADS, BWD_SMEM, st>>>(tm_qkv, tm_do, tm_dqkv, tm_o, N, H, items, scale,
                                                         static_cast<const __nv_bfloat16*>(out),
                                                         static_cast<const __nv_bfloat16*>(dout), lse);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

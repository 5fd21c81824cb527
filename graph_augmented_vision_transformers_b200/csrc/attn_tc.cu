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
// Forward (one CTA per 128-query tile x head x image, 2 CTAs/SM):
//   for each 128-key block j:  S_j = Q K_j^T  (tcgen05.mma -> TMEM cols [0,128))
//     softmax warps (thread <-> query row = TMEM lane): row max, p = exp2(s*c - m), bf16 P_j -> smem
//     O_j = P_j V_j (tcgen05.mma -> TMEM cols [128,192)), folded into fp32 registers with the online rescale.
// Backward (one CTA per head x image, N <= 256), key-major so every transposed product is a plain
//   K-major/MN-major operand:  S^T = K Q^T, dP^T = V dO^T  -> threads form P^T, dS^T (and dS, transposed
//   through shared memory) -> dV += P^T dO, dK += dS^T Q, dQ += dS K, all accumulated in TMEM.
//
// Warp roles: 0-3 softmax / epilogue, 4 TMA producer, 5 MMA issuer (+ TMEM allocation).
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr int TILE_BYTES = 128 * 128;   // [128 rows][64 bf16]
constexpr int THREADS = 192;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// write 32 consecutive bf16 (columns c0..c0+31 of row `row`) of a K-major operand made of [128][64] blocks
__device__ __forceinline__ void store_row32(uint8_t* tile, int row, int c0, const float (&p)[32]) {
  uint8_t* blk = tile + (c0 >> 6) * TILE_BYTES;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack_bf16(p[8 * q + 0], p[8 * q + 1]);
    w.y = pack_bf16(p[8 * q + 2], p[8 * q + 3]);
    w.z = pack_bf16(p[8 * q + 4], p[8 * q + 5]);
    w.w = pack_bf16(p[8 * q + 6], p[8 * q + 7]);
    *reinterpret_cast<uint4*>(blk + swz128(row, (c0 & 63) + 8 * q)) = w;
  }
}
__device__ __forceinline__ void store_out64(__nv_bfloat16* dst, const float (&o)[64], float mul) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 w;
    w.x = pack_bf16(o[8 * q + 0] * mul, o[8 * q + 1] * mul);
    w.y = pack_bf16(o[8 * q + 2] * mul, o[8 * q + 3] * mul);
    w.z = pack_bf16(o[8 * q + 4] * mul, o[8 * q + 5] * mul);
    w.w = pack_bf16(o[8 * q + 6] * mul, o[8 * q + 7] * mul);
    *reinterpret_cast<uint4*>(dst + 8 * q) = w;
  }
}

// =================================================================================================
// forward
// =================================================================================================
struct __align__(8) FwdCtrl {
  uint64_t q_full, kv_full[2], kv_empty[2], s_full, p_full, o_full;
  uint32_t tmem_base;
};
constexpr size_t FWD_SMEM = 1024 + 7 * TILE_BYTES + sizeof(FwdCtrl);   // Q, K x2, V x2, P x2 (two 64-key blocks)

__global__ void __launch_bounds__(THREADS, 2) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, int N,
                                                                 int H, float scale, __nv_bfloat16* __restrict__ out,
                                                                 float* __restrict__ lse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;
  uint8_t* sK = sm + TILE_BYTES;          // 2 stages
  uint8_t* sV = sm + 3 * TILE_BYTES;      // 2 stages
  uint8_t* sP = sm + 5 * TILE_BYTES;      // [128 rows][128 keys] as two 64-key blocks
  FwdCtrl* ctl = reinterpret_cast<FwdCtrl*>(sm + 7 * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (N + 127) / 128;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    mbar_init(&ctl->q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->kv_full[s], 1); mbar_init(&ctl->kv_empty[s], 1); }
    mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->p_full, 128);
    mbar_init(&ctl->o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&ctl->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&ctl->q_full, TILE_BYTES);
      tma_load_3d(sQ, &tm_qkv, h * 64, q0, b, &ctl->q_full);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(&ctl->kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->kv_full[s], 2 * TILE_BYTES);
        tma_load_3d(sK + s * TILE_BYTES, &tm_qkv, (H + h) * 64, j * 128, b, &ctl->kv_full[s]);
        tma_load_3d(sV + s * TILE_BYTES, &tm_qkv, (2 * H + h) * 64, j * 128, b, &ctl->kv_full[s]);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const int nj = (min(128, N - j * 128) + 15) & ~15;
        mbar_wait(&ctl->kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + s * TILE_BYTES);
        const uint32_t idesc = make_idesc(128, nj, false, false);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tS, make_sdesc(aQ + kk * 32), make_sdesc(aK + kk * 32), idesc, kk > 0);
        umma_commit(&ctl->s_full);
      };
      mbar_wait(&ctl->q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        const int nj = (min(128, N - j * 128) + 15) & ~15;
        mbar_wait(&ctl->p_full, j & 1);      // P_j in smem; S_j and O_{j-1} fully read by the softmax warps
        tc_fence_after();
        if (j + 1 < nblk) issue_s(j + 1);    // next scores overlap this block's P.V
        const uint32_t aV = smem_u32(sV + s * TILE_BYTES);
        const uint32_t idesc = make_idesc(128, 64, false, true);
        for (int ks = 0; ks < nj / 16; ++ks)
          umma_ss(tO, make_sdesc(aP + (ks >> 2) * TILE_BYTES + (ks & 3) * 32), make_sdesc(aV + ks * 2048), idesc, ks > 0);
        umma_commit(&ctl->o_full);
        umma_commit(&ctl->kv_empty[s]);
      }
    }
  } else {
    const int r = warp * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lS = tmem_lane_base(tS, warp), lO = tmem_lane_base(tO, warp);
    const float sl2 = scale * LOG2E;
    float m_run = -1e30f, l_run = 0.f;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int len = min(128, N - j * 128);
      const int nj = (len + 15) & ~15;
      mbar_wait(&ctl->s_full, j & 1);
      tc_fence_after();
      float mx = -1e30f;
      for (int c0 = 0; c0 < nj; c0 += 32) {
        float v[32];
        tmem_ld32(lS + c0, v);
#pragma unroll
        for (int t = 0; t < 32; ++t) mx = (c0 + t < len) ? fmaxf(mx, v[t]) : mx;
      }
      const float m_new = fmaxf(m_run, mx * sl2);
      float l_blk = 0.f;
      for (int c0 = 0; c0 < nj; c0 += 32) {
        float v[32];
        tmem_ld32(lS + c0, v);
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          v[t] = (c0 + t < len) ? ex2(fmaf(v[t], sl2, -m_new)) : 0.f;
          l_blk += v[t];
        }
        store_row32(sP, r, c0, v);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&ctl->p_full);
      const float corr = ex2(m_run - m_new);
      l_run = l_run * corr + l_blk;
      m_run = m_new;
      mbar_wait(&ctl->o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        float v[32];
        tmem_ld32(lO + hlf * 32, v);
#pragma unroll
        for (int t = 0; t < 32; ++t) o[hlf * 32 + t] = fmaf(o[hlf * 32 + t], corr, v[t]);
      }
    }
    const int q = q0 + r;
    if (q < N) {
      store_out64(out + (((int64_t)b * N + q) * H + h) * 64, o, 1.0f / l_run);
      lse[((int64_t)b * H + h) * N + q] = (m_run + log2f(l_run)) * LN2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 256);
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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  F2Ctrl* ctl = reinterpret_cast<F2Ctrl*>(sm + 2 * F2_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MT = (N + 127) >> 7;                 // 128-row query tiles (1 or 2) == 128-row key/value tiles
  const int NT = (N + 15) & ~15;                 // key extent of the MMAs

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
  const uint32_t tmem = ctl->tmem_base;

  if (warp == 8) {
    if (lane == 0) {
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
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(128, NT, false, false);
      const uint32_t idesc_o = make_idesc(128, 64, false, true);
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t aQ = smem_u32(sm + s * F2_STAGE_BYTES), aK = aQ + 2 * TILE_BYTES, aV = aQ + 4 * TILE_BYTES;
        mbar_wait(&ctl->full[s], (it >> 1) & 1);
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
        mbar_wait(&ctl->s_full[g], it & 1);
        tc_fence_after();
        float mx = -3.0e38f;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld32(tR + c0, v);
#pragma unroll
          for (int t = 0; t < 32; ++t) mx = (c0 + t < N) ? fmaxf(mx, v[t]) : mx;
        }
        const float msc = mx * sl2;
        float l = 0.f;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld32(tR + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const float p0 = (c0 + t < N) ? ex2(fmaf(v[t], sl2, -msc)) : 0.f;
            const float p1 = (c0 + t + 1 < N) ? ex2(fmaf(v[t + 1], sl2, -msc)) : 0.f;
            l += p0 + p1;
            pk[t >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16(tR + (c0 >> 1), pk);                     // in place: columns [c0/2, c0/2+16) were read already
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ctl->p_full[g]);

        mbar_wait(&ctl->o_full[g], it & 1);
        tc_fence_after();
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
        mbar_arrive(&ctl->t_free[g]);                        // TMEM region g may be overwritten by the next S_g
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
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// =================================================================================================
// backward (N <= 256)
// =================================================================================================
struct __align__(8) BwdCtrl {
  float lse2[256], delta[256];
  uint64_t qdo_full, kv_full, kv_empty, s_full, p_full, mma_done;
  uint32_t tmem_base;
};
// Q[2], dO[2], K, V tiles + P^T, dS^T (each [128][128] = 2 blocks)
constexpr size_t BWD_SMEM = 1024 + 10 * TILE_BYTES + sizeof(BwdCtrl);

__global__ void __launch_bounds__(THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                 const __grid_constant__ CUtensorMap tm_do, int N,
                                                                 int H, float scale,
                                                                 const __nv_bfloat16* __restrict__ out,
                                                                 const __nv_bfloat16* __restrict__ dout,
                                                                 const float* __restrict__ lse,
                                                                 float* __restrict__ delta_ws,
                                                                 __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;                        // 2 tiles
  uint8_t* sdO = sm + 2 * TILE_BYTES;      // 2 tiles
  uint8_t* sK = sm + 4 * TILE_BYTES;
  uint8_t* sV = sm + 5 * TILE_BYTES;
  uint8_t* sPT = sm + 6 * TILE_BYTES;      // [keys][q]   K-major A of dV
  uint8_t* sdST = sm + 8 * TILE_BYTES;     // [keys][q]   K-major A of dK and MN-major A of dQ
  BwdCtrl* ctl = reinterpret_cast<BwdCtrl*>(sm + 10 * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int T = (N + 127) / 128;           // tiles along queries == tiles along keys (1 or 2)

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    mbar_init(&ctl->qdo_full, 1);
    mbar_init(&ctl->kv_full, 1);
    mbar_init(&ctl->kv_empty, 1);
    mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->p_full, 128);
    mbar_init(&ctl->mma_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;
  const uint32_t tST = tmem, tdPT = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&ctl->qdo_full, 2 * T * TILE_BYTES);
      for (int t = 0; t < T; ++t) {
        tma_load_3d(sQ + t * TILE_BYTES, &tm_qkv, h * 64, t * 128, b, &ctl->qdo_full);
        tma_load_3d(sdO + t * TILE_BYTES, &tm_do, h * 64, t * 128, b, &ctl->qdo_full);
      }
      for (int kt = 0; kt < T; ++kt) {
        mbar_wait(&ctl->kv_empty, (kt & 1) ^ 1);
        mbar_expect_tx(&ctl->kv_full, 2 * TILE_BYTES);
        tma_load_3d(sK, &tm_qkv, (H + h) * 64, kt * 128, b, &ctl->kv_full);
        tma_load_3d(sV, &tm_qkv, (2 * H + h) * 64, kt * 128, b, &ctl->kv_full);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t aPT = smem_u32(sPT), adST = smem_u32(sdST);
      mbar_wait(&ctl->qdo_full, 0);
      int it = 0;
      for (int kt = 0; kt < T; ++kt) {
        const int nk = (min(128, N - kt * 128) + 15) & ~15;
        mbar_wait(&ctl->kv_full, kt & 1);
        for (int qt = 0; qt < T; ++qt, ++it) {
          const int nq = (min(128, N - qt * 128) + 15) & ~15;
          tc_fence_after();
          // S^T = K Q^T, dP^T = V dO^T   (M = keys, N = queries, K = head dim)
          const uint32_t idesc_s = make_idesc(128, nq, false, false);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tST, make_sdesc(aK + kk * 32), make_sdesc(aQ + qt * TILE_BYTES + kk * 32), idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tdPT, make_sdesc(aV + kk * 32), make_sdesc(adO + qt * TILE_BYTES + kk * 32), idesc_s, kk > 0);
          umma_commit(&ctl->s_full);
          mbar_wait(&ctl->p_full, it & 1);   // P^T, dS^T, dS staged in smem
          tc_fence_after();
          const uint32_t idesc_mn = make_idesc(128, 64, false, true);
          for (int ks = 0; ks < nq / 16; ++ks) {   // K = queries of this tile
            const uint32_t aoff = (ks >> 2) * TILE_BYTES + (ks & 3) * 32, boff = qt * TILE_BYTES + ks * 2048;
            umma_ss(tdV, make_sdesc(aPT + aoff), make_sdesc(adO + boff), idesc_mn, qt > 0 || ks > 0);
            umma_ss(tdK, make_sdesc(adST + aoff), make_sdesc(aQ + boff), idesc_mn, qt > 0 || ks > 0);
          }
          // dQ += dS K with dS = (dS^T)^T: the [keys][queries] tile is read as an MN-major A operand (M = queries
          // contiguous, two 64-query atoms LBO = one tile apart), so dS is never transposed through shared memory
          const uint32_t idesc_tt = make_idesc(128, 64, true, true);
          for (int ks = 0; ks < nk / 16; ++ks)     // K = keys of this tile
            umma_ss(tdQ + qt * 64, make_sdesc_lbo(adST + ks * 2048, TILE_BYTES), make_sdesc(aK + ks * 2048), idesc_tt,
                    kt > 0 || ks > 0);
          umma_commit(&ctl->mma_done);
          if (qt == T - 1) umma_commit(&ctl->kv_empty);
        }
      }
    }
  } else {
    const int t = warp * 32 + lane;              // key row of the tile (S^T lanes) / query row (dQ lanes)
    const float sl2 = scale * LOG2E;
    // prologue: delta = rowsum(dO * O), lse in log2 units
    for (int q = t; q < 256; q += 128) {
      float d = 0.f, l2 = 0.f;
      if (q < N) {
        const __nv_bfloat16* orow = out + (((int64_t)b * N + q) * H + h) * 64;
        const __nv_bfloat16* drow = dout + (((int64_t)b * N + q) * H + h) * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a[8], g[8];
          load8(orow + 8 * c, a);
          load8(drow + 8 * c, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(a[e], g[e], d);
        }
        l2 = lse[((int64_t)b * H + h) * N + q] * LOG2E;
        delta_ws[((int64_t)b * H + h) * N + q] = d;
      }
      ctl->delta[q] = d;
      ctl->lse2[q] = l2;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint32_t lST = tmem_lane_base(tST, warp), ldPT = tmem_lane_base(tdPT, warp);
    int it = 0;
    for (int kt = 0; kt < T; ++kt) {
      const int key = kt * 128 + t;
      const bool kvalid = key < N;
      for (int qt = 0; qt < T; ++qt, ++it) {
        const int nq = (min(128, N - qt * 128) + 15) & ~15;
        mbar_wait(&ctl->s_full, it & 1);
        tc_fence_after();
        for (int c0 = 0; c0 < nq; c0 += 32) {
          float s[32], dp[32];
          tmem_ld32(lST + c0, s);
          tmem_ld32(ldPT + c0, dp);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int q = qt * 128 + c0 + e;
            const float p = (kvalid && q < N) ? ex2(fmaf(s[e], sl2, -ctl->lse2[q])) : 0.f;
            s[e] = p;
            dp[e] = p * (dp[e] - ctl->delta[q]) * scale;
          }
          store_row32(sPT, t, c0, s);
          store_row32(sdST, t, c0, dp);
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(&ctl->p_full);
        mbar_wait(&ctl->mma_done, it & 1);       // P^T/dS^T/dS consumed; dV, dK, dQ updated
        tc_fence_after();
        if (qt == T - 1) {
          float v[64];
          tmem_ld32(tmem_lane_base(tdV, warp), *reinterpret_cast<float(*)[32]>(&v[0]));
          tmem_ld32(tmem_lane_base(tdV, warp) + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
          if (kvalid) store_out64(dqkv + ((((int64_t)b * N + key) * 3 + 2) * H + h) * 64, v, 1.0f);
          tmem_ld32(tmem_lane_base(tdK, warp), *reinterpret_cast<float(*)[32]>(&v[0]));
          tmem_ld32(tmem_lane_base(tdK, warp) + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
          if (kvalid) store_out64(dqkv + ((((int64_t)b * N + key) * 3 + 1) * H + h) * 64, v, 1.0f);
          tc_fence_before();                      // order these reads before the next tile's MMAs (via p_full)
        }
      }
    }
    for (int qt = 0; qt < T; ++qt) {
      float v[64];
      tmem_ld32(tmem_lane_base(tdQ, warp) + qt * 64, *reinterpret_cast<float(*)[32]>(&v[0]));
      tmem_ld32(tmem_lane_base(tdQ, warp) + qt * 64 + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
      const int q = qt * 128 + t;
      if (q < N) store_out64(dqkv + ((((int64_t)b * N + q) * 3 + 0) * H + h) * 64, v, 1.0f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

}  // namespace

bool attn_fwd_tc_supported(int N, int dh) { return dh == 64 && N >= 1; }
bool attn_bwd_tc_supported(int N, int dh) { return dh == 64 && N >= 1 && N <= 256; }

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
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  dim3 grid((N + 127) / 128, H, B);
  attn_fwd_tc_kernel<<<grid, THREADS, FWD_SMEM, st>>>(tm, N, H, scale, static_cast<__nv_bfloat16*>(out), lse);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, float scale,
                float* delta_ws, void* dqkv, cudaStream_t st) {
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_bf16_3d(&tm_qkv, qkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_do, dout, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  dim3 grid(H, B);
  attn_bwd_tc_kernel<<<grid, THREADS, BWD_SMEM, st>>>(tm_qkv, tm_do, N, H, scale, static_cast<const __nv_bfloat16*>(out),
                                                      static_cast<const __nv_bfloat16*>(dout), lse, delta_ws,
                                                      static_cast<__nv_bfloat16*>(dqkv));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

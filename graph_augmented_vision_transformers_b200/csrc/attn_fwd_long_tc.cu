// attn_fwd_long_tc.cu - fused softmax(QK^T * scale)V forward for N > 256 keys (ViT-L/16 at 384x384: N = 577), bf16,
// head_dim 64; replaces /root/reference/src/models/vit.py:64-69 at the sequence lengths of BASELINE configs[3].
//
// The first long-sequence kernel (attn_fwd_tc_kernel, attn_tc.cu) gave every 128-query tile its own short-lived CTA with ONE
// softmax warpgroup and a strictly serial S -> softmax -> P (through shared memory) -> PV -> rescale chain per 128-key
// block: 0.248 ms per layer at B = 32, N = 577 = 176 TF/s, a tenth of the tensor peak (profiles/README.md).  This one is the
// N <= 256 kernel (attn_fwd_tc2_kernel) carried over to key BLOCKS with an online softmax:
//   * persistent CTAs, one per SM; a work item is (image, head, PAIR of 128-query tiles); consecutive items are the query
//     pairs of one head, so a head's K / V come from HBM once and from L2 for the other pairs;
//   * the keys are cut into nkb = ceil(N / 256) equal blocks of KB <= 256 keys (3 x 208 at N = 577); K_j, V_j stream through
//     a 2-stage TMA ring, the two Q tiles of an item are double-buffered across items;
//   * per key block and query tile g:  S_g = Q_g K_j^T -> TMEM region g;  softmax warpgroup g (thread <-> query row = TMEM
//     lane): block max, running max / sum (online softmax), p = exp2(..) written back IN PLACE over S_g as packed bf16;
//     O_g(j) = P_g V_j with P read straight from TMEM; the warpgroup folds O_g(j) into its fp32 registers with the
//     rescale factor exp2(m_old - m_new);
//   * the MMA warp issues in the order PV_0(u), S_0(u+1), PV_1(u), S_1(u+1): tile 0's tensor work runs under tile 1's
//     softmax and vice versa (the two warpgroups settle half a period apart), so the MUFU pipe - the bound of a head_dim-64
//     softmax, one MUFU.EX2 per score at 4 lanes/clk/SMSP, every 4th exponential emulated on the FMA pipe - stays busy;
//   * O_g / l -> bf16 -> the (dead) Q_g tile -> one TMA tile store.
// TMEM region g (256 of the 512 columns): S_g at [0, KB) -> P_g at [0, KB/2), O_g(j) at [128, 192).
// Warp roles: 0-3 / 4-7 softmax warpgroups of query tile 0 / 1, 8 TMA producer, 9 MMA issuer + TMEM owner.
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr int TILE_BYTES = 128 * 128;              // [128 rows][64 bf16]
constexpr int L_THREADS = 320;
constexpr int L_Q_BYTES = 2 * TILE_BYTES;          // the two query tiles of an item
constexpr int L_KV_BYTES = 4 * TILE_BYTES;         // K block [256][64] | V block [256][64]

struct __align__(8) LCtrl {
  uint64_t q_full[2], q_empty[2], kv_full[2], kv_empty[2], s_full[2], p_full[2], o_full[2], t_free[2];
  uint32_t tmem_base;
};
constexpr size_t L_SMEM = 2 * L_Q_BYTES + 2 * L_KV_BYTES + sizeof(LCtrl);

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA / ALU pipes: Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5,
// far below the bf16 resolution of the probabilities it feeds); x <= 0 expected, x < -125 clamped
__device__ __forceinline__ float ex2_emul(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;
  const float f = x - (xr - 12582912.0f);
  float p = fmaf(0.05517164617776871f, f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(L_THREADS, 1) attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                     const __grid_constant__ CUtensorMap tm_out, int N, int H,
                                                                     int QP, int nkb, int KB, int items, float scale,
                                                                     float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  uint8_t* sQ = sm;                                  // 2 buffers x 2 tiles
  uint8_t* sKV = sm + 2 * L_Q_BYTES;                 // 2 stages x (K 2 tiles | V 2 tiles)
  LCtrl* ctl = reinterpret_cast<LCtrl*>(sKV + 2 * L_KV_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_out);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->q_full[s], 1);
      mbar_init(&ctl->q_empty[s], 2);                // one arrival per softmax warpgroup
      mbar_init(&ctl->kv_full[s], 1);
      mbar_init(&ctl->kv_empty[s], 1);
      mbar_init(&ctl->s_full[s], 1);
      mbar_init(&ctl->p_full[s], 128);
      mbar_init(&ctl->o_full[s], 1);
      mbar_init(&ctl->t_free[s], 128);
    }
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
      uint32_t n = 0, u = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++n) {
        const int qp = w % QP, bh = w / QP, b = bh / H, h = bh - b * H;
        const uint32_t qb = n & 1;
        mbar_wait(&ctl->q_empty[qb], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->q_full[qb], (uint32_t)L_Q_BYTES);
        tma_load_3d(sQ + qb * L_Q_BYTES, &tm_qkv, h * 64, qp * 256, b, &ctl->q_full[qb]);                 // rows >= N: zeros
        tma_load_3d(sQ + qb * L_Q_BYTES + TILE_BYTES, &tm_qkv, h * 64, qp * 256 + 128, b, &ctl->q_full[qb]);
        for (int j = 0; j < nkb; ++j, ++u) {
          const uint32_t s = u & 1;
          mbar_wait(&ctl->kv_empty[s], ((u >> 1) & 1) ^ 1);
          mbar_expect_tx(&ctl->kv_full[s], (uint32_t)L_KV_BYTES);
          uint8_t* st = sKV + s * L_KV_BYTES;
          tma_load_3d(st, &tm_qkv, (H + h) * 64, j * KB, b, &ctl->kv_full[s]);
          tma_load_3d(st + TILE_BYTES, &tm_qkv, (H + h) * 64, j * KB + 128, b, &ctl->kv_full[s]);
          tma_load_3d(st + 2 * TILE_BYTES, &tm_qkv, (2 * H + h) * 64, j * KB, b, &ctl->kv_full[s]);
          tma_load_3d(st + 3 * TILE_BYTES, &tm_qkv, (2 * H + h) * 64, j * KB + 128, b, &ctl->kv_full[s]);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(128, KB, false, false);
      const uint32_t idesc_o = make_idesc(128, 64, false, true);
      const uint32_t aQ0 = smem_u32(sQ), aKV0 = smem_u32(sKV);
      // flat sequence of (item, key block) steps u; `valid1` = the item has a second query tile
      auto issue_s = [&](uint32_t n, uint32_t u, int g) {          // S_g(u) = Q_g K_j^T
        const uint32_t aQ = aQ0 + (n & 1) * L_Q_BYTES + g * TILE_BYTES, aK = aKV0 + (u & 1) * L_KV_BYTES;
        mbar_wait(&ctl->t_free[g], (u & 1) ^ 1);                   // region g: O_g(u-1) read, S_g(u-1) long consumed
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + g * 256, make_sdesc(aQ + kk * 32), make_sdesc(aK + kk * 32), idesc_s, kk > 0);
        umma_commit(&ctl->s_full[g]);
      };
      auto issue_pv = [&](uint32_t u, int g) {                     // O_g(u) = P_g V_j, P read from TMEM
        const uint32_t aV = aKV0 + (u & 1) * L_KV_BYTES + 2 * TILE_BYTES;
        mbar_wait(&ctl->p_full[g], u & 1);
        tc_fence_after();
        for (int ks = 0; ks < KB / 16; ++ks)
          umma_ts(tmem + g * 256 + 128, tmem + g * 256 + ks * 8, make_sdesc(aV + ks * 2048), idesc_o, ks > 0);
        umma_commit(&ctl->o_full[g]);
      };
      // NOTE: a query tile that lies wholly past N still runs (its Q rows are zero-filled, nothing of it is stored): the
      // per-region barrier phases then advance uniformly with u for both regions
      uint32_t u = 0;
      const int my_items = blockIdx.x < items ? (items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const uint32_t U = (uint32_t)my_items * nkb;
      if (U > 0) {
        mbar_wait(&ctl->q_full[0], 0);
        mbar_wait(&ctl->kv_full[0], 0);
        tc_fence_after();
        issue_s(0, 0, 0);
        issue_s(0, 0, 1);
      }
      for (; u < U; ++u) {
        const uint32_t un = u + 1, nn = un / nkb;                  // next step and its item
        const bool more = un < U;
        issue_pv(u, 0);
        if (more) {
          if (un % nkb == 0) mbar_wait(&ctl->q_full[nn & 1], (nn >> 1) & 1);    // first key block of the next item: its Q
          mbar_wait(&ctl->kv_full[un & 1], (un >> 1) & 1);
          tc_fence_after();
          issue_s(nn, un, 0);
        }
        issue_pv(u, 1);
        umma_commit(&ctl->kv_empty[u & 1]);                          // K_j / V_j of step u are consumed
        if (more) issue_s(nn, un, 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;                           // row inside the tile == TMEM lane
    const uint32_t tR = tmem_lane_base(tmem + g * 256, warp);
    const float sl2 = scale * LOG2E;
    uint32_t n = 0, u = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++n) {
      const int qp = w % QP, bh = w / QP, b = bh / H, h = bh - b * H;
      const int q = qp * 256 + g * 128 + r;                         // query row
      float m_run = -1.0e30f, l_run = 0.f;
      float o[64];
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] = 0.f;
      for (int j = 0; j < nkb; ++j, ++u) {
        const int len = min(KB, N - j * KB);                        // valid keys of this block (>= 1)
        mbar_wait(&ctl->s_full[g], u & 1);
        tc_fence_after();
        float mx = -3.0e38f;
        for (int c0 = 0; c0 < KB; c0 += 32) {
          if (c0 >= len) break;
          float v[32];
          tmem_ld32(tR + c0, v);
          if (c0 + 32 <= len) {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = fmaxf(mx, v[t]);
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = (c0 + t < len) ? fmaxf(mx, v[t]) : mx;
          }
        }
        const float m_new = fmaxf(m_run, mx * sl2);
        float l_blk = 0.f;
        for (int c0 = 0; c0 < KB; c0 += 32) {                       // KB is a multiple of 16: the last chunk may be half
          float v[32];
          uint32_t pk[16];
          if (c0 + 32 <= len) {                                     // full chunk: no masks, every 4th exp2 off the MUFU pipe
            tmem_ld32(tR + c0, v);
#pragma unroll
            for (int t = 0; t < 32; t += 4) {
              const float p0 = ex2(fmaf(v[t], sl2, -m_new));
              const float p1 = ex2(fmaf(v[t + 1], sl2, -m_new));
              const float p2 = ex2(fmaf(v[t + 2], sl2, -m_new));
              const float p3 = ex2_emul(fmaf(v[t + 3], sl2, -m_new));
              l_blk += (p0 + p1) + (p2 + p3);
              pk[t >> 1] = pack_bf16(p0, p1);
              pk[(t >> 1) + 1] = pack_bf16(p2, p3);
            }
          } else if (c0 < len) {
            tmem_ld32(tR + c0, v);
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
              const float p0 = (c0 + t < len) ? ex2(fmaf(v[t], sl2, -m_new)) : 0.f;
              const float p1 = (c0 + t + 1 < len) ? ex2(fmaf(v[t + 1], sl2, -m_new)) : 0.f;
              l_blk += p0 + p1;
              pk[t >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[t] = 0u;                // keys past N: P = 0
          }
          if (c0 + 32 <= KB) {
            tmem_st16(tR + (c0 >> 1), pk);                          // in place: columns [c0/2, c0/2+16) were read already
          } else {
            uint32_t pk8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) pk8[t] = pk[t];
            tmem_st8(tR + (c0 >> 1), pk8);                          // half chunk (16 keys)
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ctl->p_full[g]);
        const float corr = ex2(m_run - m_new);
        l_run = fmaf(l_run, corr, l_blk);
        m_run = m_new;
        mbar_wait(&ctl->o_full[g], u & 1);
        tc_fence_after();
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float v[32];
          tmem_ld32(tR + 128 + hlf * 32, v);
#pragma unroll
          for (int t = 0; t < 32; ++t) o[hlf * 32 + t] = fmaf(o[hlf * 32 + t], corr, v[t]);
        }
        tc_fence_before();
        mbar_arrive(&ctl->t_free[g]);                               // region g may be overwritten by the next S_g
      }
      // ---- epilogue: O / l -> bf16 -> the Q_g tile of this item (dead: every S_g of the item has completed) -> TMA store
      const float inv = 1.0f / l_run;
      uint8_t* so = sQ + (n & 1) * L_Q_BYTES + g * TILE_BYTES;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint4 o4;
        o4.x = pack_bf16(o[8 * c8 + 0] * inv, o[8 * c8 + 1] * inv);
        o4.y = pack_bf16(o[8 * c8 + 2] * inv, o[8 * c8 + 3] * inv);
        o4.z = pack_bf16(o[8 * c8 + 4] * inv, o[8 * c8 + 5] * inv);
        o4.w = pack_bf16(o[8 * c8 + 6] * inv, o[8 * c8 + 7] * inv);
        *reinterpret_cast<uint4*>(so + swz128(r, 8 * c8)) = o4;
      }
      if (q < N) lse[((int64_t)b * H + h) * N + q] = (m_run + log2f(l_run)) * LN2;
      fence_async_smem();                                           // generic-proxy tile writes -> visible to the TMA store
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if ((warp & 3) == 0 && lane == 0) {
        if (qp * 256 + g * 128 < N) {
          tma_store_3d(&tm_out, so, h * 64, qp * 256 + g * 128, b);  // rows >= N are clipped by the TMA unit
          tma_store_commit();
          tma_store_wait_read();
        }
        mbar_arrive(&ctl->q_empty[n & 1]);                          // this item's Q buffer (and O staging) may be refilled
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");         // outstanding TMA stores of this thread (no-op for most)
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace

int attn_fwd_long_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, cudaStream_t st) {
  CUtensorMap tm, tm_out;
  int rc = make_tmap_bf16_3d(&tm, qkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_out, out, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  const int nkb = (N + 255) / 256;                                  // key blocks ...
  const int KB = (((N + nkb - 1) / nkb) + 15) & ~15;                // ... of equal size, a multiple of 16, <= 256
  const int QP = ((N + 127) / 128 + 1) / 2;                         // pairs of 128-query tiles
  const int64_t items64 = (int64_t)B * H * QP;
  GVIT_REQUIRE(items64 < (1LL << 30), GVIT_ERR_SHAPE, "attn_fwd: B * H * query pairs = %lld is too large", (long long)items64);
  const int items = (int)items64;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM));
  const int grid = items < num_sms() ? items : num_sms();
  attn_fwd_long_kernel<<<grid, L_THREADS, L_SMEM, st>>>(tm, tm_out, N, H, QP, nkb, KB, items, scale, lse);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

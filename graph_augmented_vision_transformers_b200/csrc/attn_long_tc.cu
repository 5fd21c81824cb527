// attn_long_tc.cu - attention backward (vit.py:64-69) for MORE than 256 tokens on tcgen05: ViT-L/16 at 384x384 has
// N = 577 (BASELINE configs[3]).  attn_tc.cu's backward keeps a whole head in TMEM / shared memory (N <= 256); beyond
// that the round-1 library fell back to the exact-fp32-FMA kernels (~30 TFLOP/s).  Here the backward is two passes of
// ONE warp-specialised kernel over 128 x 128 score tiles, each pass owning its accumulators (no atomics, deterministic):
//
//   pass 0 (key-major, CTA = key tile r of one head):    for every query tile c
//        S^T = K_r Q_c^T, dP^T = V_r dO_c^T               (tcgen05, fp32 in TMEM)
//        P^T = exp2(S^T scale log2e - lse_q), dS^T = P^T (dP^T - delta_q)   (thread <-> key row; written back IN PLACE as
//                                                          packed bf16: the A operands of the next two products)
//        dV_r += P^T dO_c,  dK_r += dS^T Q_c               (A from TMEM, B = the same column tiles read MN-major)
//   pass 1 (query-major, CTA = query tile r):              for every key tile c
//        S = Q_r K_c^T, dP = dO_r V_c^T,  dS = P (dP - delta_row),  dQ_r += dS K_c
//
// delta = rowsum(dO * O) comes from a small streaming kernel into the caller's workspace (the ABI's delta_ws).  S and dP
// are recomputed in pass 1 (7 GEMMs instead of 5) - the price of keeping dQ out of global atomics.
// Tile conventions are tc.cuh's: [128][64] bf16, 128-byte swizzled rows, K-major or MN-major by descriptor; q / k / v are
// read in place from the packed (B, N, 3, H, 64) projection output (vit.py:59) through one 3-D tensor map.
#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f;
constexpr int TILE_BYTES = 128 * 128;          // [128 rows][64 bf16]
constexpr int L_THREADS = 320;                 // warps 0-7 softmax / epilogue, warp 8 TMA producer, warp 9 MMA issuer
constexpr int L_STAGES = 3;                    // column-tile pairs in flight

struct __align__(16) LCtrl {
  float nl[2][128], dl[2][128];                // pass 0: negated lse * log2e and negated delta of the current query tile
  uint64_t row_full, col_full[L_STAGES], col_empty[L_STAGES], s_full, p_full, out_full;
  uint32_t tmem_base;
};
constexpr size_t L_SMEM = (size_t)(2 + 2 * L_STAGES) * TILE_BYTES + sizeof(LCtrl);

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// delta[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d]; 8 lanes share one 128-byte head row (coalesced)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                                         int64_t rows, int H, int N, float* __restrict__ delta) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // (token, head) pair
  const int ch = threadIdx.x & 7;
  float d = 0.f;
  const bool on = g < rows * H;
  if (on) {
    const uint4 ov = *reinterpret_cast<const uint4*>(out + g * 64 + ch * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(dout + g * 64 + ch * 8);
    const __nv_bfloat162* oa = reinterpret_cast<const __nv_bfloat162*>(&ov);
    const __nv_bfloat162* da = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(oa[e]), y = __bfloat1622float2(da[e]);
      d = fmaf(x.x, y.x, d);
      d = fmaf(x.y, y.y, d);
    }
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (on && ch == 0) {
    const int64_t tok = g / H;
    const int h = (int)(g - tok * H);
    const int64_t b = tok / N;
    const int q = (int)(tok - b * N);
    delta[(b * H + h) * N + q] = d;
  }
}

// two 32-column (16-column) TMEM reads of this thread's lane, one wait
__device__ __forceinline__ void ld2(uint32_t ta, uint32_t tb, float (&a)[32], float (&b)[32]) {
  tmem_ld32(ta, a);
  tmem_ld32(tb, b);
}
__device__ __forceinline__ void ld2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  tmem_ld16(ta, a);
  tmem_ld16(tb, b);
}
__device__ __forceinline__ void st_packed(uint32_t t, const uint32_t (&r)[16]) { tmem_st16(t, r); }
__device__ __forceinline__ void st_packed(uint32_t t, const uint32_t (&r)[8]) { tmem_st8(t, r); }

// P = exp2(S sl2 + nl), dS = P (dP + dl) for W consecutive columns of this thread's row; both written back IN PLACE as packed
// bf16 (the A operands of the accumulating products).  nl / dl are the NEGATED lse * log2e / delta of each column (pass 0)
// or of the row (pass 1); -inf masks a padded query exactly (p = 0).
template <int W, bool PER_COL>
__device__ __forceinline__ void chunk(uint32_t tS, uint32_t tdP, uint32_t tPk, uint32_t tdSk, const float* nlc, const float* dlc,
                                      float nlr, float dlr, float sl2) {
  float s[W], dp[W];
  ld2(tS, tdP, s, dp);
  uint32_t pk[W / 2], dk[W / 2];
#pragma unroll
  for (int e = 0; e < W; ++e) {
    const float nl = PER_COL ? nlc[e] : nlr, dl = PER_COL ? dlc[e] : dlr;
    s[e] = ex2(fmaf(s[e], sl2, nl));
    dp[e] = s[e] * (dp[e] + dl);
  }
#pragma unroll
  for (int e = 0; e < W; e += 2) {
    pk[e >> 1] = pack_bf16(s[e], s[e + 1]);
    dk[e >> 1] = pack_bf16(dp[e], dp[e + 1]);
  }
  st_packed(tPk, pk);
  st_packed(tdSk, dk);
}

// 64 fp32 -> bf16 into row `row` of a swizzled [128][64] staging tile (the layout a SWIZZLE_128B TMA store reads)
__device__ __forceinline__ void stage64(uint8_t* tile, int row, const float (&a)[32], const float (&b)[32], float mul) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack_bf16(a[8 * q + 0] * mul, a[8 * q + 1] * mul); w.y = pack_bf16(a[8 * q + 2] * mul, a[8 * q + 3] * mul);
    w.z = pack_bf16(a[8 * q + 4] * mul, a[8 * q + 5] * mul); w.w = pack_bf16(a[8 * q + 6] * mul, a[8 * q + 7] * mul);
    *reinterpret_cast<uint4*>(tile + swz128(row, 8 * q)) = w;
    w.x = pack_bf16(b[8 * q + 0] * mul, b[8 * q + 1] * mul); w.y = pack_bf16(b[8 * q + 2] * mul, b[8 * q + 3] * mul);
    w.z = pack_bf16(b[8 * q + 4] * mul, b[8 * q + 5] * mul); w.w = pack_bf16(b[8 * q + 6] * mul, b[8 * q + 7] * mul);
    *reinterpret_cast<uint4*>(tile + swz128(row, 32 + 8 * q)) = w;
  }
}

template <int PASS>
__global__ void __launch_bounds__(L_THREADS, 1) attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                     const __grid_constant__ CUtensorMap tm_do,
                                                                     const __grid_constant__ CUtensorMap tm_dqkv, int N, int H,
                                                                     float scale, const float* __restrict__ lse,
                                                                     const float* __restrict__ delta) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sRow = smem_raw;                                 // row-side tiles: pass 0 K_r | V_r, pass 1 Q_r | dO_r
  if ((smem_u32(sRow) & 1023u) != 0) __trap();
  uint8_t* sCol = sRow + 2 * TILE_BYTES;                    // L_STAGES x { first | second column tile }
  LCtrl* ctl = reinterpret_cast<LCtrl*>(sCol + 2 * L_STAGES * TILE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = (N + 127) / 128;
  // columns of the packed projection: q at h*64, k at (H+h)*64, v at (2H+h)*64 (vit.py:59-61)
  const int cq = h * 64, ck = (H + h) * 64, cv = (2 * H + h) * 64;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dqkv);
    mbar_init(&ctl->row_full, 1);
    for (int s = 0; s < L_STAGES; ++s) { mbar_init(&ctl->col_full[s], 1); mbar_init(&ctl->col_empty[s], 1); }
    mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->p_full, 256);
    mbar_init(&ctl->out_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
  const uint32_t tS = tmem, tdP = tmem + 128, tO1 = tmem + 256, tO2 = tmem + 320;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_expect_tx(&ctl->row_full, 2 * TILE_BYTES);
      if (PASS == 0) {
        tma_load_3d(sRow, &tm_qkv, ck, r0, b, &ctl->row_full);
        tma_load_3d(sRow + TILE_BYTES, &tm_qkv, cv, r0, b, &ctl->row_full);
      } else {
        tma_load_3d(sRow, &tm_qkv, cq, r0, b, &ctl->row_full);
        tma_load_3d(sRow + TILE_BYTES, &tm_do, cq, r0, b, &ctl->row_full);
      }
      for (int j = 0; j < T; ++j) {
        const int s = j % L_STAGES;
        mbar_wait(&ctl->col_empty[s], ((j / L_STAGES) & 1) ^ 1);
        mbar_expect_tx(&ctl->col_full[s], 2 * TILE_BYTES);
        uint8_t* dst = sCol + (size_t)(2 * s) * TILE_BYTES;
        if (PASS == 0) {
          tma_load_3d(dst, &tm_qkv, cq, j * 128, b, &ctl->col_full[s]);
          tma_load_3d(dst + TILE_BYTES, &tm_do, cq, j * 128, b, &ctl->col_full[s]);
        } else {
          tma_load_3d(dst, &tm_qkv, ck, j * 128, b, &ctl->col_full[s]);
          tma_load_3d(dst + TILE_BYTES, &tm_qkv, cv, j * 128, b, &ctl->col_full[s]);
        }
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t aRow = smem_u32(sRow), aCol = smem_u32(sCol);
      const uint32_t idesc2 = make_idesc(128, 64, false, true);          // A from TMEM, B = column tile read MN-major
      auto mma1 = [&](int j) {                                           // both score-shaped products of step j
        const int s = j % L_STAGES;
        const int nj = (min(128, N - j * 128) + 15) & ~15;
        mbar_wait(&ctl->col_full[s], (j / L_STAGES) & 1);
        tc_fence_after();
        const uint32_t idesc1 = make_idesc(128, nj, false, false);
        const uint32_t c1 = aCol + (2 * s) * TILE_BYTES, c2 = c1 + TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tS, make_sdesc(aRow + kk * 32), make_sdesc(c1 + kk * 32), idesc1, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tdP, make_sdesc(aRow + TILE_BYTES + kk * 32), make_sdesc(c2 + kk * 32), idesc1, kk > 0);
        umma_commit(&ctl->s_full);
      };
      mbar_wait(&ctl->row_full, 0);
      mma1(0);
      for (int j = 0; j < T; ++j) {
        const int s = j % L_STAGES;
        const int nj = (min(128, N - j * 128) + 15) & ~15;
        mbar_wait(&ctl->p_full, j & 1);                                  // P / dS of step j are packed in TMEM
        tc_fence_after();
        const uint32_t c1 = aCol + (2 * s) * TILE_BYTES, c2 = c1 + TILE_BYTES;
        for (int ks = 0; ks < nj / 16; ++ks) {                           // K = the 16 columns [16 ks, 16 ks + 16) of the score tile
          const uint32_t aoff = (ks >> 2) * 64 + (ks & 3) * 8;           // packed operands: 8 columns per slice, 64-column halves
          const bool acc = j > 0 || ks > 0;
          if (PASS == 0) {
            umma_ts(tO1, tS + aoff, make_sdesc(c2 + ks * 2048), idesc2, acc);    // dV += P^T dO_c
            umma_ts(tO2, tdP + aoff, make_sdesc(c1 + ks * 2048), idesc2, acc);   // dK += dS^T Q_c
          } else {
            umma_ts(tO1, tdP + aoff, make_sdesc(c1 + ks * 2048), idesc2, acc);   // dQ += dS K_c
          }
        }
        umma_commit(&ctl->col_empty[s]);
        if (j + 1 < T) mma1(j + 1);                                      // in-order pipe: runs after the products that read P / dS
      }
      umma_commit(&ctl->out_full);
    }
  } else {
    // ---------------------------------------------------------------- softmax-backward warps (thread <-> tile row = TMEM lane)
    const int hc = warp >> 2;                                            // column half [64 hc, 64 hc + 64) of every score tile
    const int row = (warp & 3) * 32 + lane;
    const int tid = threadIdx.x;                                         // 0..255
    const float sl2 = scale * LOG2E;
    const uint32_t lS = tmem_lane_base(tS, warp), ldP = tmem_lane_base(tdP, warp);
    const float ninf = __int_as_float(0xff800000);
    const int64_t lbase = ((int64_t)b * H + h) * N;
    float nlr = 0.f, dlr = 0.f;
    if (PASS == 1) {
      const int q = r0 + row;
      nlr = q < N ? -lse[lbase + q] * LOG2E : ninf;
      dlr = q < N ? -delta[lbase + q] : 0.f;
    }
    for (int j = 0; j < T; ++j) {
      const int nj = (min(128, N - j * 128) + 15) & ~15;
      const int par = j & 1;
      if (PASS == 0) {
        if (tid < 128) {
          const int q = j * 128 + tid;
          ctl->nl[par][tid] = q < N ? -lse[lbase + q] * LOG2E : ninf;
          ctl->dl[par][tid] = q < N ? -delta[lbase + q] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&ctl->s_full, par);
      tc_fence_after();
      for (int c0 = hc * 64; c0 < hc * 64 + 64 && c0 < nj; c0 += 32) {
        const uint32_t pk = hc * 64 + ((c0 - hc * 64) >> 1);             // packed home of these columns
        if (nj - c0 >= 32) chunk<32, PASS == 0>(lS + c0, ldP + c0, lS + pk, ldP + pk, &ctl->nl[par][c0], &ctl->dl[par][c0], nlr, dlr, sl2);
        else chunk<16, PASS == 0>(lS + c0, ldP + c0, lS + pk, ldP + pk, &ctl->nl[par][c0], &ctl->dl[par][c0], nlr, dlr, sl2);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ctl->p_full);
    }
    // ---- epilogue: accumulators -> bf16 -> the (dead) row tiles -> TMA tile stores (rows >= N are clipped) ----
    mbar_wait(&ctl->out_full, 0);
    tc_fence_after();
    if (hc == 0 || PASS == 0) {
      float v0[32], v1[32];
      const uint32_t tO = tmem_lane_base(hc == 0 ? tO1 : tO2, warp);
      tmem_ld32(tO, v0);
      tmem_ld32(tO + 32, v1);
      // dV as is; dK and dQ carry the softmax scale (dS was formed without it)
      stage64(sRow + hc * TILE_BYTES, row, v0, v1, (PASS == 0 && hc == 0) ? 1.0f : scale);
    }
    fence_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) {
      if (PASS == 0) {
        tma_store_3d(&tm_dqkv, sRow, cv, r0, b);
        tma_store_3d(&tm_dqkv, sRow + TILE_BYTES, ck, r0, b);
      } else {
        tma_store_3d(&tm_dqkv, sRow, cq, r0, b);
      }
      tma_store_commit();
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace

int attn_bwd_long_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, float scale,
                     float* delta_ws, void* dqkv, cudaStream_t st) {
  CUtensorMap tm_qkv, tm_do, tm_dqkv;
  int rc = make_tmap_bf16_3d(&tm_qkv, qkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_do, dout, (uint64_t)H * 64, N, B, (uint64_t)H * 64, (uint64_t)N * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_dqkv, dqkv, (uint64_t)3 * H * 64, N, B, (uint64_t)3 * H * 64, (uint64_t)N * 3 * H * 64, 128);
  if (rc != GVIT_OK) return rc;
  const int64_t rows = (int64_t)B * N;
  attn_delta_kernel<<<(unsigned)((rows * H * 8 + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                                           static_cast<const __nv_bfloat16*>(dout), rows, H, N, delta_ws);
  GVIT_CHECK_LAUNCH();
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_long_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM));
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_long_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM));
  dim3 grid((N + 127) / 128, H, B);
  attn_bwd_long_kernel<0><<<grid, L_THREADS, L_SMEM, st>>>(tm_qkv, tm_do, tm_dqkv, N, H, scale, lse, delta_ws);
  GVIT_CHECK_LAUNCH();
  attn_bwd_long_kernel<1><<<grid, L_THREADS, L_SMEM, st>>>(tm_qkv, tm_do, tm_dqkv, N, H, scale, lse, delta_ws);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

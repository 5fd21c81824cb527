// mlp_tc.cu - Linear layers of the block with their whole epilogue in one kernel:
//
//   MODE 0, fc1 of the Mlp:
//     u   = x W^T + b                         (vit.py:90, nn.Linear)            -> saved for the backward, bf16
//     out = dropout(gelu(u), p)               (vit.py:91-92, nn.GELU, nn.Dropout) -> bf16 + 1-bit keep mask
//   MODE 1, attention output projection:
//     out = resid + dropout(x W^T + b, p)     (vit.py:70-71 proj + proj_drop, residual add of vit.py:117)
//
//   MODE 2, backward of fc2 -> (drop, GELU): the input gradient of fc2 fused with the backward of vit.py:91-92:
//     du  = (dout W2) * keep / (1 - p) * gelu'(u)     and per-tile column sums of du (fc1's bias gradient, summed by a
//     second tiny kernel in a fixed order).  Here the B operand is W2 itself ((out, in) row-major = [K][N] with N
//     contiguous): MN-major, four [64 k][64 n] atoms per stage.
//
// (The text below describes MODE 0; the other modes share the main loop and swap the epilogue.)
//
// As a library GEMM plus the gvit_gelu_dropout_fwd pass this is 173 us + 163 us at B = 256 (M = 50432, N = 3072,
// K = 768): the elementwise pass is issue-bound ALU work (GELU + Philox) over 640 MB.  Here that ALU work runs in the
// epilogue warps of a persistent tcgen05 GEMM, under the tensor pipe's shadow: the accumulator tile (128 x 256 fp32) is
// double-buffered in TMEM, so the MMAs of tile i+1 overlap bias + GELU + dropout + the two stores of tile i, and the
// pre-activation never makes the extra HBM round trip.
//
// Tiles: 256 x 256 per CTA PAIR, K in 64-deep slabs: `tcgen05.mma.cta_group::2` (M = 256, N = 256, K = 16) issued by the
// leader CTA of a 2-CTA cluster.  Each CTA stages its own 128 rows of x and its own 128 of the tile's 256 W rows (16 + 16 KB
// per stage, six stages); completion bytes of both CTAs land on the leader's full barrier, ring slots and accumulators are
// released to both CTAs by multicast tcgen05.commit - the same main loop as gemm2_tc.cu.  Each CTA's 128 x 256 fp32 half of
// the accumulator is double-buffered in its own TMEM (2 x 256 columns).
// History: un-paired 128 x 256 tiles ran the main loop at 1.07 PF/s (48 KB of operands per stage per CTA through L2 ->
// shared memory); W multicast across a 2-CTA cluster cut the L2 side to 32 KB but not the shared-memory fill (0.29 ms for
// fc1 at B = 256); pairing the MMA itself halves both.
// Tile order: consecutive tile ids share the x row block (the 12 column tiles of one row block run on neighbouring
// pairs at the same time, so x is read from HBM once and from L2 eleven times); W (4.7 MB) stays L2-resident.
// The epilogue is ~30 instructions per element, so SIXTEEN epilogue warps (4 per scheduler) are needed to keep it off
// the critical path: with eight, the kernel ran at the speed of GEMM + separate elementwise pass.
// Warp roles: 0-15 epilogue (warpgroup g = warp / 4 owns accumulator columns [64g, 64g + 64), thread <-> row = TMEM lane),
// 16 TMA producer, 17 MMA issuer (leader CTA only).
#include "gelu.cuh"
#include "kernels.cuh"
#include "philox.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * 128, B_BYTES = (BN / 2) * 128, STAGE_BYTES = A_BYTES + B_BYTES;   // per CTA: own rows, own half of W
// CTAs per cluster = the MMA pair.  The grid is sized from cudaOccupancyMaxActiveClusters (whole co-resident pairs).
constexpr int CL = 2;
constexpr int EPI_WARPS = 16;
constexpr int WSTAGE = 2 * 1024;            // per-warp staging ([32 rows][64 B]) for coalesced global stores
constexpr int THREADS = (EPI_WARPS + 2) * 32;

struct __align__(8) Ctrl {
  uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_free[2];
  uint64_t rbar[EPI_WARPS];                   // fp32-stream MODE 1: this warp's residual tile has landed (TMA load)
  uint32_t tmem_base;
};
// Shared-memory layout per instantiation.  The fp32-stream Linear + dropout + residual (MODE 1, RES32) moves its residual / output
// tiles by TMA as [32 rows][32 fp32] = 4 KB per epilogue warp (whole 128-byte lines; no registers, no LSU round trips), paid for
// with one ring stage (5 instead of 6) and a separate 128-byte bias row per warp (the tile is the TMA target).
// TILES = false keeps the register-staged fp32 epilogue and all six stages: the K = 4 D product (fc2) is tensor-bound, its epilogue is
// hidden, and the sixth stage is worth 2.5 % there (0.1910 vs 0.1958 ms).
template <int MODE, bool RES32, bool TILES>
struct Lay {
  static constexpr bool T32 = MODE == 1 && RES32 && TILES;
  static constexpr int NST = T32 ? 5 : STAGES;
  static constexpr int WS = T32 ? 4096 : WSTAGE;
  static constexpr int BIAS = T32 ? EPI_WARPS * 128 : 0;
  static constexpr size_t SMEM = (size_t)NST * STAGE_BYTES + EPI_WARPS * WS + BIAS + sizeof(Ctrl);
};

struct Params {
  int64_t M;
  int N, K;
  float p;
  uint64_t seed, offset;
  const uint64_t* offset_dev;
  const __nv_bfloat16* bias;
  __nv_bfloat16* u;                           // MODE 0: pre-activation output.  MODE 1: the RESIDUAL input (read only).
                                              // MODE 2: the saved pre-activation (read only); mask is read only too
  __nv_bfloat16* out;
  uint8_t* mask;
  float* partial;                             // MODE 2: (ceil(M/128) * 4, N) fp32 column partial sums of du
  int factor;                                 // MODE 0: store keep * gelu'(u) / (1 - p) in `u` instead of the pre-activation (u may
                                              // be NULL: inference, nothing saved).  MODE 2: `u` holds that factor (mask unused)
  uint32_t rk[2 * GVIT_PHILOX_ROUNDS];        // Philox round keys, precomputed on the host: constant-bank operands, no
};                                            // per-call key schedule in the epilogue

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// RES32 (MODE 1 only): the residual input and the output are fp32 - the residual stream torch.autocast keeps in fp32
// (vit.py:207-211,117-118: cat / add with the fp32 cls_token / pos_embed promote) - while the branch value is still
// rounded to bf16 first, exactly what `x + proj_drop(proj(o))` computes under autocast.
template <int MODE, bool RES32 = false, bool TILES = false>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) fc1_gelu_dropout_tc_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                         const __grid_constant__ CUtensorMap tm_w,
                                                                         const __grid_constant__ CUtensorMap tm_res,
                                                                         const __grid_constant__ CUtensorMap tm_o32,
                                                                         const Params P) {
  using L = Lay<MODE, RES32, TILES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  if ((smem_u32(ring) & 1023u) != 0) __trap();
  uint8_t* sStg = ring + (size_t)L::NST * STAGE_BYTES;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(sStg + EPI_WARPS * L::WS + L::BIAS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nM = (int)((P.M + BM - 1) / BM), nN = P.N / BN, nK = P.K / BK;
  const int rank = (int)cluster_ctarank();                 // which row block of the group, which slice of W to fetch
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int npairs = ((nM + CL - 1) / CL) * nN;            // (group of CL row blocks, column tile); row blocks past the end
                                                           // are empty (TMA zero-fills, the stores are row-guarded)

  if (warp == EPI_WARPS && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < L::NST; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int w8 = 0; w8 < EPI_WARPS; ++w8) mbar_init(&ctl->rbar[w8], 1);
    if (L::T32) { prefetch_tmap(&tm_res); prefetch_tmap(&tm_o32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->acc_full[s], 1); mbar_init(&ctl->acc_free[s], CL * EPI_WARPS); }   // leader's: one arrival per epilogue warp of the pair
    fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) tmem_alloc_2sm(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // the peer's barriers exist before anything is multicast to it
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == EPI_WARPS) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int pr = cid; pr < npairs; pr += ncl) {
        const int mp = pr / nN, n = pr - mp * nN, m = CL * mp + rank;
        for (int kb = 0; kb < nK; ++kb, ++it) {
          const int s = it % L::NST;
          mbar_wait(&ctl->empty[s], ((it / L::NST) & 1) ^ 1);            // released by the leader's multicast commit
          const uint32_t fullL = mapa_u32(smem_u32(&ctl->full[s]), 0);
          if (rank == 0) mbar_expect_tx(&ctl->full[s], (uint32_t)(CL * STAGE_BYTES));   // both CTAs' tiles
          tma_load_3d_2sm(ring + (size_t)s * STAGE_BYTES, &tm_x, kb * BK, m * BM, 0, fullL);                 // rows >= M: zeros
          if constexpr (MODE != 2) {                                     // W (N, K): K-major, this CTA's BN / 2 rows of the tile
            tma_load_3d_2sm(ring + (size_t)s * STAGE_BYTES + A_BYTES, &tm_w, kb * BK, n * BN + rank * (BN / CL), 0, fullL);
          } else {                                                       // W2 (K, N): MN-major, [64 k][64 n] atoms of 8 KB
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_3d_2sm(ring + (size_t)s * STAGE_BYTES + A_BYTES + j * 8192, &tm_w, n * BN + rank * (BN / CL) + j * 64, kb * BK, 0, fullL);
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(CL * BM, BN, false, MODE == 2);
      const uint32_t aR = smem_u32(ring);
      int it = 0, tc = 0;
      for (int pr = cid; pr < npairs; pr += ncl, ++tc) {
        const int buf = tc & 1;
        if (tc >= 2) {                              // both CTAs' epilogues have drained this accumulator buffer (use tc/2 - 1)
          mbar_wait(&ctl->acc_free[buf], ((tc >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nK; ++kb, ++it) {
          const int s = it % L::NST;
          mbar_wait(&ctl->full[s], (it / L::NST) & 1);
          tc_fence_after();
          const uint32_t aA = aR + s * STAGE_BYTES, aB = aA + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_2sm(tmem + buf * BN, make_sdesc(aA + kk * 32),
                        MODE == 2 ? make_sdesc_lbo(aB + kk * 2048, 8192) : make_sdesc(aB + kk * 32), idesc, kb > 0 || kk > 0);
          umma_commit_2sm_mc(&ctl->empty[s], (uint16_t)((1u << CL) - 1));  // slot s of BOTH CTAs may be refilled
        }
        umma_commit_2sm_mc(&ctl->acc_full[buf], (uint16_t)((1u << CL) - 1));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int g = warp >> 2;                                           // accumulator column quarter
    uint8_t* stg = sStg + warp * L::WS;
    uint32_t rph = 0;                                                  // T32: residual-tile arrivals consumed
    const uint32_t tl = tmem_lane_base(tmem, warp) + g * 64;
    const int ch4 = lane & 3, r4 = lane >> 2;                          // coalesced pattern: 4 lanes per 64-byte row segment
    const float scale = P.p > 0.f ? 1.0f / (1.0f - P.p) : 1.0f;
    const uint32_t th = dropout_thresh16(P.p);
    // Philox-4x32 with the host-made round keys
    auto philox = [&](uint32_t c0, uint32_t c1) -> uint4 {
      uint4 ctr = make_uint4(c0, c1, 0u, 0u);
#pragma unroll
      for (int r = 0; r < GVIT_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ P.rk[2 * r], lo1, hi0 ^ ctr.w ^ P.rk[2 * r + 1], lo0);
      }
      return ctr;
    };
    const uint64_t off = P.offset + (P.offset_dev ? __ldg(P.offset_dev) : 0ull);
    // staging tile [32 rows][64 B]: 16-byte chunk q of row r lives at chunk q ^ ((r >> 1) & 3) - conflict-free for the
    // row-per-lane writes and for the 4-lanes-per-row reads
    auto stage = [&](const uint32_t (&w)[16]) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    };
    // this warp's 32 staged rows x 32 columns -> global, eight 64-byte row segments per instruction
    auto flush = [&](__nv_bfloat16* dst, int64_t row0, int col0) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r4 + 8 * i;
        if (row0 + r < P.M) {
          const uint4 v4 = *reinterpret_cast<const uint4*>(stg + r * 64 + ((ch4 ^ ((r >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(dst + (row0 + r) * P.N + col0 + ch4 * 8) = v4;
        }
      }
      __syncwarp();
    };
    int tc = 0;
    for (int pr = cid; pr < npairs; pr += ncl, ++tc) {
      const int mp = pr / nN, n = pr - mp * nN, m = CL * mp + rank;
      const int buf = tc & 1;
      const int64_t wrow0 = (int64_t)m * BM + (warp & 3) * 32;         // first row of this warp
      const int64_t row = wrow0 + lane;
      // MODE 2: the saved tile (pre-activation or backward factor) of this warp's rows is the one global read of the
      // epilogue; it is fetched one 32-column half ahead - the first before the accumulator wait - so that its latency
      // lies under the MMAs / the previous half instead of in front of every half
      // MODE 1 on a bf16 stream: the same for the residual tile (0.0777 -> 0.0701 ms at B = 256, same box).  On the fp32 stream the
      // same prefetch (one 16-column quarter ahead) spilled and measured 2 % SLOWER (0.1085 -> 0.1105 ms): loaded in place there
      uint4 pre[4];
      auto prefetch_saved = [&](int h) {
        if constexpr (MODE == 2 || (MODE == 1 && !RES32)) {
          const int c0 = n * BN + g * 64 + h * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r4 + 8 * i;
            pre[i] = make_uint4(0, 0, 0, 0);
            if (wrow0 + r < P.M) pre[i] = *reinterpret_cast<const uint4*>(P.u + (wrow0 + r) * P.N + c0 + ch4 * 8);
          }
        }
      };
      prefetch_saved(0);
      // T32: the residual tile [32 rows][32 fp32] of the NEXT 32-column half is requested as soon as the staging tile's previous
      // store has been read - here before the accumulator wait, below behind the previous half's store
      auto request_resid = [&](int h) {
        if constexpr (L::T32) {
          if (lane == 0) {
            tma_store_wait_read();
            if (wrow0 < P.M) {
              mbar_expect_tx(&ctl->rbar[warp], 4096u);
              tma_load_3d(stg, &tm_res, n * BN + g * 64 + h * 32, (int)wrow0, 0, &ctl->rbar[warp]);
            }
          }
          __syncwarp();
        }
      };
      request_resid(0);
      mbar_wait(&ctl->acc_full[buf], (tc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {                                    // 32 columns at a time
        const int col0 = n * BN + g * 64 + h * 32;
        // bias of these 32 columns as fp32 in the (idle) staging area: every lane then reads it with broadcast LDS.128
        float* sbias = L::T32 ? reinterpret_cast<float*>(sStg + EPI_WARPS * L::WS + warp * 128) : reinterpret_cast<float*>(stg);
        if (MODE != 2 && lane < 8) {
          float4 bf4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(P.bias + col0 + lane * 4));
            bf4 = make_float4(bf_lo(raw.x), bf_hi(raw.x), bf_lo(raw.y), bf_hi(raw.y));
          }
          *reinterpret_cast<float4*>(sbias + lane * 4) = bf4;
        }
        float v[32];
        tmem_ld32(tl + buf * BN + h * 32, v);
        if (h == 1) {                                                  // this warp's accumulator rows have been read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&ctl->acc_free[buf]), 0));
        }
        __syncwarp();
        if constexpr (MODE == 2) {
          // ---- du = dh * keep / (1 - p) * gelu'(u); column sums of du over this warp's 32 rows
          uint32_t uu[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {                                // saved tile: coalesced global (prefetched) -> staging -> own row
            const int r = r4 + 8 * i;
            *reinterpret_cast<uint4*>(stg + r * 64 + ((ch4 ^ ((r >> 1) & 3)) << 4)) = pre[i];
          }
          if (h == 0) prefetch_saved(1);
          uint32_t keep = 0xffffffffu;
          if (!P.factor && P.p > 0.f && row < P.M) keep = *reinterpret_cast<const uint32_t*>(P.mask + ((row * P.N + col0) >> 3));
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
            uu[4 * q] = v4.x; uu[4 * q + 1] = v4.y; uu[4 * q + 2] = v4.z; uu[4 * q + 3] = v4.w;
          }
          __syncwarp();
          uint32_t dpk[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float gq[8], uq[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uq[2 * e] = bf_lo(uu[4 * q + e]); uq[2 * e + 1] = bf_hi(uu[4 * q + e]);
              gq[2 * e] = v[8 * q + 2 * e]; gq[2 * e + 1] = v[8 * q + 2 * e + 1];
            }
            if (P.factor) {                                            // the forward saved keep * gelu'(u) / (1 - p): one multiply
#pragma unroll
              for (int t = 0; t < 8; ++t) gq[t] *= uq[t];
            } else {
              Gelu<false>::grad8(gq, uq, scale);
#pragma unroll
              for (int t = 0; t < 8; ++t) gq[t] = (keep >> (8 * q + t)) & 1u ? gq[t] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) dpk[4 * q + e] = pack2(gq[2 * e], gq[2 * e + 1]);
          }
          stage(dpk);
          __syncwarp();
          // column sums of the STAGED (bf16, as stored) tile: lane = (row half, column pair); two halves joined by one shuffle
          {
            const int cp = lane & 15, rh = lane >> 4;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = rh * 16 + i;
              const uint32_t w2 = *reinterpret_cast<const uint32_t*>(stg + r * 64 + (((cp >> 2) ^ ((r >> 1) & 3)) << 4) + (cp & 3) * 4);
              s0 += bf_lo(w2);
              s1 += bf_hi(w2);
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            if (lane < 16)
              *reinterpret_cast<float2*>(P.partial + ((int64_t)m * 4 + (warp & 3)) * P.N + col0 + 2 * cp) = make_float2(s0, s1);
          }
          flush(P.out, wrow0, col0);                                   // du tile (flush syncs the warp before and after)
          continue;
        }
        uint32_t upk[16];                                              // u as stored: 32 bf16
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const ulonglong2 bq = *reinterpret_cast<const ulonglong2*>(sbias + q4 * 4);    // two packed fp32 pairs
          float x0, x1, x2, x3;
          up2(add2(pk2(v[4 * q4], v[4 * q4 + 1]), f32x2{bq.x}), x0, x1);
          up2(add2(pk2(v[4 * q4 + 2], v[4 * q4 + 3]), f32x2{bq.y}), x2, x3);
          upk[2 * q4] = pack2(x0, x1);
          upk[2 * q4 + 1] = pack2(x2, x3);
        }
        __syncwarp();                                                  // bias reads done: the staging area is free again
        uint32_t rres[16];                                             // MODE 1: this row's 32 residual values (bf16 pairs)
        if constexpr (MODE == 0) {
          if (!P.factor && P.u != nullptr) {
            stage(upk);
            flush(P.u, wrow0, col0);                                   // pre-activation tile
          }
        } else if constexpr (!RES32) {
          // residual tile: coalesced global -> staging (the store pattern in reverse), then every lane reads its row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r4 + 8 * i;
            *reinterpret_cast<uint4*>(stg + r * 64 + ((ch4 ^ ((r >> 1) & 3)) << 4)) = pre[i];
          }
          if (h == 0) prefetch_saved(1);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
            rres[4 * q] = v4.x; rres[4 * q + 1] = v4.y; rres[4 * q + 2] = v4.z; rres[4 * q + 3] = v4.w;
          }
          __syncwarp();
        }
        // Keep decisions for these 32 columns, BIT-SLICED: 16 Philox words w[15..0], bit j of w[i] = bit i of element j's
        // 16-bit uniform number r_j; keep_j = (r_j >= th) comes out of a serial comparator over the 16 bit planes (one
        // LOP3 per plane for all 32 elements at once) instead of 32 extract / compare / select / merge sequences, and the
        // result IS the keep-mask word of these columns.  (Same distribution, different bit assignment than
        // gvit_gelu_dropout_fwd: the two paths draw different - equally valid - masks from the same seed.)
        uint32_t keep = 0xffffffffu;
        if (P.p > 0.f) {
          const uint64_t ctr0 = off + (uint64_t)((row * P.N + col0) >> 3);           // 4 consecutive counters per 32 columns
          // borrow chain of r - th, least significant plane first: b' = t ? (~r | b) : (~r & b), which is ONE three-input
          // LOP3 per plane on (r, b, T) with T = the th bit spread to all 32 bits (uniform, hoisted): no borrow out = keep
          uint32_t bor = 0u;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const uint64_t ctr = ctr0 + c4;
            const uint4 r = philox((uint32_t)ctr, (uint32_t)(ctr >> 32));
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t T = 0u - ((th >> (4 * c4 + i)) & 1u);
              bor = (~w[i] & bor) | (T & (~w[i] | bor));
            }
          }
          keep = ~bor;                                                 // r >= th
        }
        if constexpr (L::T32) {
          // fp32 stream: the residual tile of these 32 columns has landed in the staging tile by TMA (128-byte swizzle: 16-byte
          // chunk q of row r at q ^ (r & 7)); every lane adds keep * scale * y to its row in place, one TMA store sends it out
          if (wrow0 < P.M) {
            mbar_wait(&ctl->rbar[warp], rph & 1);
            ++rph;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t o = lane * 128 + ((q ^ (lane & 7)) << 4);
              float4 r4v = *reinterpret_cast<const float4*>(stg + o);
              const uint32_t y01 = upk[2 * q], y23 = upk[2 * q + 1];
              const int bit = 4 * q;
              r4v.x += (keep >> bit) & 1u ? bf_lo(y01) * scale : 0.f;
              r4v.y += (keep >> (bit + 1)) & 1u ? bf_hi(y01) * scale : 0.f;
              r4v.z += (keep >> (bit + 2)) & 1u ? bf_lo(y23) * scale : 0.f;
              r4v.w += (keep >> (bit + 3)) & 1u ? bf_hi(y23) * scale : 0.f;
              *reinterpret_cast<float4*>(stg + o) = r4v;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tm_o32, stg, col0, (int)wrow0, 0);         // rows >= M are clipped by the TMA unit
              tma_store_commit();
            }
          }
          if (h == 0) request_resid(1);                                // waits for that store's read, then refills the tile
          if (P.p > 0.f && row < P.M) *reinterpret_cast<uint32_t*>(P.mask + ((row * P.N + col0) >> 3)) = keep;
          continue;
        }
        if constexpr (MODE == 1 && RES32 && !TILES) {
          // fp32 stream: 16 columns (64 bytes) at a time through the same staging tile - coalesced residual load, every lane
          // picks up its row, adds keep * scale * y, writes its row back, coalesced store
          const float* resid32 = reinterpret_cast<const float*>(P.u);
          float* out32 = reinterpret_cast<float*>(P.out);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c16 = col0 + hh * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r4 + 8 * i;
              uint4 v4 = make_uint4(0, 0, 0, 0);
              if (wrow0 + r < P.M) v4 = *reinterpret_cast<const uint4*>(resid32 + (wrow0 + r) * P.N + c16 + ch4 * 4);
              *reinterpret_cast<uint4*>(stg + r * 64 + ((ch4 ^ ((r >> 1) & 3)) << 4)) = v4;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 r4v = *reinterpret_cast<const float4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
              const uint32_t y01 = upk[8 * hh + 2 * q], y23 = upk[8 * hh + 2 * q + 1];
              const int bit = 16 * hh + 4 * q;
              r4v.x += (keep >> bit) & 1u ? bf_lo(y01) * scale : 0.f;
              r4v.y += (keep >> (bit + 1)) & 1u ? bf_hi(y01) * scale : 0.f;
              r4v.z += (keep >> (bit + 2)) & 1u ? bf_lo(y23) * scale : 0.f;
              r4v.w += (keep >> (bit + 3)) & 1u ? bf_hi(y23) * scale : 0.f;
              *reinterpret_cast<float4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = r4v;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r4 + 8 * i;
              if (wrow0 + r < P.M)
                *reinterpret_cast<uint4*>(out32 + (wrow0 + r) * P.N + c16 + ch4 * 4) =
                    *reinterpret_cast<const uint4*>(stg + r * 64 + ((ch4 ^ ((r >> 1) & 3)) << 4));
            }
            __syncwarp();
          }
          if (P.p > 0.f && row < P.M) *reinterpret_cast<uint32_t*>(P.mask + ((row * P.N + col0) >> 3)) = keep;
          continue;
        }
        uint32_t fpk[16];                                              // MODE 0, factor mode: keep * gelu'(u) / (1 - p), 32 bf16
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                  // 8 columns per step
          float a[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) { a[2 * e] = bf_lo(upk[4 * q + e]); a[2 * e + 1] = bf_hi(upk[4 * q + e]); }
          if constexpr (MODE == 0) {
            if (P.factor && P.u != nullptr) {                          // value and backward factor from one evaluation
              float dg[8];
              Gelu<false>::fwd_grad8(a, dg, scale);
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const bool kp = (keep >> (8 * q + t)) & 1u;
                a[t] = kp ? a[t] : 0.f;
                dg[t] = kp ? dg[t] : 0.f;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) fpk[4 * q + e] = pack2(dg[2 * e], dg[2 * e + 1]);
            } else {
              Gelu<false>::fwd8(a, scale);                             // GELU of the stored value, dropout scale folded in
#pragma unroll
              for (int t = 0; t < 8; ++t) a[t] = (keep >> (8 * q + t)) & 1u ? a[t] : 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {                              // resid + keep * scale * y (y as a bf16 GEMM would store it)
              const uint32_t rr = rres[4 * q + e];
              a[2 * e] = bf_lo(rr) + ((keep >> (8 * q + 2 * e)) & 1u ? a[2 * e] * scale : 0.f);
              a[2 * e + 1] = bf_hi(rr) + ((keep >> (8 * q + 2 * e + 1)) & 1u ? a[2 * e + 1] * scale : 0.f);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) upk[4 * q + e] = pack2(a[2 * e], a[2 * e + 1]);
        }
        stage(upk);
        flush(P.out, wrow0, col0);                                     // activation tile
        if constexpr (MODE == 0) {
          if (P.factor && P.u != nullptr) {
            stage(fpk);
            flush(P.u, wrow0, col0);                                   // backward factor tile (instead of the pre-activation)
          }
        }
        if (P.p > 0.f && P.mask != nullptr && row < P.M) *reinterpret_cast<uint32_t*>(P.mask + ((row * P.N + col0) >> 3)) = keep;
      }
    }
  }
  if (L::T32 && warp < EPI_WARPS && lane == 0) tma_store_wait_all();   // every output tile of this warp has landed
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // no CTA leaves while its peer may still multicast into it
  if (warp == EPI_WARPS + 1) tmem_dealloc_2sm(tmem, 512);
}

// out[c] = sum over the partial rows, fixed order (32 columns x 8 row lanes per CTA, then lanes 0..7): deterministic
__global__ void __launch_bounds__(256) partial_colsum_kernel(const float* __restrict__ partial, int nrows, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float sacc = 0.f;
  if (c < N)
    for (int i = ly; i < nrows; i += 8) sacc += partial[(int64_t)i * N + c];
  red[ly][cx] = sacc;
  __syncthreads();
  if (ly == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][cx];
    out[c] = t;
  }
}

}  // namespace

bool fc1_tc_supported(int64_t M, int N, int K) { return M >= 1 && N >= BN && N % BN == 0 && K >= BK && K % BK == 0; }

template <int MODE, bool RES32 = false, bool TILES = false>
static int fused_linear_launch(const void* x, const void* w, const void* bias, int64_t M, int N, int K, float p, uint64_t seed,
                               uint64_t offset, const uint64_t* offset_dev, void* u_or_resid, void* out, uint8_t* mask,
                               cudaStream_t st, float* partial = nullptr, int factor = 0) {
  using L = Lay<MODE, RES32, TILES>;
  CUtensorMap tm_x, tm_w, tm_res, tm_o32;
  int rc = make_tmap_bf16_3d(&tm_x, x, (uint64_t)K, (uint64_t)M, 1, (uint64_t)K, (uint64_t)M * K, BM);
  if (rc != GVIT_OK) return rc;
  if (MODE != 2)
    rc = make_tmap_bf16_3d(&tm_w, w, (uint64_t)K, (uint64_t)N, 1, (uint64_t)K, (uint64_t)N * K, BN / CL);   // one CTA's half of the tile
  else
    rc = make_tmap_bf16_3d(&tm_w, w, (uint64_t)N, (uint64_t)K, 1, (uint64_t)N, (uint64_t)N * K, 64);        // W2 (K, N): [64 k][64 n] atoms
  if (rc != GVIT_OK) return rc;
  if (L::T32) {                                                       // residual in / stream out as [32 rows][32 fp32] TMA tiles
    rc = make_tmap_f32_3d(&tm_res, u_or_resid, (uint64_t)N, (uint64_t)M, 1, (uint64_t)N, (uint64_t)M * N, 32);
    if (rc != GVIT_OK) return rc;
    rc = make_tmap_f32_3d(&tm_o32, out, (uint64_t)N, (uint64_t)M, 1, (uint64_t)N, (uint64_t)M * N, 32);
    if (rc != GVIT_OK) return rc;
  } else {
    tm_res = tm_x; tm_o32 = tm_x;                                      // never dereferenced
  }
  Params P{M, N, K, p, seed, offset, offset_dev, static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(u_or_resid),
           static_cast<__nv_bfloat16*>(out), mask, partial, factor, {}};
  for (int r = 0; r < GVIT_PHILOX_ROUNDS; ++r) {
    P.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
    P.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(fc1_gelu_dropout_tc_kernel<MODE, RES32, TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
  const int64_t ngroups = (((M + BM - 1) / BM + CL - 1) / CL) * (N / BN);
  int max_clusters = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((num_sms() / CL) * CL);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = L::SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, fc1_gelu_dropout_tc_kernel<MODE, RES32, TILES>, &cfg) != cudaSuccess || max_clusters < 1) {
      (void)cudaGetLastError();
      max_clusters = num_sms() / CL;
    }
  }
  const int64_t want = CL * ngroups, cap = (int64_t)CL * max_clusters;
  const int grid = (int)(want < cap ? want : cap);                  // whole, co-resident clusters: a persistent grid
  fc1_gelu_dropout_tc_kernel<MODE, RES32, TILES><<<grid, THREADS, L::SMEM, st>>>(tm_x, tm_w, tm_res, tm_o32, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int fc1_gelu_dropout_fwd_tc(const void* x, const void* w, const void* bias, int64_t M, int N, int K, float p, uint64_t seed,
                            uint64_t offset, const uint64_t* offset_dev, int save_mode, void* u, void* out, uint8_t* mask, cudaStream_t st) {
  return fused_linear_launch<0>(x, w, bias, M, N, K, p, seed, offset, offset_dev, u, out, mask, st, nullptr, save_mode);
}

int linear_dropout_residual_fwd_tc(const void* x, const void* w, const void* bias, const void* resid, int64_t M, int N, int K, float p,
                                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int resid_dtype, void* out, uint8_t* mask,
                                   cudaStream_t st) {
  if (resid_dtype == GVIT_F32) {
    // HBM-bound products (K up to 2 D): TMA tiles for the fp32 residual / output; tensor-bound ones keep the sixth ring stage
    if (K <= 1536) return fused_linear_launch<1, true, true>(x, w, bias, M, N, K, p, seed, offset, offset_dev, const_cast<void*>(resid), out, mask, st);
    return fused_linear_launch<1, true, false>(x, w, bias, M, N, K, p, seed, offset, offset_dev, const_cast<void*>(resid), out, mask, st);
  }
  return fused_linear_launch<1>(x, w, bias, M, N, K, p, seed, offset, offset_dev, const_cast<void*>(resid), out, mask, st);
}

// rows of the column-partial workspace of linear_gelu_dropout_bwd_tc: one per (row block incl. cluster padding, lane quarter)
int64_t fc2_bwd_partial_rows(int64_t M) { return (((M + BM - 1) / BM + CL - 1) / CL) * CL * 4; }

int linear_gelu_dropout_bwd_tc(const void* dout, const void* w2, const void* u, const uint8_t* mask, int64_t M, int N, int K, float p,
                               int saved_mode, void* du, float* colsum_out, float* partial_ws, cudaStream_t st) {
  int rc = fused_linear_launch<2>(dout, w2, nullptr, M, N, K, p, 0, 0, nullptr, const_cast<void*>(u), du, const_cast<uint8_t*>(mask), st,
                                  partial_ws, saved_mode);
  if (rc != GVIT_OK) return rc;
  partial_colsum_kernel<<<(N + 31) / 32, 256, 0, st>>>(partial_ws, (int)fc2_bwd_partial_rows(M), N, colsum_out);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

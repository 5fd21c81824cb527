// gemm2_tc.cu - the dense Linear GEMMs of the block as ONE persistent 2-SM tcgen05 kernel:
//
//   forward            y  = x W^T + b        (vit.py:59 qkv, :93 fc2, :28 patch projection)    A, B K-major,  bf16 out
//   input gradient     dx = dy W             (autograd of the same nn.Linear)                   A K-major, B MN-major, bf16 out
//   weight gradient    dW = dy^T x           (fp32, the master parameter's dtype)               A, B MN-major, fp32 out, split-K
//
// Tile = 256 x 256 per CTA PAIR (__cluster_dims__(2)): `tcgen05.mma.cta_group::2` with M = 256, N = 256, K = 16.  Each CTA
// stages its own 128 rows of A and its own 128 of the tile's 256 B columns per 64-deep K slab (16 + 16 KB per stage, six
// stages), i.e. per 2 x 4.2 MFLOP of tile work a CTA pulls 32 KB through L2 -> shared memory where an un-paired
// 128 x 256 tile pulls 48 KB, and every MMA reads 8 KB of shared-memory operands per CTA instead of 12 KB.  Accumulators:
// 128 lanes x 256 fp32 columns per CTA, double-buffered in TMEM (2 x 256 columns), so the epilogue of tile t runs under
// the MMAs of tile t + 1.
//
// Roles per CTA: warps 0-7 epilogue (warp & 3 = TMEM lane quadrant, warp >> 2 = column half), warp 8 TMA producer (both
// CTAs; completion bytes of both land on the LEADER's full barrier), warp 9 MMA issuer (leader only) and TMEM owner.
// Barriers: full[s] (leader, expect 64 KB), empty[s] (each CTA, released by a multicast tcgen05.commit), acc_full[b] (each
// CTA, multicast commit), acc_free[b] (leader, 16 arrivals: one per epilogue warp of the pair, the peer's through a
// shared::cluster address).
//
// Split-K (weight gradient: the reduction runs over all B * N tokens while dW has only 9..48 tiles, fewer than the 74 CTA
// pairs): K is cut into `splits` equal pieces and the work items are ordered split-major, so the pairs of one wave walk the
// SAME K range together - every operand slab comes from HBM once and from L2 for the other tiles (equalising the work
// stream-K style, each pair at its own K offset, was measured 25-55 % slower: the re-reads went to DRAM).  Each item
// leaves its fp32 partial tile in a workspace slot (plain coalesced stores, nothing waits) and splitk_reduce_kernel adds the
// pieces of every tile in split order into `out`: deterministic, and the partials are read back from L2.
#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int T_THREADS = 320;
constexpr int T_STAGES = 6;
constexpr int T_A_BYTES = 128 * 128, T_B_BYTES = 128 * 128, T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
constexpr int T_WSTG = 4096;                       // per-warp staging [32 rows][128 B]
constexpr int T_BM = 256, T_BN = 256;

struct __align__(8) TCtrl {
  uint64_t full[T_STAGES], empty[T_STAGES], acc_full[2], acc_free[2];
  uint32_t tmem_base;
};
constexpr size_t T_SMEM = (size_t)T_STAGES * T_STAGE_BYTES + 8 * T_WSTG + sizeof(TCtrl);

struct TParams {
  int64_t M;                                       // rows of out
  int N, K;                                        // columns of out, reduction length
  int a_t, b_t;                                    // 0: K-major ([rows][K]); 1: MN-major ([K][rows])
  int splits, slabs_per_split;                     // splits > 1: item = (piece, tile), partial tiles go to `ws`
  float* ws;                                       // splits * tiles slots of 256 x 256 fp32
  const __nv_bfloat16* bias;                       // N values added to every row (nullable)
  void* out;
  int out_f32;
  int64_t out_rs;                                  // row stride of out, elements
};

struct TItem { int tile, k0, k1, slot; };           // slot < 0: the tile goes to `out`, else to workspace slot `slot`

// i-th work item of CTA pair `cid` (of `ncl`): items are (piece, tile) pairs, piece-major, dealt round robin
__device__ __forceinline__ bool get_item(const TParams& P, int tiles, int nslabs, int cid, int ncl, int i, TItem& it) {
  const int t = cid + i * ncl;
  if (t >= tiles * P.splits) return false;
  const int sp = t / tiles;
  it.tile = t - sp * tiles;
  it.k0 = sp * P.slabs_per_split;
  it.k1 = min(nslabs, it.k0 + P.slabs_per_split);
  it.slot = P.splits > 1 ? t : -1;
  return true;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T_THREADS, 1) gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                          const __grid_constant__ CUtensorMap tmB,
                                                                                          const TParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  if ((smem_u32(ring) & 1023u) != 0) __trap();
  uint8_t* sStg = ring + (size_t)T_STAGES * T_STAGE_BYTES;
  TCtrl* ctl = reinterpret_cast<TCtrl*>(sStg + 8 * T_WSTG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int mtiles = (int)((P.M + T_BM - 1) / T_BM), ntiles = P.N / T_BN;
  const int tiles = mtiles * ntiles;
  const int nslabs = (P.K + 63) / 64;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->acc_full[s], 1); mbar_init(&ctl->acc_free[s], 16); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2sm(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // the peer's barriers and TMEM exist before anything reaches them
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t c = 0;
      TItem it;
      for (int i = 0; get_item(P, tiles, nslabs, cid, ncl, i, it); ++i) {
        const int mt = it.tile / ntiles, nt = it.tile - mt * ntiles;
        const int m0 = mt * T_BM + rank * 128, n0 = nt * T_BN + rank * 128;
        for (int ks = it.k0; ks < it.k1; ++ks, ++c) {
          const uint32_t s = c % T_STAGES;
          mbar_wait(&ctl->empty[s], ((c / T_STAGES) & 1) ^ 1);
          const uint32_t fullL = mapa_u32(smem_u32(&ctl->full[s]), 0);
          if (rank == 0) mbar_expect_tx(&ctl->full[s], 2u * T_STAGE_BYTES);     // this CTA's 32 KB and the peer's
          uint8_t* dA = ring + (size_t)s * T_STAGE_BYTES;
          uint8_t* dB = dA + T_A_BYTES;
          if (P.a_t == 0) {
            tma_load_3d_2sm(dA, &tmA, ks * 64, m0, 0, fullL);                    // [128 rows][64 k]
          } else {                                                               // stored [K][M]: two [64 k][64 m] boxes
            tma_load_3d_2sm(dA, &tmA, m0, ks * 64, 0, fullL);
            tma_load_3d_2sm(dA + 8192, &tmA, m0 + 64, ks * 64, 0, fullL);
          }
          if (P.b_t == 0) {
            tma_load_3d_2sm(dB, &tmB, ks * 64, n0, 0, fullL);                    // [128 n][64 k]
          } else {                                                               // stored [K][N]: two [64 k][64 n] boxes
            tma_load_3d_2sm(dB, &tmB, n0, ks * 64, 0, fullL);
            tma_load_3d_2sm(dB + 8192, &tmB, n0 + 64, ks * 64, 0, fullL);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(T_BM, T_BN, P.a_t != 0, P.b_t != 0);
      uint32_t c = 0, tc_ = 0;
      TItem it;
      for (int i = 0; get_item(P, tiles, nslabs, cid, ncl, i, it); ++i, ++tc_) {
        const uint32_t buf = tc_ & 1;
        mbar_wait(&ctl->acc_free[buf], ((tc_ >> 1) & 1) ^ 1);                    // both CTAs' epilogues drained tile t-2
        tc_fence_after();
        bool acc = false;
        for (int ks = it.k0; ks < it.k1; ++ks, ++c) {
          const uint32_t s = c % T_STAGES;
          mbar_wait(&ctl->full[s], (c / T_STAGES) & 1);
          tc_fence_after();
          const uint32_t aA = smem_u32(ring + (size_t)s * T_STAGE_BYTES), aB = aA + T_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = P.a_t ? make_sdesc_lbo(aA + kk * 2048, 8192) : make_sdesc(aA + kk * 32);
            const uint64_t db = P.b_t ? make_sdesc_lbo(aB + kk * 2048, 8192) : make_sdesc(aB + kk * 32);
            umma_ss_2sm(tmem + buf * 256, da, db, idesc, acc);
            acc = true;
          }
          umma_commit_2sm_mc(&ctl->empty[s], 3);                                 // slot s of BOTH CTAs may be refilled
        }
        umma_commit_2sm_mc(&ctl->acc_full[buf], 3);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3, hc = warp >> 2;
    uint8_t* stg = sStg + warp * T_WSTG;
    const int ch8 = lane & 7, r8 = lane >> 3;
    uint32_t tc_ = 0;
    TItem it;
    for (int i = 0; get_item(P, tiles, nslabs, cid, ncl, i, it); ++i, ++tc_) {
      const int mt = it.tile / ntiles, nt = it.tile - mt * ntiles;
      const int64_t wrow0 = (int64_t)mt * T_BM + rank * 128 + q * 32;            // first row of this warp
      const int n0 = nt * T_BN + hc * 128;                                       // first column of this warp
      const uint32_t buf = tc_ & 1;
      mbar_wait(&ctl->acc_full[buf], (tc_ >> 1) & 1);
      tc_fence_after();
      const uint32_t tA = tmem_lane_base(tmem, warp) + buf * 256 + hc * 128;
      const uint32_t freeL = mapa_u32(smem_u32(&ctl->acc_free[buf]), 0);
      if (!P.out_f32) {
        __nv_bfloat16* out = static_cast<__nv_bfloat16*>(P.out);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 64) {
          float v0[32], v1[32];
          tmem_ld32(tA + c0, v0);
          tmem_ld32(tA + c0 + 32, v1);
          if (c0 == 64) {                                                        // last read of this accumulator by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(freeL);
          }
          if (P.bias) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(P.bias + n0 + c0 + 8 * g));
              const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(P.bias + n0 + c0 + 32 + 8 * g));
              const uint32_t w0[4] = {b0.x, b0.y, b0.z, b0.w}, w1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v0[8 * g + 2 * e] += bf_lo(w0[e]); v0[8 * g + 2 * e + 1] += bf_hi(w0[e]);
                v1[8 * g + 2 * e] += bf_lo(w1[e]); v1[8 * g + 2 * e + 1] += bf_hi(w1[e]);
              }
            }
          }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((qq ^ (lane & 7)) << 4)) =
                make_uint4(pack2(v0[8 * qq], v0[8 * qq + 1]), pack2(v0[8 * qq + 2], v0[8 * qq + 3]),
                           pack2(v0[8 * qq + 4], v0[8 * qq + 5]), pack2(v0[8 * qq + 6], v0[8 * qq + 7]));
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + qq) ^ (lane & 7)) << 4)) =
                make_uint4(pack2(v1[8 * qq], v1[8 * qq + 1]), pack2(v1[8 * qq + 2], v1[8 * qq + 3]),
                           pack2(v1[8 * qq + 4], v1[8 * qq + 5]), pack2(v1[8 * qq + 6], v1[8 * qq + 7]));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r8 + 4 * i;
            if (wrow0 + r < P.M)
              *reinterpret_cast<uint4*>(out + (wrow0 + r) * P.out_rs + n0 + c0 + ch8 * 8) =
                  *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
          }
          __syncwarp();
        }
      } else {
        // fp32: straight into `out`, or (stream-K) this segment's partial tile into its workspace slot, row-major 256 x 256
        const bool part = it.slot >= 0;
        float* base = part ? P.ws + (size_t)it.slot * (T_BM * T_BN) + (size_t)(rank * 128 + q * 32) * T_BN + hc * 128
                           : static_cast<float*>(P.out) + wrow0 * P.out_rs + n0;
        const int64_t rs = part ? T_BN : P.out_rs;
        const int64_t rows_ok = part ? 32 : P.M - wrow0;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float v0[32];
          tmem_ld32(tA + c0, v0);
          if (c0 == 96) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(freeL);
          }
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((qq ^ (lane & 7)) << 4)) =
                make_float4(v0[4 * qq], v0[4 * qq + 1], v0[4 * qq + 2], v0[4 * qq + 3]);
          __syncwarp();
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = r8 + 4 * i8;
            if (r < rows_ok)
              *reinterpret_cast<float4*>(base + r * rs + c0 + ch8 * 4) =
                  *reinterpret_cast<const float4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // no CTA leaves (or frees TMEM) while the pair may still touch it
  if (warp == 9) tmem_dealloc_2sm(tmem, 512);
}

// out tile t = sum over the pieces s = 0..splits-1 of workspace slot s * tiles + t, in that order.  One block = 16 rows of
// one tile; thread = one float4 of a row (64 per row, 4 rows per pass).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int64_t out_rs,
                                                            int64_t M, int tiles, int ntiles, int splits) {
  const int tile = blockIdx.x >> 4, rblk = blockIdx.x & 15;
  const int mt = tile / ntiles, nt = tile - mt * ntiles;
  const int c4 = threadIdx.x & 63, rr = threadIdx.x >> 6;
  float4 acc[4];
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) acc[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < splits; ++sp) {
    const float* src = ws + ((size_t)sp * tiles + tile) * (T_BM * T_BN) + c4 * 4;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(rblk * 16 + pass * 4 + rr) * T_BN));
      acc[pass].x += v.x; acc[pass].y += v.y; acc[pass].z += v.z; acc[pass].w += v.w;
    }
  }
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int64_t row = (int64_t)mt * T_BM + rblk * 16 + pass * 4 + rr;
    if (row < M) *reinterpret_cast<float4*>(out + row * out_rs + nt * T_BN + c4 * 4) = acc[pass];
  }
}

// operand -> tensor map.  K-major: stored [rows][K], box {64 k, 128 rows}; MN-major: stored [K][rows], box {64 rows, 64 k}
int operand_map2(CUtensorMap* tm, const void* ptr, int t, int64_t rows, int K, int64_t rs) {
  if (t == 0) {
    const int64_t inner = (K + 63) & ~63;
    GVIT_REQUIRE(rs >= inner, GVIT_ERR_SHAPE, "gemm: a K-major operand needs K=%d padded to 64 inside its row stride %lld", K, (long long)rs);
    return make_tmap_bf16_3d(tm, ptr, (uint64_t)inner, (uint64_t)rows, 1, (uint64_t)rs, (uint64_t)rs * rows, 128);
  }
  const int64_t inner = (rows + 63) & ~63;
  GVIT_REQUIRE(rs >= inner, GVIT_ERR_SHAPE, "gemm: an MN-major operand needs its %lld rows padded to 64 inside its row stride %lld",
               (long long)rows, (long long)rs);
  return make_tmap_bf16_3d(tm, ptr, (uint64_t)inner, (uint64_t)K, 1, (uint64_t)rs, (uint64_t)rs * K, 64);
}

int max_pairs() {
  static int cached = 0;
  if (cached > 0) return cached;
  int n = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((num_sms() / 2) * 2);
  cfg.blockDim = dim3(T_THREADS);
  cfg.dynamicSmemBytes = T_SMEM;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&n, gemm2_tc_kernel, &cfg) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    n = num_sms() / 2;
  }
  cached = n;
  return n;
}

}  // namespace

bool gemm2_tc_supported(int64_t M, int N, int K) { return M >= 1 && N >= T_BN && N % T_BN == 0 && K >= 1; }

// How many pieces the reduction of an fp32-output product is cut into.  Cost model in units of one K slab of one pair
// (512 tensor cycles, ~0.3 us): the main loops take waves * slabs_per_piece; every partial tile costs ~0.3 slabs of reduce
// time (256 KB written and read back at L2 speed, spread over the chip) and splitting at all costs ~16 (an exposed epilogue
// and the second launch).  ViT-B/16, 50432 rows: qkv 27 tiles -> 5 pieces, proj 9 -> 8, fc1 / fc2 36 -> 2.
int gemm2_splits(int64_t M, int N, int K) {
  const int tiles = (int)((M + T_BM - 1) / T_BM) * (N / T_BN);
  const int nslabs = (K + 63) / 64, pairs = num_sms() / 2;
  if (tiles >= pairs) return 1;
  int best = 1;
  double best_cost = (double)((tiles + pairs - 1) / pairs) * nslabs;
  for (int s = 2; s <= 16 && s * 16 <= nslabs; ++s) {
    const int items = tiles * s, waves = (items + pairs - 1) / pairs, per = (nslabs + s - 1) / s;
    const double cost = (double)waves * per + 0.3 * items + 16.0;
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// bytes of partial-tile workspace an fp32 product wants (0: none)
int64_t gemm2_ws_bytes(int64_t M, int N, int K) {
  const int s = gemm2_splits(M, N, K);
  if (s <= 1) return 0;
  return (int64_t)s * ((M + T_BM - 1) / T_BM) * (N / T_BN) * T_BM * T_BN * sizeof(float);
}

int gemm2_tc(const void* a, int a_t, int64_t a_rs, const void* b, int b_t, int64_t b_rs, int64_t M, int N, int K, const void* bias,
             int out_dtype, void* out, int64_t out_rs, float* ws, int64_t ws_bytes, cudaStream_t st) {
  GVIT_REQUIRE(gemm2_tc_supported(M, N, K), GVIT_ERR_SHAPE, "gemm: M=%lld N=%d K=%d (N must be a multiple of 256)", (long long)M, N, K);
  TParams P;
  P.M = M; P.N = N; P.K = K; P.a_t = a_t; P.b_t = b_t;
  P.bias = static_cast<const __nv_bfloat16*>(bias);
  P.out = out; P.out_f32 = out_dtype == GVIT_F32; P.out_rs = out_rs;
  P.ws = ws;
  const int ntiles = N / T_BN;
  const int tiles = (int)((M + T_BM - 1) / T_BM) * ntiles;
  const int nslabs = (K + 63) / 64;
  GVIT_REQUIRE(!(bias != nullptr && P.out_f32), GVIT_ERR_SHAPE, "gemm: bias is added on the bf16 output path only");
  P.splits = 1;
  if (P.out_f32 && ws != nullptr && ws_bytes >= gemm2_ws_bytes(M, N, K)) P.splits = gemm2_splits(M, N, K);
  P.slabs_per_split = (nslabs + P.splits - 1) / P.splits;
  P.splits = (nslabs + P.slabs_per_split - 1) / P.slabs_per_split;          // no empty piece
  CUtensorMap tmA, tmB;
  int rc = operand_map2(&tmA, a, a_t, M, K, a_rs);
  if (rc != GVIT_OK) return rc;
  rc = operand_map2(&tmB, b, b_t, N, K, b_rs);
  if (rc != GVIT_OK) return rc;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(gemm2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM));
  const int64_t items = (int64_t)tiles * P.splits;
  const int pairs = max_pairs();
  const int ncl = (int)(items < pairs ? items : pairs);
  gemm2_tc_kernel<<<2 * ncl, T_THREADS, T_SMEM, st>>>(tmA, tmB, P);
  GVIT_CHECK_LAUNCH();
  if (P.splits > 1) {
    splitk_reduce_kernel<<<tiles * 16, 256, 0, st>>>(ws, static_cast<float*>(out), out_rs, M, tiles, ntiles, P.splits);
    GVIT_CHECK_LAUNCH();
  }
  return GVIT_OK;
}

}  // namespace gvit

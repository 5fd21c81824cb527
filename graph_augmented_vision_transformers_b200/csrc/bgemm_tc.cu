// bgemm_tc.cu - batched bf16 GEMM on tcgen05 for the DENSE-adjacency graph layer (mode 'dense';
// BASELINE configs[3]: ViT-L/16 at 384x384, Np = 576, D = 1024).
//
//   out[b] = diag(row_scale[b]) * sum_p  op(A_p[b]) op(B_p[b])          b = 0..batch-1, up to two products per launch
//
// With the adjacency dense, every stage of the layer is a per-image matrix product over operands that already sit in
// HBM in the layouts the layer uses - tokens (Np, D) inside the (B, 1+Np, D) stream, the (Np, Np) adjacency - and the
// four transposition cases all occur:
//     G  = P P^T            A K-major  (rows x K),  B K-major  (N rows x K)        similarity Gram matrix (fp32 out)
//     Z  = A~ P             A K-major,              B MN-major (K rows x N)        aggregation  (row_scale = 1)
//     dA~ = dZ P^T          A K-major,              B K-major                      (fp32 out)
//     T  = A~^T dZ          A MN-major (K rows x M), B MN-major
//     V  = dG P + dG^T P    two products into ONE accumulator (A K-major, then MN-major; B MN-major)
// so this is one persistent warp-specialised kernel with the operand majors as runtime flags: a [rows][64] 128B-swizzled
// shared-memory tile is read K-major or MN-major by descriptor only (tc.cuh), never transposed in memory.
// Tiles: 128 x BN (BN in {64,128,192,256}, the largest that divides N), K in slabs of 64 through a 4-stage TMA ring;
// fp32 accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
// Warp roles: 0-7 epilogue (warp & 3 = TMEM lane quadrant, warp >> 2 = column half), 8 TMA producer, 9 MMA issuer.
#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int G_THREADS = 320;
constexpr int G_STAGES = 4;
constexpr int G_A_BYTES = 128 * 128;               // [128][64] K-major, or two [64][64] MN-major sub-tiles
constexpr int G_B_BYTES = 256 * 128;               // up to [256][64] K-major, or four [64][64] MN-major sub-tiles
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_WSTG = 4096;                       // per-warp staging [32 rows][128 B] for coalesced stores

struct __align__(8) GCtrl {
  uint64_t full[G_STAGES], empty[G_STAGES], acc_full[2], acc_free[2];
  uint32_t tmem_base;
};
constexpr size_t G_SMEM = (size_t)G_STAGES * G_STAGE_BYTES + 8 * G_WSTG + sizeof(GCtrl);

struct GParams {
  int batch, M, N, BN, nprod;
  int K[2], a_t[2], b_t[2];
  const float* row_scale;                          // batch * M (nullable)
  void* out;
  int out_f32;
  int64_t out_rs, out_bs;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(G_THREADS, 1) bgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                const __grid_constant__ CUtensorMap tmB0,
                                                                const __grid_constant__ CUtensorMap tmA1,
                                                                const __grid_constant__ CUtensorMap tmB1, const GParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  if ((smem_u32(ring) & 1023u) != 0) __trap();
  uint8_t* sStg = ring + (size_t)G_STAGES * G_STAGE_BYTES;
  GCtrl* ctl = reinterpret_cast<GCtrl*>(sStg + 8 * G_WSTG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int BN = P.BN;
  const int mtiles = (P.M + 127) / 128, ntiles = (P.N + BN - 1) / BN;
  const int tiles = P.batch * mtiles * ntiles;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB0);
    if (P.nprod > 1) { prefetch_tmap(&tmA1); prefetch_tmap(&tmB1); }
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->acc_full[s], 1); mbar_init(&ctl->acc_free[s], 256); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t c = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int nt = t % ntiles, mt = (t / ntiles) % mtiles, b = t / (ntiles * mtiles);
        const int m0 = mt * 128, n0 = nt * BN;
        for (int p = 0; p < P.nprod; ++p) {
          const CUtensorMap* ta = p == 0 ? &tmA0 : &tmA1;
          const CUtensorMap* tb = p == 0 ? &tmB0 : &tmB1;
          const int ksl = (P.K[p] + 63) / 64;
          for (int ks = 0; ks < ksl; ++ks, ++c) {
            const uint32_t s = c % G_STAGES;
            mbar_wait(&ctl->empty[s], ((c / G_STAGES) & 1) ^ 1);
            mbar_expect_tx(&ctl->full[s], (uint32_t)(G_A_BYTES + BN * 128));
            uint8_t* dA = ring + (size_t)s * G_STAGE_BYTES;
            uint8_t* dB = dA + G_A_BYTES;
            if (P.a_t[p] == 0) {
              tma_load_3d(dA, ta, ks * 64, m0, b, &ctl->full[s]);                       // [128 rows][64 k]
            } else {                                                                     // stored [K][M]: two [64 k][64 m] boxes
              tma_load_3d(dA, ta, m0, ks * 64, b, &ctl->full[s]);
              tma_load_3d(dA + 8192, ta, m0 + 64, ks * 64, b, &ctl->full[s]);
            }
            if (P.b_t[p] == 0) {
              tma_load_3d(dB, tb, ks * 64, n0, b, &ctl->full[s]);                       // [BN rows][64 k]
            } else {                                                                     // stored [K][N]: BN/64 [64 k][64 n] boxes
              for (int j = 0; j < BN / 64; ++j) tma_load_3d(dB + j * 8192, tb, n0 + j * 64, ks * 64, b, &ctl->full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      uint32_t c = 0, tc_ = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tc_) {
        const uint32_t buf = tc_ & 1;
        mbar_wait(&ctl->acc_free[buf], ((tc_ >> 1) & 1) ^ 1);                            // drained by the epilogue of tile t-2
        tc_fence_after();
        bool acc = false;
        for (int p = 0; p < P.nprod; ++p) {
          const uint32_t idesc = make_idesc(128, BN, P.a_t[p] != 0, P.b_t[p] != 0);
          const int ksl = (P.K[p] + 63) / 64;
          for (int ks = 0; ks < ksl; ++ks, ++c) {
            const uint32_t s = c % G_STAGES;
            mbar_wait(&ctl->full[s], (c / G_STAGES) & 1);
            tc_fence_after();
            const uint32_t aA = smem_u32(ring + (size_t)s * G_STAGE_BYTES), aB = aA + G_A_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = P.a_t[p] ? make_sdesc_lbo(aA + kk * 2048, 8192) : make_sdesc(aA + kk * 32);
              const uint64_t db = P.b_t[p] ? make_sdesc_lbo(aB + kk * 2048, 8192) : make_sdesc(aB + kk * 32);
              umma_ss(tmem + buf * 256, da, db, idesc, acc);
              acc = true;
            }
            umma_commit(&ctl->empty[s]);
          }
        }
        umma_commit(&ctl->acc_full[buf]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int q = warp & 3, hc = warp >> 2;
    const int row = q * 32 + lane;
    uint8_t* stg = sStg + warp * G_WSTG;
    const int ch8 = lane & 7, r8 = lane >> 3;
    const int half = BN / 2;                                           // columns of this warp: [hc * half, hc * half + half)
    uint32_t tc_ = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tc_) {
      const int nt = t % ntiles, mt = (t / ntiles) % mtiles, b = t / (ntiles * mtiles);
      const int m0 = mt * 128, n0 = nt * BN;
      const uint32_t buf = tc_ & 1;
      const int m = m0 + row;
      const float rs = (P.row_scale && m < P.M) ? P.row_scale[(int64_t)b * P.M + m] : 1.0f;
      mbar_wait(&ctl->acc_full[buf], (tc_ >> 1) & 1);
      tc_fence_after();
      const uint32_t tA = tmem_lane_base(tmem, warp) + buf * 256 + hc * half;
      const int wrow0 = m0 + q * 32;
      // one pass = 128 bytes per row: 64 bf16 or 32 fp32 columns, staged per warp, stored as whole row segments
      const int cpp = P.out_f32 ? 32 : 64;
      for (int c0 = 0; c0 < half; c0 += cpp) {
        const int ncol = min(cpp, half - c0);                          // 32 or 64 (bf16 tail: 32)
        float v0[32], v1[32];
        tmem_ld32(tA + c0, v0);
        if (ncol > 32) tmem_ld32(tA + c0 + 32, v1);
        if (c0 + cpp >= half) {                                        // last read of this accumulator by this thread
          tc_fence_before();
          mbar_arrive(&ctl->acc_free[buf]);
        }
        if (P.out_f32) {
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((qq ^ (lane & 7)) << 4)) =
                make_float4(v0[4 * qq] * rs, v0[4 * qq + 1] * rs, v0[4 * qq + 2] * rs, v0[4 * qq + 3] * rs);
        } else {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((qq ^ (lane & 7)) << 4)) =
                make_uint4(pack2(v0[8 * qq] * rs, v0[8 * qq + 1] * rs), pack2(v0[8 * qq + 2] * rs, v0[8 * qq + 3] * rs),
                           pack2(v0[8 * qq + 4] * rs, v0[8 * qq + 5] * rs), pack2(v0[8 * qq + 6] * rs, v0[8 * qq + 7] * rs));
          if (ncol > 32) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq)
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + qq) ^ (lane & 7)) << 4)) =
                  make_uint4(pack2(v1[8 * qq] * rs, v1[8 * qq + 1] * rs), pack2(v1[8 * qq + 2] * rs, v1[8 * qq + 3] * rs),
                             pack2(v1[8 * qq + 4] * rs, v1[8 * qq + 5] * rs), pack2(v1[8 * qq + 6] * rs, v1[8 * qq + 7] * rs));
          }
        }
        __syncwarp();
        const int esz = P.out_f32 ? 4 : 2;
        const int nchunk = ncol * esz / 16;                            // valid 16-byte chunks per row in this pass
        const int col = n0 + hc * half + c0 + ch8 * (16 / esz);        // first column of this lane's chunk
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r8 + 4 * i;
          if (wrow0 + r < P.M && ch8 < nchunk && col < P.N) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch8 ^ (r & 7)) << 4));
            uint8_t* dst = static_cast<uint8_t*>(P.out) + ((int64_t)b * P.out_bs + (int64_t)(wrow0 + r) * P.out_rs + col) * esz;
            *reinterpret_cast<uint4*>(dst) = v4;
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// operand -> tensor map.  K-major (t == 0): stored [rows][K], box {64 k, box_rows}; MN-major: stored [K][rows], box {64, 64}
int operand_map(CUtensorMap* tm, const void* ptr, int t, int rows, int K, int batch, int64_t rs, int64_t bs, int box_rows) {
  if (t == 0) {
    const int inner = (K + 63) & ~63;
    GVIT_REQUIRE(rs >= inner, GVIT_ERR_SHAPE, "bgemm: a K-major operand needs its K extent (%d) padded to 64 inside the row stride (%lld)",
                 K, (long long)rs);
    return make_tmap_bf16_3d(tm, ptr, inner, rows, batch, rs, bs, box_rows);
  }
  const int inner = (rows + 63) & ~63;
  GVIT_REQUIRE(rs >= inner, GVIT_ERR_SHAPE, "bgemm: an MN-major operand needs its row extent (%d) padded to 64 inside the row stride (%lld)",
               rows, (long long)rs);
  return make_tmap_bf16_3d(tm, ptr, inner, K, batch, rs, bs, 64);
}

}  // namespace

int bgemm_tc(int batch, int M, int N, int nprod, const BgemmProduct* prods, const float* row_scale, int out_dtype, void* out,
             int64_t out_rs, int64_t out_bs, cudaStream_t st) {
  GVIT_REQUIRE(batch >= 1 && M >= 1 && N >= 1 && nprod >= 1 && nprod <= 2, GVIT_ERR_SHAPE,
               "bgemm: batch=%d M=%d N=%d nprod=%d", batch, M, N, nprod);
  {                                                     // rows are written in whole 16-byte chunks: the row stride must hold the last one
    const int per = out_dtype == GVIT_F32 ? 4 : 8;
    GVIT_REQUIRE(out_rs >= (int64_t)((N + per - 1) / per * per), GVIT_ERR_SHAPE,
                 "bgemm: out row stride %lld is shorter than N=%d rounded up to a 16-byte chunk", (long long)out_rs, N);
  }
  GParams P;
  P.batch = batch; P.M = M; P.N = N; P.nprod = nprod;
  const int Npad = (N + 63) & ~63;
  P.BN = Npad % 256 == 0 ? 256 : Npad % 192 == 0 ? 192 : Npad % 128 == 0 ? 128 : 64;
  if (Npad > 256 && P.BN == 64) P.BN = 256;            // odd multiples of 64: full-width tiles, the last one clipped
  P.row_scale = row_scale;
  P.out = out; P.out_f32 = out_dtype == GVIT_F32;
  P.out_rs = out_rs; P.out_bs = out_bs;
  CUtensorMap tm[4];
  for (int p = 0; p < 2; ++p) {
    const BgemmProduct& pr = prods[p < nprod ? p : 0];
    P.K[p] = pr.K; P.a_t[p] = pr.a_t; P.b_t[p] = pr.b_t;
    GVIT_REQUIRE(pr.K >= 1 && pr.a && pr.b, GVIT_ERR_SHAPE, "bgemm: product %d: K=%d or null operand", p, pr.K);
    int rc = operand_map(&tm[2 * p], pr.a, pr.a_t, M, pr.K, batch, pr.a_rs, pr.a_bs, 128);
    if (rc != GVIT_OK) return rc;
    rc = operand_map(&tm[2 * p + 1], pr.b, pr.b_t, N, pr.K, batch, pr.b_rs, pr.b_bs, P.BN);
    if (rc != GVIT_OK) return rc;
  }
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(bgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM));
  const int64_t tiles = (int64_t)batch * ((M + 127) / 128) * ((N + P.BN - 1) / P.BN);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  bgemm_tc_kernel<<<grid, G_THREADS, G_SMEM, st>>>(tm[0], tm[1], tm[2], tm[3], P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit

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
#pragma unroll
      for (int s = 0; s < KT; ++s) key[s] = 0u;             // below every real key
      const float ci = rn_i * KEY_SCALE;                    // fixed-point scale folded into the row factor
      // one 32-column block of this thread's similarity row; MASK only for the block that straddles Np
      auto block = [&](int c0, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        float v[32];
        tmem_ld32(trow + c0, v);
        const uint32_t cbase = 255u - (uint32_t)c0;
#pragma unroll
        for (int t8 = 0; t8 < 32; t8 += 8) {
          if (MASK && c0 + t8 >= Np) break;                 // whole group beyond the last token (warp-uniform)
          const float4 rna = *reinterpret_cast<const float4*>(&ctl->rn[c0 + t8]);
          const float4 rnb = *reinterpret_cast<const float4*>(&ctl->rn[c0 + t8 + 4]);
          const float rnv[8] = {rna.x, rna.y, rna.z, rna.w, rnb.x, rnb.y, rnb.z, rnb.w};
          uint32_t cand[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            // round(s * KEY_SCALE) lands in the low mantissa bits of (x + 1.5 * 2^23); the << 8 drops the exponent byte
            // and leaves 2^22 + round(s * KEY_SCALE) in bits 8..31
            const float x = fmaf(v[t8 + u] * ci, rnv[u], 12582912.0f);
            cand[u] = (__float_as_uint(x) << 8) + (cbase - (uint32_t)(t8 + u));
            if (MASK && c0 + t8 + u >= Np) cand[u] = 0u;    // columns >= Np can never be selected
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
      int c0 = 0;
      for (; c0 + 32 <= Np; c0 += 32) block(c0, std::false_type{});
      if (c0 < Np) block(c0, std::true_type{});
      GVIT_TR(13);
      tc_fence_before();
      mbar_arrive(&ctl->accum_free);                        // this thread's accumulator row is in registers (keys)
      if (row < Np) {
        const int64_t o = ((int64_t)b * Np + row) * k;
#pragma unroll
        for (int s = 0; s < KT; ++s)
          if (s < k) {
            idx[o + s] = 255 - (int)(key[s] & 0xffu);
            vals[o + s] = (float)((int)(key[s] >> 8) - 4194304) * (1.0f / KEY_SCALE);
          }
      }
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
  if (k <= 4) return launch<4>(tmap, t, k, NT, idx, vals, rnorm, st);
  if (k <= 8) return launch<8>(tmap, t, k, NT, idx, vals, rnorm, st);
  if (k <= 16) return launch<16>(tmap, t, k, NT, idx, vals, rnorm, st);
  return launch<32>(tmap, t, k, NT, idx, vals, rnorm, st);
}

}  // namespace gvit

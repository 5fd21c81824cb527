// agg_tc.cu - fused aggregation G4+G5+G6 + residual for bf16 (north_star: "gather /
// normalise, then a tcgen05 GEMM with TMA-staged tiles that keeps A.X in shared memory with no HBM
// round-trip").
//
//   out[b,1+i,:] = resid[b,1+i,:] + ( sum_j softmax_k(vals)_ij * p[b, idx_ij, :] ) Wg^T + bias
//
// One CTA owns 128 token rows of one image and NC (<= 256) output features.  The reduction dimension
// (the D input features) streams in 64-wide slabs; per slab TMA stages (a) the image's token slab
// [Np][64] and (b) the weight slab Wg[n0:n0+NC][64].  The gather warps read neighbour rows out of the
// staged token slab (shared memory, not HBM/L2), form the aggregated bf16 tile Z[128][64] directly in
// the 128B-swizzled K-major layout tcgen05 wants, and the MMA warp multiplies it with the weight slab
// into a TMEM accumulator.  Z only ever exists in shared memory (training optionally streams a copy out
// for the weight gradient).  Bias, residual and the CLS-row pass-through are folded into the epilogue.
//
// Warp roles: 0-7 gather + epilogue (row = tid & 127, half of the slab's columns = tid >> 7),
//             8 TMA producer, 9 MMA issuer (+ TMEM allocation).
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int THREADS = 320;
constexpr int ZS_BYTES = 128 * 128;
constexpr int MAX_STAGES = 4;

struct __align__(8) Ctrl {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], zs_full[2], zs_free[2], y_full;
  uint32_t tmem_base;
};

struct Params {
  int Np, D, k, NT, NC, stages, stage_bytes, tmem_cols;
  const int32_t* idx;
  const float* vals;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid;
  __nv_bfloat16* out;
  float* w_save;
  __nv_bfloat16* z_save;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int KT>
__global__ void __launch_bounds__(THREADS, 1) agg_tc_kernel(const __grid_constant__ CUtensorMap tm_tok,
                                                            const __grid_constant__ CUtensorMap tm_w, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sZ = sm;                                        // 2 x [128][64] bf16
  uint8_t* sStage = sm + 2 * ZS_BYTES;                     // stages x (token slab | weight slab)
  Ctrl* ctl = reinterpret_cast<Ctrl*>(sStage + (size_t)P.stages * P.stage_bytes);
  const int tok_bytes = P.NT * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nchunk = blockIdx.y, b = blockIdx.z;
  const int n0 = nchunk * P.NC;
  const int slabs = P.D / 64;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tm_tok);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < P.stages; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    for (int z = 0; z < 2; ++z) { mbar_init(&ctl->zs_full[z], 256); mbar_init(&ctl->zs_free[z], 1); }
    mbar_init(&ctl->y_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&ctl->tmem_base, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tY = ctl->tmem_base;

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < slabs; ++s) {
        const int st = s % P.stages;
        mbar_wait(&ctl->empty[st], ((s / P.stages) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[st], (uint32_t)(tok_bytes + P.NC * 128));
        uint8_t* dst = sStage + (size_t)st * P.stage_bytes;
        tma_load_3d(dst, &tm_tok, s * 64, 0, b, &ctl->full[st]);
        tma_load_3d(dst + tok_bytes, &tm_w, s * 64, n0, 0, &ctl->full[st]);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, P.NC, false, false);
      for (int s = 0; s < slabs; ++s) {
        const int st = s % P.stages, zb = s & 1;
        mbar_wait(&ctl->full[st], (s / P.stages) & 1);       // weight slab landed
        mbar_wait(&ctl->zs_full[zb], (s >> 1) & 1);          // aggregated tile staged (token slab fully read)
        tc_fence_after();
        const uint32_t aZ = smem_u32(sZ + zb * ZS_BYTES);
        const uint32_t aW = smem_u32(sStage + (size_t)st * P.stage_bytes + tok_bytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tY, make_sdesc(aZ + kk * 32), make_sdesc(aW + kk * 32), idesc, s > 0 || kk > 0);
        umma_commit(&ctl->empty[st]);
        umma_commit(&ctl->zs_free[zb]);
      }
      umma_commit(&ctl->y_full);
    }
  } else {
    const int row = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int rowg = mt * 128 + row;
    const bool valid = rowg < P.Np;
    // ---- G4: softmax over the k selected similarities (fp32), rounded to bf16 for the GEMM-shaped G5
    int nb[KT];
    float w[KT];
    {
      float mx = -FLT_MAX, sum = 0.f;
      const int64_t o = ((int64_t)b * P.Np + rowg) * P.k;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const bool on = valid && j < P.k;
        nb[j] = on ? P.idx[o + j] : 0;
        w[j] = on ? P.vals[o + j] : -FLT_MAX;
        mx = fmaxf(mx, w[j]);
      }
#pragma unroll
      for (int j = 0; j < KT; ++j) { w[j] = (valid && j < P.k) ? expf(w[j] - mx) : 0.f; sum += w[j]; }
      const float inv = valid ? 1.0f / sum : 0.f;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        w[j] *= inv;
        if (P.w_save && nchunk == 0 && half == 0 && valid && j < P.k) P.w_save[o + j] = w[j];
        w[j] = __bfloat162float(__float2bfloat16_rn(w[j]));
      }
    }
    // CLS row: out[b,0,n0:n0+NC] = resid[b,0,...] (graph leaves CLS untouched)
    if (mt == 0 && threadIdx.x < P.NC / 8) {
      const int64_t o = (int64_t)b * (P.Np + 1) * P.D + n0 + threadIdx.x * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (P.resid) v = *reinterpret_cast<const uint4*>(P.resid + o);
      *reinterpret_cast<uint4*>(P.out + o) = v;
    }
    // ---- G5: per slab, gather neighbour rows from the staged token slab into the swizzled A tile
    for (int s = 0; s < slabs; ++s) {
      const int st = s % P.stages, zb = s & 1;
      mbar_wait(&ctl->full[st], (s / P.stages) & 1);
      mbar_wait(&ctl->zs_free[zb], ((s >> 1) & 1) ^ 1);
      const uint8_t* tok = sStage + (size_t)st * P.stage_bytes;
      uint8_t* zt = sZ + zb * ZS_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = (half * 4 + c) * 8;
        float acc[8] = {};
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          if (j < P.k) {
            const uint4 raw = *reinterpret_cast<const uint4*>(tok + swz128(nb[j], col));
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hp[e]);
              acc[2 * e] = fmaf(w[j], f.x, acc[2 * e]);
              acc[2 * e + 1] = fmaf(w[j], f.y, acc[2 * e + 1]);
            }
          }
        }
        uint4 zq;
        zq.x = pack2(acc[0], acc[1]); zq.y = pack2(acc[2], acc[3]); zq.z = pack2(acc[4], acc[5]); zq.w = pack2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(zt + swz128(row, col)) = zq;
        if (P.z_save && nchunk == 0 && valid)
          *reinterpret_cast<uint4*>(P.z_save + ((int64_t)b * P.Np + rowg) * P.D + s * 64 + col) = zq;
      }
      fence_async_smem();
      mbar_arrive(&ctl->zs_full[zb]);
    }
    // ---- G6 epilogue: + bias + residual, bf16, straight to HBM
    mbar_wait(&ctl->y_full, 0);
    tc_fence_after();
    const int ewarp = warp & 3, chalf = warp >> 2;
    const int erow = mt * 128 + ewarp * 32 + lane;
    const uint32_t lY = tmem_lane_base(tY, warp);
    const int cw = P.NC / 2;
    for (int c0 = chalf * cw; c0 < (chalf + 1) * cw; c0 += 32) {
      float v[32];
      tmem_ld32(lY + c0, v);
      if (erow < P.Np) {
        const int64_t o = ((int64_t)b * (P.Np + 1) + 1 + erow) * P.D + n0 + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float r[8] = {}, bi[8] = {};
          if (P.resid) load8(P.resid + o + 8 * q, r);
          if (P.bias) load8(P.bias + n0 + c0 + 8 * q, bi);
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = v[8 * q + e] + bi[e] + r[e];
          store8(P.out + o + 8 * q, y);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tY, P.tmem_cols);
}

inline int pick_nc(int D) {
  for (int nc : {256, 192, 128, 64})
    if (D % nc == 0) return nc;
  return 0;
}

template <int KT>
int launch(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const Params& P, int B, size_t smem, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(agg_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((P.Np + 127) / 128, P.D / P.NC, B);
  agg_tc_kernel<KT><<<grid, THREADS, smem, st>>>(tm_tok, tm_w, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

bool agg_tc_supported(int Np, int D, int k) {
  return Np >= 16 && Np <= 256 && D >= 64 && D % 64 == 0 && D <= 1024 && k <= 16 && pick_nc(D) != 0;
}

int agg_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
               const void* bias, const void* resid, void* out, float* w_save, void* z_save, cudaStream_t st) {
  Params P;
  P.Np = Np; P.D = D; P.k = k;
  P.NT = (Np + 15) & ~15;
  P.NC = pick_nc(D);
  P.stage_bytes = P.NT * 128 + P.NC * 128;
  const int budget = 224 * 1024 - 1024 - 2 * ZS_BYTES - (int)sizeof(Ctrl);
  P.stages = budget / P.stage_bytes;
  if (P.stages > MAX_STAGES) P.stages = MAX_STAGES;
  GVIT_REQUIRE(P.stages >= 2, GVIT_ERR_SHAPE, "agg_fwd: stage of %d bytes does not fit twice in shared memory", P.stage_bytes);
  P.tmem_cols = P.NC <= 64 ? 64 : (P.NC <= 128 ? 128 : 256);
  P.idx = idx; P.vals = vals;
  P.bias = static_cast<const __nv_bfloat16*>(bias);
  P.resid = static_cast<const __nv_bfloat16*>(resid);
  P.out = static_cast<__nv_bfloat16*>(out);
  P.w_save = w_save;
  P.z_save = static_cast<__nv_bfloat16*>(z_save);

  CUtensorMap tm_tok, tm_w;
  const __nv_bfloat16* tok = static_cast<const __nv_bfloat16*>(h) + D;   // skip the CLS row (section 9, G0)
  int rc = make_tmap_bf16_3d(&tm_tok, tok, D, Np, B, D, (uint64_t)(Np + 1) * D, P.NT);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_w, Wg, D, D, 1, D, (uint64_t)D * D, P.NC);
  if (rc != GVIT_OK) return rc;
  const size_t smem = 1024 + 2 * ZS_BYTES + (size_t)P.stages * P.stage_bytes + sizeof(Ctrl);
  if (k <= 4) return launch<4>(tm_tok, tm_w, P, B, smem, st);
  if (k <= 8) return launch<8>(tm_tok, tm_w, P, B, smem, st);
  return launch<16>(tm_tok, tm_w, P, B, smem, st);
}

}  // namespace gvit

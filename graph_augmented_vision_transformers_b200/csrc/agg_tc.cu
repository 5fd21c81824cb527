// agg_tc.cu - fused aggregation G4+G5+G6 + residual for bf16 (north_star: "gather / normalise,
// then a tcgen05 GEMM with TMA-staged tiles that keeps A.X in shared memory or TMEM with no HBM round-trip").
//
//   out[b,1+i,:] = resid[b,1+i,:] + ( sum_j softmax_k(vals)_ij * p[b, idx_ij, :] ) Wg^T + bias
//
// Both products run on the tensor cores.  A CTA owns 128 token rows of one image and NC (<= 384) output features:
//   prologue : the softmax weights of its rows are scattered into a dense bf16 adjacency tile A~ [128][NT] in shared
//              memory (k non-zeros per row) - "A" of A.X.W, built once per CTA;
//   per 64-feature slab s of the reduction dimension (TMA stages the image's token slab P_s [NT][64] and the
//   weight slab W_s [NC][64]):
//       Z_s  = A~ . P_s          tcgen05.mma, M=128 N=64 K=NT, fp32 in TMEM        (the "gather", as a GEMM)
//       Z_s -> bf16, in place    tcgen05.ld / tcgen05.st by the 4 convert warps     (A.X never leaves TMEM;
//                                                                                    training also streams it out
//                                                                                    for the weight gradient)
//       OUT += Z_s . W_s^T       tcgen05.mma with A read from TMEM, N = NC, fp32 in TMEM
//   epilogue : + bias + residual, bf16, straight to HBM; the CLS row is passed through.
// The dense A~.P product does ~NT/k times the FLOPs of a sparse gather but costs 416 tensor-pipe cycles per slab
// against 768+ for the projection, replaces a bank-conflicted shared-memory gather that was 6x slower than the MMAs
// it fed, and is exactly what the dense-adjacency mode (BASELINE config 4) needs.
//
// TMEM (512 columns): OUT [0, NC) | Z ping [384, 448) | Z pong [448, 512).
// Warp roles: 0-3 adjacency build + Z conversion + epilogue (thread <-> row = TMEM lane), 4 TMA producer,
//             5 MMA issuer (+ TMEM allocation).
#include <float.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {

using namespace tc;

constexpr int THREADS = 192;
constexpr int TILE = 128 * 128;             // [128 rows][64 bf16]
constexpr int A_BYTES = 4 * TILE;           // A~ [128][256] as four 64-column blocks
constexpr int TOK_BYTES = 2 * TILE;         // token slab, up to [256][64]
constexpr int W_BYTES = 3 * TILE;           // weight slab, up to [384][64]
constexpr int STAGE_BYTES = TOK_BYTES + W_BYTES;
constexpr int TMEM_Z = 384;

struct __align__(8) Ctrl {
  uint64_t full[2], empty[2], a_ready, z_full[2], zb_ready[2], out_full;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 + A_BYTES + 2 * STAGE_BYTES + sizeof(Ctrl);

struct Params {
  int Np, D, k, NT, NC, NH, nsplit;        // NC output features per CTA, issued as nsplit MMAs of NH columns
  const int32_t* idx;
  const float* vals;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid;
  __nv_bfloat16* out;
  float* w_save;
  __nv_bfloat16* z_save;
  int64_t zbs;                             // batch stride of z_save in elements (rows are D apart)
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int KT>
__global__ void __launch_bounds__(THREADS, 1) agg_tc_kernel(const __grid_constant__ CUtensorMap tm_tok,
                                                            const __grid_constant__ CUtensorMap tm_w, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // no static __shared__ in these kernels: base is 1024-aligned
  uint8_t* sm = smem_raw;   // NOT rounded through an integer: that made every access a generic LD/ST instead of LDS/STS
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  uint8_t* sA = sm;
  uint8_t* sStage = sm + A_BYTES;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(sStage + 2 * STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role id
  const int mt = blockIdx.x, nchunk = blockIdx.y, b = blockIdx.z;
  const int n0 = nchunk * P.NC;
  const int slabs = P.D / 64;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tm_tok);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
      mbar_init(&ctl->z_full[s], 1);
      mbar_init(&ctl->zb_ready[s], 128);
    }
    mbar_init(&ctl->a_ready, 128);
    mbar_init(&ctl->out_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 4) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      for (int s = 0; s < slabs; ++s) {
        const int st = s & 1;
        mbar_wait(&ctl->empty[st], ((s >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->full[st], (uint32_t)((P.NT + P.NC) * 128));
        uint8_t* dst = sStage + st * STAGE_BYTES;
        tma_load_3d(dst, &tm_tok, s * 64, 0, b, &ctl->full[st]);
        for (int j = 0; j < P.nsplit; ++j)
          tma_load_3d(dst + TOK_BYTES + j * P.NH * 128, &tm_w, s * 64, n0 + j * P.NH, 0, &ctl->full[st]);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {   // ONE elected thread runs the whole role loop (see tc.cuh)
      const uint32_t idesc_z = make_idesc(128, 64, false, true);       // A~ K-major, token slab MN-major
      const uint32_t idesc_w = make_idesc(128, P.NH, false, false);    // Z from TMEM, W slab K-major
      const uint32_t aA = smem_u32(sA);
      auto issue_w = [&](int t) {                                      // OUT += Z_t . W_t^T
        const int st = t & 1;
        mbar_wait(&ctl->zb_ready[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t aW = smem_u32(sStage + st * STAGE_BYTES + TOK_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          for (int j = 0; j < P.nsplit; ++j)
            umma_ts(tmem + j * P.NH, tmem + TMEM_Z + st * 64 + kk * 8, make_sdesc(aW + j * P.NH * 128 + kk * 32), idesc_w,
                    t > 0 || kk > 0);
        umma_commit(&ctl->empty[st]);                                  // slab stage (and Z buffer) reusable
      };
      mbar_wait(&ctl->a_ready, 0);
      tc_fence_after();
      for (int s = 0; s < slabs; ++s) {
        const int st = s & 1;
        mbar_wait(&ctl->full[st], (s >> 1) & 1);
        tc_fence_after();
        const uint32_t aTok = smem_u32(sStage + st * STAGE_BYTES);
        for (int ks = 0; ks < P.NT / 16; ++ks)                         // Z_s = A~ . P_s
          umma_ss(tmem + TMEM_Z + st * 64, make_sdesc(aA + (ks >> 2) * TILE + (ks & 3) * 32), make_sdesc(aTok + ks * 2048),
                  idesc_z, ks > 0);
        umma_commit(&ctl->z_full[st]);
        if (s > 0) issue_w(s - 1);                                     // overlaps the conversion of Z_s
      }
      issue_w(slabs - 1);
      umma_commit(&ctl->out_full);
    }
  } else {
    const int row = threadIdx.x;                                       // 0..127 == TMEM lane
    const int rowg = mt * 128 + row;
    const bool valid = rowg < P.Np;
    // ---- G4 + adjacency tile: zero A~, then scatter this row's k softmax weights (bf16) at its neighbour columns
    {
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < A_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = z4;
      asm volatile("bar.sync 1, 128;" ::: "memory");
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
            if (P.w_save && nchunk == 0) P.w_save[o + j] = wj;
            const int c = nb[j];
            *reinterpret_cast<__nv_bfloat16*>(sA + (c >> 6) * TILE + swz128(row, c & 63) + (c & 7) * 2) = __float2bfloat16_rn(wj);
          }
        }
      }
      fence_async_smem();
      mbar_arrive(&ctl->a_ready);
    }
    // CLS row: out[b,0,n0:n0+NC] = resid[b,0,...] (the graph leaves CLS untouched, section 9 G0)
    if (mt == 0 && threadIdx.x < P.NC / 8) {
      const int64_t o = (int64_t)b * (P.Np + 1) * P.D + n0 + threadIdx.x * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (P.resid) v = *reinterpret_cast<const uint4*>(P.resid + o);
      *reinterpret_cast<uint4*>(P.out + o) = v;
    }
    // ---- per slab: Z_s fp32 -> bf16 in place (the A operand of the projection), optional copy for the backward
    const uint32_t tZ = tmem_lane_base(tmem + TMEM_Z, warp);
    for (int s = 0; s < slabs; ++s) {
      const int st = s & 1;
      mbar_wait(&ctl->z_full[st], (s >> 1) & 1);
      tc_fence_after();
      float v0[32], v1[32];
      tmem_ld32(tZ + st * 64, v0);
      tmem_ld32(tZ + st * 64 + 32, v1);
      uint32_t pk0[16], pk1[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) { pk0[t] = pack2(v0[2 * t], v0[2 * t + 1]); pk1[t] = pack2(v1[2 * t], v1[2 * t + 1]); }
      tmem_st16(tZ + st * 64, pk0);
      tmem_st16(tZ + st * 64 + 16, pk1);
      if (P.z_save && nchunk == 0 && valid) {
        uint4* dst = reinterpret_cast<uint4*>(P.z_save + (int64_t)b * P.zbs + (int64_t)rowg * P.D + s * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dst[q] = make_uint4(pk0[4 * q], pk0[4 * q + 1], pk0[4 * q + 2], pk0[4 * q + 3]);
          dst[4 + q] = make_uint4(pk1[4 * q], pk1[4 * q + 1], pk1[4 * q + 2], pk1[4 * q + 3]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ctl->zb_ready[st]);
    }
    // ---- G6 epilogue: + bias + residual, bf16, straight to HBM
    mbar_wait(&ctl->out_full, 0);
    tc_fence_after();
    const uint32_t tO = tmem_lane_base(tmem, warp);
    for (int c0 = 0; c0 < P.NC; c0 += 32) {
      float v[32];
      tmem_ld32(tO + c0, v);
      if (valid) {
        const int64_t o = ((int64_t)b * (P.Np + 1) + 1 + rowg) * P.D + n0 + c0;
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
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// output features per CTA: the largest NC <= 384 dividing D, issued as one MMA (NC <= 256) or two (NC = 384 = 2 x 192)
inline bool pick_nc(int D, int* NC, int* NH, int* nsplit) {
  for (int nc : {384, 256, 192, 128, 64}) {
    if (D % nc) continue;
    *NC = nc;
    *nsplit = nc > 256 ? 2 : 1;
    *NH = nc / *nsplit;
    return true;
  }
  return false;
}

template <int KT>
int launch(const CUtensorMap& tm_tok, const CUtensorMap& tm_w, const Params& P, int B, cudaStream_t st) {
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(agg_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  dim3 grid((P.Np + 127) / 128, P.D / P.NC, B);
  agg_tc_kernel<KT><<<grid, THREADS, SMEM_BYTES, st>>>(tm_tok, tm_w, P);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace

bool agg_tc_supported(int Np, int D, int k) {
  int nc, nh, ns;
  return Np >= 16 && Np <= 256 && D >= 64 && D % 64 == 0 && D <= 1024 && k <= 16 && pick_nc(D, &nc, &nh, &ns);
}

int agg_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
               const void* bias, const void* resid, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
               cudaStream_t st) {
  Params P;
  P.Np = Np; P.D = D; P.k = k;
  P.NT = (Np + 15) & ~15;
  GVIT_REQUIRE(pick_nc(D, &P.NC, &P.NH, &P.nsplit), GVIT_ERR_SHAPE, "agg_fwd: D=%d has no supported column split", D);
  GVIT_REQUIRE(B <= 65535, GVIT_ERR_SHAPE, "agg_fwd: batch %d exceeds the grid limit 65535", B);
  P.idx = idx; P.vals = vals;
  P.bias = static_cast<const __nv_bfloat16*>(bias);
  P.resid = static_cast<const __nv_bfloat16*>(resid);
  P.out = static_cast<__nv_bfloat16*>(out);
  P.w_save = w_save;
  P.z_save = static_cast<__nv_bfloat16*>(z_save);
  P.zbs = z_batch_stride;

  CUtensorMap tm_tok, tm_w;
  const __nv_bfloat16* tok = static_cast<const __nv_bfloat16*>(h) + D;   // skip the CLS row (section 9, G0)
  int rc = make_tmap_bf16_3d(&tm_tok, tok, D, Np, B, D, (uint64_t)(Np + 1) * D, P.NT);
  if (rc != GVIT_OK) return rc;
  rc = make_tmap_bf16_3d(&tm_w, Wg, D, D, 1, D, (uint64_t)D * D, P.NH);
  if (rc != GVIT_OK) return rc;
  if (k <= 4) return launch<4>(tm_tok, tm_w, P, B, st);
  if (k <= 8) return launch<8>(tm_tok, tm_w, P, B, st);
  return launch<16>(tm_tok, tm_w, P, B, st);
}

}  // namespace gvit

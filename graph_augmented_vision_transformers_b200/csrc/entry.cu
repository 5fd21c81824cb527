// entry.cu - the extern "C" surface declared in include/gvit.h: argument validation and routing.
// bf16 requests go to the tcgen05/TMEM/TMA kernels whenever the shape is in their range and to the
// fp32-FMA kernels (bf16 storage, fp32 arithmetic) otherwise; fp32 requests always run exact FMA.
#include <stdlib.h>

#include "kernels.cuh"

using namespace gvit;

namespace {

inline int check_dtype(int dtype, const char* who) {
  if (dtype != GVIT_F32 && dtype != GVIT_BF16) return fail(GVIT_ERR_DTYPE, "%s: dtype %d is not GVIT_F32/GVIT_BF16", who, dtype);
  return GVIT_OK;
}

inline int check_tokens(const char* who, const void* p, int64_t bs, int64_t rs, int B, int Np, int D, int k) {
  GVIT_REQUIRE(p != nullptr, GVIT_ERR_SHAPE, "%s: null token pointer", who);
  GVIT_REQUIRE(B >= 1 && Np >= 1 && D >= 8, GVIT_ERR_SHAPE, "%s: bad sizes B=%d Np=%d D=%d", who, B, Np, D);
  GVIT_REQUIRE(D % 8 == 0 && D <= 1024, GVIT_ERR_SHAPE, "%s: D=%d must be a multiple of 8 and <= 1024", who, D);
  GVIT_REQUIRE(k >= 1 && k <= GVIT_MAX_K && k <= Np, GVIT_ERR_SHAPE, "%s: k=%d out of range [1, min(Np=%d, %d)]", who, k, Np, GVIT_MAX_K);
  GVIT_REQUIRE(rs >= D && rs % 8 == 0 && bs % 8 == 0 && bs >= 0, GVIT_ERR_ALIGN, "%s: strides (%lld, %lld) must be multiples of 8 elements", who, (long long)bs, (long long)rs);
  GVIT_REQUIRE(aligned16(p), GVIT_ERR_ALIGN, "%s: token pointer must be 16-byte aligned", who);
  return GVIT_OK;
}

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != GVIT_OK) return rc__; \
  } while (0)

}  // namespace

extern "C" {

int gvit_knn_fwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 int32_t* idx, float* vals, float* rnorm, void* stream) {
  TRY(check_dtype(dtype, "knn_fwd"));
  TRY(check_tokens("knn_fwd", p, batch_stride, row_stride, B, Np, D, k));
  GVIT_REQUIRE(idx && vals && rnorm, GVIT_ERR_SHAPE, "knn_fwd: null output");
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GVIT_BF16 && knn_tc_supported(Np, D, k)) return knn_fwd_tc(t, k, idx, vals, rnorm, st);
  return knn_fwd_simt(t, k, dtype, idx, vals, rnorm, st);
}

int gvit_graph_reverse(const int32_t* idx, int B, int Np, int k, int32_t* rev_ptr, int32_t* rev_src, void* stream) {
  GVIT_REQUIRE(idx && rev_ptr && rev_src, GVIT_ERR_SHAPE, "graph_reverse: null pointer");
  GVIT_REQUIRE(B >= 1 && Np >= 1 && k >= 1 && k <= GVIT_MAX_K && k <= Np, GVIT_ERR_SHAPE, "graph_reverse: bad sizes B=%d Np=%d k=%d", B, Np, k);
  return graph_reverse(idx, B, Np, k, rev_ptr, rev_src, static_cast<cudaStream_t>(stream));
}

int gvit_knn_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 const int32_t* idx, const float* rnorm, const float* dvals, const int32_t* rev_ptr,
                 const int32_t* rev_src, void* dp, void* stream) {
  TRY(check_dtype(dtype, "knn_bwd"));
  TRY(check_tokens("knn_bwd", p, batch_stride, row_stride, B, Np, D, k));
  GVIT_REQUIRE(idx && rnorm && dvals && rev_ptr && rev_src && dp, GVIT_ERR_SHAPE, "knn_bwd: null pointer");
  GVIT_REQUIRE(aligned16(dp), GVIT_ERR_ALIGN, "knn_bwd: dp must be 16-byte aligned");
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return knn_bwd_simt(t, k, dtype, idx, rnorm, dvals, rev_ptr, rev_src, dp, static_cast<cudaStream_t>(stream));
}

int gvit_agg_gather_fwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k,
                        int dtype, const int32_t* idx, const float* vals, float* w, void* z, void* stream) {
  TRY(check_dtype(dtype, "agg_gather_fwd"));
  TRY(check_tokens("agg_gather_fwd", p, batch_stride, row_stride, B, Np, D, k));
  GVIT_REQUIRE(idx && vals && w && z, GVIT_ERR_SHAPE, "agg_gather_fwd: null pointer");
  GVIT_REQUIRE(aligned16(z), GVIT_ERR_ALIGN, "agg_gather_fwd: z must be 16-byte aligned");
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return agg_gather_fwd_simt(t, k, dtype, idx, vals, w, z, static_cast<cudaStream_t>(stream));
}

int gvit_agg_fwd(const void* h, int B, int Np, int D, int k, int dtype, const int32_t* idx, const float* vals,
                 const void* Wg, const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save,
                 int64_t z_batch_stride, void* stream) {
  TRY(check_dtype(dtype, "agg_fwd"));
  TRY(check_dtype(resid_dtype, "agg_fwd"));
  GVIT_REQUIRE(h && idx && vals && Wg && out, GVIT_ERR_SHAPE, "agg_fwd: null pointer");
  GVIT_REQUIRE(B >= 1 && Np >= 1 && k >= 1 && k <= GVIT_MAX_K && k <= Np, GVIT_ERR_SHAPE, "agg_fwd: bad sizes B=%d Np=%d k=%d", B, Np, k);
  GVIT_REQUIRE(dtype == GVIT_BF16, GVIT_ERR_UNSUPPORTED,
               "agg_fwd: the fused kernel is bf16-only; fp32 composes gvit_agg_gather_fwd with a library GEMM");
  GVIT_REQUIRE(agg3_tc_supported(Np, D, k) || agg_tc_supported(Np, D, k), GVIT_ERR_UNSUPPORTED,
               "agg_fwd: shape Np=%d D=%d k=%d outside the fused kernels' range", Np, D, k);
  GVIT_REQUIRE(aligned16(h) && aligned16(Wg) && aligned16(out) && (!resid || aligned16(resid)) && (!z_save || aligned16(z_save)),
               GVIT_ERR_ALIGN, "agg_fwd: pointers must be 16-byte aligned");
  GVIT_REQUIRE(!z_save || (z_batch_stride >= (int64_t)Np * D && z_batch_stride % 8 == 0), GVIT_ERR_ALIGN,
               "agg_fwd: z_batch_stride=%lld must be >= Np*D and a multiple of 8", (long long)z_batch_stride);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Default for D <= 768: one CTA pair per image with cta_group::2 MMAs (agg4_tc.cu): with the projection issued as N = 128
  // instructions it measured 0.107 ms against 0.119 ms for the one-CTA-per-row-tile kernel (agg3) at B = 256 on the same box
  // (profiles/README.md).  GVIT_AGG_NOPAIR=1 selects agg3 (A/B switch).
  static const bool use_pair = getenv("GVIT_AGG_NOPAIR") == nullptr;
  if (use_pair && Np > 128 && agg4_tc_supported(Np, D, k))                  // one row tile: a pair would idle its second CTA
    return agg4_fwd_tc(h, B, Np, D, k, idx, vals, Wg, bias, resid, resid_dtype, out, w_save, z_save, z_batch_stride, st);
  if (agg3_tc_supported(Np, D, k))
    return agg3_fwd_tc(h, B, Np, D, k, idx, vals, Wg, bias, resid, resid_dtype, out, w_save, z_save, z_batch_stride, st);
  GVIT_REQUIRE(resid_dtype == GVIT_BF16, GVIT_ERR_UNSUPPORTED,
               "agg_fwd: an fp32 residual stream (resid / out) is fused for D <= 768 only (D=%d): add the residual on the host side", D);
  return agg_fwd_tc(h, B, Np, D, k, idx, vals, Wg, bias, resid, out, w_save, z_save, z_batch_stride, st);
}

int gvit_agg_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 const int32_t* idx, const float* w, const void* dz, const int32_t* rev_ptr, const int32_t* rev_src,
                 float* dvals, void* dp, void* stream) {
  TRY(check_dtype(dtype, "agg_bwd"));
  TRY(check_tokens("agg_bwd", p, batch_stride, row_stride, B, Np, D, k));
  GVIT_REQUIRE(idx && w && dz && rev_ptr && rev_src && dvals && dp, GVIT_ERR_SHAPE, "agg_bwd: null pointer");
  GVIT_REQUIRE(aligned16(dz) && aligned16(dp), GVIT_ERR_ALIGN, "agg_bwd: dz/dp must be 16-byte aligned");
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return agg_bwd_simt(t, k, dtype, idx, w, dz, rev_ptr, rev_src, dvals, dp, static_cast<cudaStream_t>(stream));
}

int gvit_graph_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                   const int32_t* idx, const float* vals, const float* w, const float* rnorm, const void* dz,
                   int64_t dz_batch_stride, float* dvals, void* dp, void* stream) {
  TRY(check_dtype(dtype, "graph_bwd"));
  TRY(check_tokens("graph_bwd", p, batch_stride, row_stride, B, Np, D, k));
  GVIT_REQUIRE(idx && vals && w && rnorm && dz && dvals && dp, GVIT_ERR_SHAPE, "graph_bwd: null pointer");
  GVIT_REQUIRE(aligned16(dz) && aligned16(dp) && dz_batch_stride >= (int64_t)Np * D && dz_batch_stride % 8 == 0, GVIT_ERR_ALIGN,
               "graph_bwd: dz/dp must be 16-byte aligned, dz_batch_stride >= Np*D and a multiple of 8");
  GVIT_REQUIRE(dtype == GVIT_BF16 && graph_bwd_tc_supported(Np, D, k), GVIT_ERR_UNSUPPORTED,
               "graph_bwd: the fused backward is bf16-only with Np=%d D=%d k=%d in range; compose gvit_graph_reverse, "
               "gvit_agg_bwd and gvit_knn_bwd instead", Np, D, k);
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return graph_bwd_tc(t, k, idx, vals, w, rnorm, dz, dz_batch_stride, dvals, dp, static_cast<cudaStream_t>(stream));
}

int gvit_bgemm(int batch, int M, int N, int nprod,
               const void* a0, int64_t a0_rs, int64_t a0_bs, int a0_t, const void* b0, int64_t b0_rs, int64_t b0_bs, int b0_t, int K0,
               const void* a1, int64_t a1_rs, int64_t a1_bs, int a1_t, const void* b1, int64_t b1_rs, int64_t b1_bs, int b1_t, int K1,
               const float* row_scale, int out_dtype, void* out, int64_t out_rs, int64_t out_bs, void* stream) {
  TRY(check_dtype(out_dtype, "bgemm"));
  GVIT_REQUIRE(out && aligned16(out) && out_rs % 8 == 0 && out_bs % 8 == 0, GVIT_ERR_ALIGN, "bgemm: out must be 16-byte aligned with strides %% 8 == 0");
  GVIT_REQUIRE(batch >= 1 && batch <= 65535 && M >= 1 && N >= 1, GVIT_ERR_SHAPE, "bgemm: batch=%d M=%d N=%d", batch, M, N);
  BgemmProduct pr[2] = {{a0, a0_rs, a0_bs, a0_t, b0, b0_rs, b0_bs, b0_t, K0}, {a1, a1_rs, a1_bs, a1_t, b1, b1_rs, b1_bs, b1_t, K1}};
  return bgemm_tc(batch, M, N, nprod, pr, row_scale, out_dtype, out, out_rs, out_bs, static_cast<cudaStream_t>(stream));
}

static int check_dense(const char* who, int B, int Np) {
  GVIT_REQUIRE(B >= 1 && Np >= 1 && Np <= 1024, GVIT_ERR_SHAPE, "%s: B=%d Np=%d (Np <= 1024)", who, B, Np);
  return GVIT_OK;
}

int gvit_dense_rownorm(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, float* rn, void* stream) {
  TRY(check_dense("dense_rownorm", B, Np));
  GVIT_REQUIRE(p && rn && D >= 8 && D % 8 == 0 && aligned16(p) && row_stride % 8 == 0 && batch_stride % 8 == 0, GVIT_ERR_SHAPE,
               "dense_rownorm: D=%d must be a multiple of 8 with 16-byte aligned rows", D);
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return dense_rownorm(t, rn, static_cast<cudaStream_t>(stream));
}

int gvit_knn_select(const float* G, int ldg, const float* rn, int B, int Np, int k, int32_t* idx, float* vals, void* stream) {
  TRY(check_dense("knn_select", B, Np));
  GVIT_REQUIRE(G && rn && idx && vals && ldg >= Np && k >= 1 && k <= Np && k <= GVIT_MAX_K, GVIT_ERR_SHAPE,
               "knn_select: ldg=%d k=%d (ldg >= Np, 1 <= k <= min(Np, %d))", ldg, k, GVIT_MAX_K);
  return knn_select(G, ldg, rn, B, Np, k, idx, vals, static_cast<cudaStream_t>(stream));
}

int gvit_dense_softmax_fwd(const float* G, int ldg, const float* rn, int B, int Np, int ldA, void* A, void* stream) {
  TRY(check_dense("dense_softmax_fwd", B, Np));
  GVIT_REQUIRE(G && rn && A && ldg >= Np && ldA >= Np && ldA % 64 == 0 && ldA <= 1024, GVIT_ERR_SHAPE,
               "dense_softmax_fwd: ldg=%d ldA=%d (>= Np, ldA %% 64 == 0)", ldg, ldA);
  return dense_softmax_fwd(G, ldg, rn, B, Np, ldA, A, static_cast<cudaStream_t>(stream));
}

int gvit_dense_softmax_bwd(const float* dA, int ldg, const void* A, int ldA, const float* rn, int B, int Np, void* dG, void* stream) {
  TRY(check_dense("dense_softmax_bwd", B, Np));
  GVIT_REQUIRE(dA && A && rn && dG && ldg >= Np && ldA >= Np && ldA % 64 == 0 && ldA <= 1024, GVIT_ERR_SHAPE,
               "dense_softmax_bwd: ldg=%d ldA=%d (>= Np, ldA %% 64 == 0)", ldg, ldA);
  return dense_softmax_bwd(dA, ldg, A, ldA, rn, B, Np, dG, static_cast<cudaStream_t>(stream));
}

int gvit_dense_combine_bwd(const void* T, const void* V, const void* p, int64_t batch_stride, int64_t row_stride, const float* rn,
                           int B, int Np, int D, void* dp, void* stream) {
  TRY(check_dense("dense_combine_bwd", B, Np));
  GVIT_REQUIRE(T && V && p && rn && dp && D >= 8 && D % 8 == 0 && D <= 1024, GVIT_ERR_SHAPE, "dense_combine_bwd: D=%d (D %% 8 == 0, D <= 1024)", D);
  GVIT_REQUIRE(aligned16(T) && aligned16(V) && aligned16(p) && aligned16(dp) && row_stride % 8 == 0 && batch_stride % 8 == 0, GVIT_ERR_ALIGN,
               "dense_combine_bwd: 16-byte alignment required");
  Tokens t{p, batch_stride, row_stride, B, Np, D};
  return dense_combine_bwd(T, V, t, rn, dp, static_cast<cudaStream_t>(stream));
}

int gvit_attn_fwd(const void* qkv, int B, int N, int H, int dh, float scale, int dtype, void* out, float* lse,
                  void* stream) {
  TRY(check_dtype(dtype, "attn_fwd"));
  GVIT_REQUIRE(qkv && out && lse, GVIT_ERR_SHAPE, "attn_fwd: null pointer");
  GVIT_REQUIRE(B >= 1 && N >= 1 && H >= 1 && B <= 65535 && H <= 65535, GVIT_ERR_SHAPE, "attn_fwd: bad sizes B=%d N=%d H=%d", B, N, H);
  GVIT_REQUIRE(aligned16(qkv) && aligned16(out), GVIT_ERR_ALIGN, "attn_fwd: qkv/out must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GVIT_BF16 && attn_fwd_tc_supported(N, dh)) return attn_fwd_tc(qkv, B, N, H, scale, out, lse, st);
  return attn_fwd_simt(qkv, B, N, H, dh, scale, dtype, out, lse, st);
}

int gvit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, int dh,
                  float scale, int dtype, float* delta_ws, void* dqkv, void* stream) {
  TRY(check_dtype(dtype, "attn_bwd"));
  GVIT_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv, GVIT_ERR_SHAPE, "attn_bwd: null pointer");
  GVIT_REQUIRE(B >= 1 && N >= 1 && H >= 1 && B <= 65535 && H <= 65535, GVIT_ERR_SHAPE, "attn_bwd: bad sizes B=%d N=%d H=%d", B, N, H);
  GVIT_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv), GVIT_ERR_ALIGN, "attn_bwd: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GVIT_BF16 && attn_bwd_tc_supported(N, dh)) return attn_bwd_tc(qkv, out, dout, lse, B, N, H, scale, delta_ws, dqkv, st);
  return attn_bwd_simt(qkv, out, dout, lse, B, N, H, dh, scale, dtype, delta_ws, dqkv, st);
}

static int check_ln_pair(int dtype, int y_dtype, const char* who) {
  TRY(check_dtype(dtype, who));
  TRY(check_dtype(y_dtype, who));
  GVIT_REQUIRE(y_dtype == dtype || (dtype == GVIT_F32 && y_dtype == GVIT_BF16), GVIT_ERR_DTYPE,
               "%s: y_dtype must equal dtype, or be GVIT_BF16 over a GVIT_F32 stream", who);
  return GVIT_OK;
}

int gvit_layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, int dtype,
                       int y_dtype, void* y, float* mean, float* rstd, void* stream) {
  TRY(check_ln_pair(dtype, y_dtype, "layernorm_fwd"));
  GVIT_REQUIRE(x && gamma && beta && y && mean && rstd, GVIT_ERR_SHAPE, "layernorm_fwd: null pointer");
  GVIT_REQUIRE(rows >= 1 && D >= 8 && D % 8 == 0 && D <= 1024, GVIT_ERR_SHAPE, "layernorm_fwd: rows=%lld D=%d (D %% 8 == 0, D <= 1024)", (long long)rows, D);
  GVIT_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), GVIT_ERR_ALIGN, "layernorm_fwd: 16-byte alignment required");
  return layernorm_fwd(x, gamma, beta, rows, D, eps, dtype, y_dtype, y, mean, rstd, static_cast<cudaStream_t>(stream));
}

int gvit_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                       int64_t rows, int D, int dtype, int y_dtype, const void* dx_add, void* dx, float* dgamma,
                       float* dbeta, float* partial_ws, void* stream) {
  TRY(check_ln_pair(dtype, y_dtype, "layernorm_bwd"));
  GVIT_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && partial_ws, GVIT_ERR_SHAPE, "layernorm_bwd: null pointer");
  GVIT_REQUIRE(rows >= 1 && D >= 8 && D % 8 == 0 && D <= 1024, GVIT_ERR_SHAPE, "layernorm_bwd: rows=%lld D=%d (D %% 8 == 0, D <= 1024)", (long long)rows, D);
  GVIT_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma) && (!dx_add || aligned16(dx_add)), GVIT_ERR_ALIGN,
               "layernorm_bwd: 16-byte alignment required");
  return layernorm_bwd(dy, x, gamma, mean, rstd, rows, D, dtype, y_dtype, dx_add, dx, dgamma, dbeta, partial_ws,
                       static_cast<cudaStream_t>(stream));
}

int gvit_colsum(const void* x, int64_t rows, int D, int dtype, int skip_period, float* out, float* partial_ws, void* stream) {
  TRY(check_dtype(dtype, "colsum"));
  GVIT_REQUIRE(x && out && partial_ws, GVIT_ERR_SHAPE, "colsum: null pointer");
  GVIT_REQUIRE(rows >= 1 && D >= 8 && D % 8 == 0, GVIT_ERR_SHAPE, "colsum: rows=%lld D=%d (D %% 8 == 0)", (long long)rows, D);
  GVIT_REQUIRE(aligned16(x), GVIT_ERR_ALIGN, "colsum: x must be 16-byte aligned");
  GVIT_REQUIRE(skip_period >= 0, GVIT_ERR_SHAPE, "colsum: skip_period=%d must be >= 0", skip_period);
  return colsum(x, rows, D, dtype, skip_period, out, partial_ws, static_cast<cudaStream_t>(stream));
}

int gvit_dropout_residual_fwd(const void* y, const void* resid, int64_t n, float p, uint64_t seed, uint64_t offset,
                              const uint64_t* offset_dev, int dtype, int y_dtype, void* out, uint8_t* keep_mask, void* stream) {
  TRY(check_ln_pair(dtype, y_dtype, "dropout_residual_fwd"));
  GVIT_REQUIRE(y && out, GVIT_ERR_SHAPE, "dropout_residual_fwd: null pointer");
  GVIT_REQUIRE(resid || dtype == y_dtype, GVIT_ERR_DTYPE, "dropout_residual_fwd: without a residual, dtype must equal y_dtype");
  GVIT_REQUIRE(n >= 8 && n % 8 == 0, GVIT_ERR_SHAPE, "dropout_residual_fwd: n=%lld must be a positive multiple of 8", (long long)n);
  GVIT_REQUIRE(p >= 0.f && p < 1.f, GVIT_ERR_SHAPE, "dropout_residual_fwd: p=%f not in [0,1)", p);
  GVIT_REQUIRE(p == 0.f || keep_mask, GVIT_ERR_SHAPE, "dropout_residual_fwd: keep_mask required when p > 0");
  GVIT_REQUIRE(aligned16(y) && aligned16(out) && (!resid || aligned16(resid)), GVIT_ERR_ALIGN, "dropout_residual_fwd: 16-byte alignment required");
  return dropout_residual_fwd(y, resid, n, p, seed, offset, offset_dev, dtype, y_dtype, out, keep_mask, static_cast<cudaStream_t>(stream));
}

static int check_colsum_args(const char* who, int64_t n, int D, const float* colsum_out, const float* partial_ws) {
  if (!colsum_out) return GVIT_OK;
  GVIT_REQUIRE(partial_ws != nullptr, GVIT_ERR_SHAPE, "%s: colsum_out needs partial_ws", who);
  GVIT_REQUIRE(D >= 8 && D % 8 == 0 && n % D == 0, GVIT_ERR_SHAPE, "%s: D=%d must be a multiple of 8 dividing n=%lld", who, D, (long long)n);
  return GVIT_OK;
}

int gvit_dropout_bwd(const void* dout, const uint8_t* keep_mask, int64_t n, float p, int dtype, int y_dtype, void* dy,
                     int D, int skip_period, float* colsum_out, float* partial_ws, void* stream) {
  TRY(check_ln_pair(dtype, y_dtype, "dropout_bwd"));
  GVIT_REQUIRE(dout && dy && (keep_mask || p == 0.f), GVIT_ERR_SHAPE, "dropout_bwd: null pointer (keep_mask is required when p > 0)");
  GVIT_REQUIRE(n >= 8 && n % 8 == 0 && p >= 0.f && p < 1.f && (p > 0.f || colsum_out || dtype != y_dtype), GVIT_ERR_SHAPE,
               "dropout_bwd: n=%lld p=%f (p == 0 without column sums is only meaningful as the fp32 -> bf16 cast)", (long long)n, p);
  GVIT_REQUIRE(aligned16(dout) && aligned16(dy), GVIT_ERR_ALIGN, "dropout_bwd: 16-byte alignment required");
  TRY(check_colsum_args("dropout_bwd", n, D, colsum_out, partial_ws));
  GVIT_REQUIRE(skip_period >= 0, GVIT_ERR_SHAPE, "dropout_bwd: skip_period=%d must be >= 0", skip_period);
  return dropout_bwd(dout, keep_mask, n, p, dtype, y_dtype, dy, D, skip_period, colsum_out, partial_ws, static_cast<cudaStream_t>(stream));
}

int gvit_gelu_dropout_fwd(const void* u, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype,
                          void* out, uint8_t* keep_mask, void* stream) {
  TRY(check_dtype(dtype, "gelu_dropout_fwd"));
  GVIT_REQUIRE(u && out, GVIT_ERR_SHAPE, "gelu_dropout_fwd: null pointer");
  GVIT_REQUIRE(n >= 8 && n % 8 == 0, GVIT_ERR_SHAPE, "gelu_dropout_fwd: n=%lld must be a positive multiple of 8", (long long)n);
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask), GVIT_ERR_SHAPE, "gelu_dropout_fwd: p=%f (keep_mask required when p > 0)", p);
  GVIT_REQUIRE(aligned16(u) && aligned16(out), GVIT_ERR_ALIGN, "gelu_dropout_fwd: 16-byte alignment required");
  return gelu_dropout_fwd(u, n, p, seed, offset, offset_dev, dtype, out, keep_mask, static_cast<cudaStream_t>(stream));
}

int gvit_gelu_dropout_bwd(const void* dout, const void* u, const uint8_t* keep_mask, int64_t n, float p, int dtype,
                          void* du, int D, float* colsum_out, float* partial_ws, void* stream) {
  TRY(check_dtype(dtype, "gelu_dropout_bwd"));
  GVIT_REQUIRE(dout && u && du, GVIT_ERR_SHAPE, "gelu_dropout_bwd: null pointer");
  GVIT_REQUIRE(n >= 8 && n % 8 == 0, GVIT_ERR_SHAPE, "gelu_dropout_bwd: n=%lld must be a positive multiple of 8", (long long)n);
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask), GVIT_ERR_SHAPE, "gelu_dropout_bwd: p=%f (keep_mask required when p > 0)", p);
  GVIT_REQUIRE(aligned16(dout) && aligned16(u) && aligned16(du), GVIT_ERR_ALIGN, "gelu_dropout_bwd: 16-byte alignment required");
  TRY(check_colsum_args("gelu_dropout_bwd", n, D, colsum_out, partial_ws));
  return gelu_dropout_bwd(dout, u, keep_mask, n, p, dtype, du, D, colsum_out, partial_ws, static_cast<cudaStream_t>(stream));
}

int gvit_linear_gelu_dropout_fwd(const void* x, const void* w, const void* bias, int64_t M, int N, int K, float p, uint64_t seed,
                                 uint64_t offset, const uint64_t* offset_dev, int dtype, int save_mode, void* u, void* out,
                                 uint8_t* keep_mask, void* stream) {
  TRY(check_dtype(dtype, "linear_gelu_dropout_fwd"));
  GVIT_REQUIRE(x && w && out, GVIT_ERR_SHAPE, "linear_gelu_dropout_fwd: null pointer");
  GVIT_REQUIRE(save_mode == 0 || save_mode == 1, GVIT_ERR_SHAPE, "linear_gelu_dropout_fwd: save_mode=%d (0: pre-activation, 1: backward factor)", save_mode);
  // the keep mask is what the backward of save_mode 0 needs; the factor of save_mode 1 already contains it, and u == NULL saves nothing
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask || save_mode == 1 || !u), GVIT_ERR_SHAPE,
               "linear_gelu_dropout_fwd: p=%f (keep_mask required when p > 0 and the pre-activation is saved)", p);
  GVIT_REQUIRE(dtype == GVIT_BF16 && fc1_tc_supported(M, N, K), GVIT_ERR_UNSUPPORTED,
               "linear_gelu_dropout_fwd: the fused kernel is bf16-only with N %% 256 == 0 and K %% 64 == 0 (M=%lld N=%d K=%d); "
               "compose a library GEMM with gvit_gelu_dropout_fwd instead", (long long)M, N, K);
  GVIT_REQUIRE(aligned16(x) && aligned16(w) && (!u || aligned16(u)) && aligned16(out) && (!bias || aligned16(bias)) &&
               (!keep_mask || (reinterpret_cast<uintptr_t>(keep_mask) & 7u) == 0), GVIT_ERR_ALIGN,
               "linear_gelu_dropout_fwd: tensors must be 16-byte aligned (keep_mask 8-byte)");
  return fc1_gelu_dropout_fwd_tc(x, w, bias, M, N, K, p, seed, offset, offset_dev, save_mode, u, out, keep_mask,
                                 static_cast<cudaStream_t>(stream));
}

int gvit_linear_dropout_residual_fwd(const void* x, const void* w, const void* bias, const void* resid, int64_t M, int N, int K, float p,
                                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int resid_dtype, void* out,
                                     uint8_t* keep_mask, void* stream) {
  TRY(check_dtype(dtype, "linear_dropout_residual_fwd"));
  TRY(check_dtype(resid_dtype, "linear_dropout_residual_fwd"));
  GVIT_REQUIRE(x && w && resid && out, GVIT_ERR_SHAPE, "linear_dropout_residual_fwd: null pointer");
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask), GVIT_ERR_SHAPE, "linear_dropout_residual_fwd: p=%f (keep_mask required when p > 0)", p);
  GVIT_REQUIRE(dtype == GVIT_BF16 && fc1_tc_supported(M, N, K), GVIT_ERR_UNSUPPORTED,
               "linear_dropout_residual_fwd: the fused kernel is bf16-only with N %% 256 == 0 and K %% 64 == 0 (M=%lld N=%d K=%d); "
               "compose a library GEMM with gvit_dropout_residual_fwd instead", (long long)M, N, K);
  GVIT_REQUIRE(aligned16(x) && aligned16(w) && aligned16(resid) && aligned16(out) && (!bias || aligned16(bias)) &&
               (!keep_mask || (reinterpret_cast<uintptr_t>(keep_mask) & 3u) == 0), GVIT_ERR_ALIGN,
               "linear_dropout_residual_fwd: tensors must be 16-byte aligned (keep_mask 4-byte)");
  return linear_dropout_residual_fwd_tc(x, w, bias, resid, M, N, K, p, seed, offset, offset_dev, resid_dtype, out, keep_mask,
                                        static_cast<cudaStream_t>(stream));
}

int64_t gvit_linear_gelu_dropout_bwd_ws_rows(int64_t M) { return fc2_bwd_partial_rows(M); }

int gvit_linear_gelu_dropout_bwd(const void* dout, const void* w2, const void* u, const uint8_t* keep_mask, int64_t M, int N, int K,
                                 float p, int dtype, int saved_mode, void* du, float* colsum_out, float* partial_ws, void* stream) {
  TRY(check_dtype(dtype, "linear_gelu_dropout_bwd"));
  GVIT_REQUIRE(dout && w2 && u && du && colsum_out && partial_ws, GVIT_ERR_SHAPE, "linear_gelu_dropout_bwd: null pointer");
  GVIT_REQUIRE(saved_mode == 0 || saved_mode == 1, GVIT_ERR_SHAPE, "linear_gelu_dropout_bwd: saved_mode=%d (0: pre-activation, 1: backward factor)", saved_mode);
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask || saved_mode == 1), GVIT_ERR_SHAPE,
               "linear_gelu_dropout_bwd: p=%f (keep_mask required when p > 0 and u is the pre-activation)", p);
  GVIT_REQUIRE(dtype == GVIT_BF16 && fc1_tc_supported(M, N, K), GVIT_ERR_UNSUPPORTED,
               "linear_gelu_dropout_bwd: the fused kernel is bf16-only with N %% 256 == 0 and K %% 64 == 0 (M=%lld N=%d K=%d); "
               "compose a library GEMM with gvit_gelu_dropout_bwd instead", (long long)M, N, K);
  GVIT_REQUIRE(aligned16(dout) && aligned16(w2) && aligned16(u) && aligned16(du) && aligned16(partial_ws) &&
               (!keep_mask || (reinterpret_cast<uintptr_t>(keep_mask) & 3u) == 0), GVIT_ERR_ALIGN,
               "linear_gelu_dropout_bwd: tensors must be 16-byte aligned (keep_mask 4-byte)");
  return linear_gelu_dropout_bwd_tc(dout, w2, u, keep_mask, M, N, K, p, saved_mode, du, colsum_out, partial_ws,
                                    static_cast<cudaStream_t>(stream));
}

int64_t gvit_linear_gemm_ws_bytes(int64_t M, int N, int K) { return gemm2_tc_supported(M, N, K) ? gemm2_ws_bytes(M, N, K) : 0; }

int gvit_linear_gemm(const void* a, int a_t, int64_t a_rs, const void* b, int b_t, int64_t b_rs, int64_t M, int N, int K,
                     const void* bias, int out_dtype, void* out, int64_t out_rs, void* workspace, int64_t workspace_bytes,
                     void* stream) {
  TRY(check_dtype(out_dtype, "linear_gemm"));
  GVIT_REQUIRE(a && b && out, GVIT_ERR_SHAPE, "linear_gemm: null pointer");
  GVIT_REQUIRE((a_t == 0 || a_t == 1) && (b_t == 0 || b_t == 1), GVIT_ERR_SHAPE, "linear_gemm: a_t=%d b_t=%d (0 or 1)", a_t, b_t);
  GVIT_REQUIRE(gemm2_tc_supported(M, N, K), GVIT_ERR_UNSUPPORTED, "linear_gemm: M=%lld N=%d K=%d (N %% 256 == 0 required)",
               (long long)M, N, K);
  GVIT_REQUIRE(M <= 0x7fffffffLL - 256 && out_rs >= N, GVIT_ERR_SHAPE, "linear_gemm: M=%lld out_rs=%lld", (long long)M, (long long)out_rs);
  GVIT_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out) && (!bias || aligned16(bias)) && a_rs % 8 == 0 && b_rs % 8 == 0 &&
               out_rs % (out_dtype == GVIT_F32 ? 4 : 8) == 0 && (!workspace || aligned16(workspace)), GVIT_ERR_ALIGN,
               "linear_gemm: operands / rows / workspace must be 16-byte aligned");
  return gemm2_tc(a, a_t, a_rs, b, b_t, b_rs, M, N, K, bias, out_dtype, out, out_rs, static_cast<float*>(workspace), workspace_bytes,
                  static_cast<cudaStream_t>(stream));
}

int gvit_patchify(const void* img, int B, int C, int H, int W, int P, int in_dtype, int out_dtype, void* out, void* stream) {
  TRY(check_dtype(in_dtype, "patchify"));
  TRY(check_dtype(out_dtype, "patchify"));
  GVIT_REQUIRE(img && out, GVIT_ERR_SHAPE, "patchify: null pointer");
  GVIT_REQUIRE(in_dtype == out_dtype || in_dtype == GVIT_F32, GVIT_ERR_DTYPE, "patchify: a bf16 image cannot produce fp32 patches");
  GVIT_REQUIRE(B >= 1 && C >= 1 && P >= 8 && P % 8 == 0 && H >= P && W >= P && H % P == 0 && W % P == 0, GVIT_ERR_SHAPE,
               "patchify: bad sizes B=%d C=%d H=%d W=%d P=%d (P %% 8 == 0, H and W multiples of P)", B, C, H, W, P);
  GVIT_REQUIRE(aligned16(img) && aligned16(out), GVIT_ERR_ALIGN, "patchify: 16-byte alignment required");
  return patchify(img, B, C, H, W, P, in_dtype, out_dtype, out, static_cast<cudaStream_t>(stream));
}

int gvit_embed_assemble(const void* y, const void* bias, const void* cls, const void* pos, int B, int N, int D, float p,
                        uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int param_dtype, int out_dtype, void* out,
                        uint8_t* keep_mask, void* stream) {
  TRY(check_dtype(dtype, "embed_assemble"));
  TRY(check_dtype(param_dtype, "embed_assemble"));
  TRY(check_dtype(out_dtype, "embed_assemble"));
  GVIT_REQUIRE(out_dtype == dtype || (out_dtype == GVIT_F32 && param_dtype == GVIT_F32), GVIT_ERR_DTYPE,
               "embed_assemble: out is of y's dtype, or fp32 (fp32 residual stream) with fp32 parameters");
  GVIT_REQUIRE(y && cls && pos && out, GVIT_ERR_SHAPE, "embed_assemble: null pointer");
  GVIT_REQUIRE(param_dtype == dtype || param_dtype == GVIT_F32, GVIT_ERR_DTYPE, "embed_assemble: parameters must be fp32 or the token dtype");
  GVIT_REQUIRE(B >= 1 && N >= 2 && D >= 8 && D % 8 == 0, GVIT_ERR_SHAPE, "embed_assemble: bad sizes B=%d N=%d D=%d", B, N, D);
  GVIT_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || keep_mask), GVIT_ERR_SHAPE, "embed_assemble: p=%f (keep_mask required when p > 0)", p);
  GVIT_REQUIRE(aligned16(y) && aligned16(out) && aligned16(cls) && aligned16(pos) && (!bias || aligned16(bias)), GVIT_ERR_ALIGN,
               "embed_assemble: 16-byte alignment required");
  return embed_assemble(y, bias, cls, pos, B, N, D, p, seed, offset, offset_dev, dtype, param_dtype, out_dtype, out, keep_mask,
                        static_cast<cudaStream_t>(stream));
}

int gvit_mt_chunk_elems(void) { return mt_chunk_elems(); }

int gvit_mt_adamw_step(const int64_t* p, const int64_t* g, const int64_t* m, const int64_t* v, const int64_t* numel, const float* lr,
                       const float* wd, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t* tstep, int ntensors, int nchunks,
                       float max_norm, int64_t warmup_steps, int64_t total_steps, float beta1, float beta2, float eps, int64_t* step,
                       float* sched, float* partial_ws, void* stream) {
  GVIT_REQUIRE(p && g && m && v && numel && lr && wd && chunk_tensor && chunk_index && tstep && step && sched && partial_ws, GVIT_ERR_SHAPE,
               "mt_adamw_step: null pointer");
  GVIT_REQUIRE(ntensors >= 1 && nchunks >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps > 0.f, GVIT_ERR_SHAPE,
               "mt_adamw_step: nchunks=%d betas=(%f, %f) eps=%g", nchunks, beta1, beta2, eps);
  return mt_adamw_step(p, g, m, v, numel, lr, wd, chunk_tensor, chunk_index, tstep, ntensors, nchunks, max_norm, warmup_steps, total_steps,
                       beta1, beta2, eps, step, sched, partial_ws, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

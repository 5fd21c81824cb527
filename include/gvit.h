/* gvit.h - C ABI of libgvit.so, the sm_100a implementation of the graph-augmented
 * ViT hot path (patch-token kNN graph construction, adjacency-weighted
 * aggregation, multi-head self-attention, and the LayerNorm / residual edges).
 *
 * Boundary contract:
 *   - plain pointers and sizes only; no torch types.  Every pointer is a CUDA
 *     device pointer owned by the caller (PyTorch's caching allocator on the
 *     Python side); the library allocates nothing persistent and keeps no
 *     mutable global state, so calls are re-entrant (PyTorch's autograd worker
 *     thread calls the *_bwd entry points).
 *   - `stream` is a cudaStream_t passed as void*; kernels are only ever enqueued
 *     on it (never on the legacy default stream).
 *   - return value: 0 = GVIT_OK, otherwise a gvit_status; the message is
 *     available per thread from gvit_last_error_string().  There is no CPU
 *     fallback and no alternative backend: unsupported shapes fail loudly.
 *   - dtype: GVIT_F32 runs exact fp32 FMA kernels (no TF32) - the parity path;
 *     GVIT_BF16 runs the tcgen05 / TMEM / TMA kernels with fp32 accumulation.
 *
 * Each entry point cites the reference interface it replaces, as
 * /root/reference/<file>:<line>.  The graph stages have no reference symbol
 * (the reference ships no graph code); they implement SURVEY.md section 9 G0-G6.
 */
#ifndef GVIT_H_
#define GVIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVIT_ABI_VERSION 18
#if defined(__GNUC__)
#define GVIT_API __attribute__((visibility("default")))
#else
#define GVIT_API
#endif

typedef enum {
  GVIT_OK = 0,
  GVIT_ERR_SHAPE = 1,       /* a size is out of the supported range            */
  GVIT_ERR_ALIGN = 2,       /* a pointer or stride breaks the alignment rules  */
  GVIT_ERR_DTYPE = 3,       /* dtype is not GVIT_F32 / GVIT_BF16               */
  GVIT_ERR_CUDA = 4,        /* a CUDA runtime / driver call failed             */
  GVIT_ERR_UNSUPPORTED = 5  /* valid request this build has no kernel for      */
} gvit_status;

typedef enum { GVIT_F32 = 0, GVIT_BF16 = 1 } gvit_dtype;

enum { GVIT_MAX_K = 32 };   /* largest neighbour count of the kNN graph */

GVIT_API int gvit_version(void);
GVIT_API const char* gvit_last_error_string(void);

/* Describes the kernels a request would be routed to (for logs / tests):
 * writes a short NUL-terminated string such as "knn:tcgen05" into buf. */
GVIT_API int gvit_describe_path(const char* op, int dtype, int n_tokens, int dim, char* buf, int buf_len);

/* ---- a7: graph construction ---------------------------------------
 * p      : patch tokens, element (b,i,d) at p[b*batch_stride + i*row_stride + d]; pass
 *          h + row_stride for a (B,1+Np,D) token tensor so the CLS row (vit.py:207-208) is skipped.
 * idx    : (B,Np,k) int32, neighbour order = descending similarity, ties -> lowest index.
 * vals   : (B,Np,k) fp32 cosine similarities of the selected neighbours.
 * rnorm  : (B,Np) fp32, 1/max(||p_i||,1e-12); saved for gvit_knn_bwd.
 * Constraints: 1 <= k <= min(Np, GVIT_MAX_K); D % 8 == 0; bf16: Np <= 1024, 16-byte aligned rows. */
GVIT_API int gvit_knn_fwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 int32_t* idx, float* vals, float* rnorm, void* stream);

/* Reverse adjacency of idx, per image, in CSR form; deterministic (ascending edge id e = i*k+s).
 * rev_ptr : (B,Np+1) int32, rev_src : (B,Np*k) int32.  Used by the two backward entry points. */
GVIT_API int gvit_graph_reverse(const int32_t* idx, int B, int Np, int k, int32_t* rev_ptr, int32_t* rev_src, void* stream);

/* Backward of G1-G3 through the selected similarities (indices are constants):
 * dp += d vals/d p contracted with dvals.  dp has p's strides and dtype and is ACCUMULATED into
 * (gvit_agg_bwd writes it first). */
GVIT_API int gvit_knn_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 const int32_t* idx, const float* rnorm, const float* dvals, const int32_t* rev_ptr,
                 const int32_t* rev_src, void* dp, void* stream);

/* ---- a8: aggregation ----------------------------------------------
 * G4+G5 only: w = softmax_k(vals) (fp32, (B,Np,k)); z_i = sum_j w_ij p[idx_ij]  ((B,Np,D) contiguous, dtype). */
GVIT_API int gvit_agg_gather_fwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k,
                        int dtype, const int32_t* idx, const float* vals, float* w, void* z, void* stream);

/* Fused G4+G5+G6 + residual: out[b,1+i,:] = resid[b,1+i,:] + (sum_j w_ij p[idx_ij]) Wg^T + bias, with the
 * aggregated tile kept on chip (shared memory / TMEM) between the two GEMMs; out[b,0,:] = resid[b,0,:]
 * (CLS row untouched, section 9 G0).  h/resid/out are (B,1+Np,D) contiguous; Wg is (D,D) row-major
 * (nn.Linear weight: out_features x in_features); bias may be NULL; resid may be NULL (treated as 0).
 * w_save ((B,Np,k) fp32) and z_save may be NULL (inference); training saves them.  z_save holds the aggregated
 * tokens z = A~ p: row i of image b at z_save[b*z_batch_stride + i*D] - pass z_full + D with stride (1+Np)*D to lay it
 * out like h (CLS row left to the caller), which lets the weight gradient run over all (B*(1+Np)) rows without a copy.
 * bf16 only (tcgen05); D % 64 == 0, D <= 1024, Np <= 1024.  resid_dtype is the dtype of resid AND out: GVIT_BF16, or
 * GVIT_F32 for the fp32 residual stream torch.autocast keeps (vit.py:117-118 `x + f(LN(x))` with x fp32: the branch value
 * is rounded to bf16, the add is fp32); fp32 is fused for D <= 768 (GVIT_ERR_UNSUPPORTED otherwise). */
GVIT_API int gvit_agg_fwd(const void* h, int B, int Np, int D, int k, int dtype, const int32_t* idx, const float* vals,
                 const void* Wg, const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save,
                 int64_t z_batch_stride, void* stream);

/* Backward of G4+G5 given dz = dY Wg (the two GEMM gradients dWg, dz are plain library GEMMs on the host
 * side): dvals (B,Np,k) fp32 and dp (strided like p, WRITTEN, dtype). */
GVIT_API int gvit_agg_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                 const int32_t* idx, const float* w, const void* dz, const int32_t* rev_ptr, const int32_t* rev_src,
                 float* dvals, void* dp, void* stream);

/* Fused backward of G1-G5 for bf16 on the tensor cores (no reverse adjacency needed): given dz = dY Wg it writes
 * dvals (B,Np,k) fp32 - the gradient w.r.t. the selected similarities - and dp (strided like p, WRITTEN in full):
 * dp = A~^T dz  +  the gradient through the cosine similarities and the L2 normalisation.  `vals` are the saved
 * similarities of gvit_knn_fwd, `w` the saved softmax weights, `rnorm` the saved reciprocal norms; dz row i of image b
 * is at dz[b*dz_batch_stride + i*D].
 * bf16 only; returns GVIT_ERR_UNSUPPORTED outside the kernel's range (use gvit_graph_reverse + gvit_agg_bwd +
 * gvit_knn_bwd there, which also serve fp32).  gvit_describe_path("graph_bwd", ...) tells which applies. */
GVIT_API int gvit_graph_bwd(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, int k, int dtype,
                   const int32_t* idx, const float* vals, const float* w, const float* rnorm, const void* dz,
                   int64_t dz_batch_stride, float* dvals, void* dp, void* stream);

/* ---- a8, dense adjacency (SURVEY section 9 with G3 skipped: `graph_mode='dense'`, BASELINE configs[3]; the ctor path is
 * /root/reference/src/models/vit.py:125-127 with the appended keyword `graph_mode`) ------------------------------------
 * With the adjacency dense every stage is a per-image matrix product or a row-wise pass over tensors that already live in
 * HBM in the layer's own layouts; bf16 storage, fp32 arithmetic; Np <= 1024, D % 64 == 0, D <= 1024.
 *
 * gvit_bgemm: out[b] = diag(row_scale[b]) * sum_{p < nprod} op(A_p[b]) op(B_p[b]),  b < batch, nprod in {1, 2}, on tcgen05.
 *   A_p: a_t == 0 -> stored [M rows][K] (row stride a_rs, batch stride a_bs, in elements); a_t == 1 -> stored [K rows][M].
 *   B_p: b_t == 0 -> stored [N rows][K];                                                  b_t == 1 -> stored [K rows][N].
 *   The contiguous extent of every operand must be padded to a multiple of 64 elements inside its row stride, and a padded
 *   K range must hold zeros.  out: (batch, M, N) with strides out_bs / out_rs, out_dtype GVIT_BF16 or GVIT_F32; rows are
 *   written in whole 16-byte chunks, so out_rs must cover N rounded up to 8 (bf16) / 4 (fp32) elements and the tail of the
 *   last chunk is overwritten.  row_scale (batch * M fp32) may be NULL.  The second product's arguments are ignored when nprod == 1. */
GVIT_API int gvit_bgemm(int batch, int M, int N, int nprod,
               const void* a0, int64_t a0_rs, int64_t a0_bs, int a0_t, const void* b0, int64_t b0_rs, int64_t b0_bs, int b0_t, int K0,
               const void* a1, int64_t a1_rs, int64_t a1_bs, int a1_t, const void* b1, int64_t b1_rs, int64_t b1_bs, int b1_t, int K1,
               const float* row_scale, int out_dtype, void* out, int64_t out_rs, int64_t out_bs, void* stream);
/* G1: rn[b,i] = 1 / max(||p[b,i,:]||, 1e-12) for the strided patch-token view p (bf16). */
GVIT_API int gvit_dense_rownorm(const void* p, int64_t batch_stride, int64_t row_stride, int B, int Np, int D, float* rn, void* stream);
/* G3 for 256 < Np <= 1024 (e.g. the 576 patch tokens of a 384x384 image; gvit_knn_fwd fuses GEMM and selection up to 256
 * tokens): per-row top-k of S_ij = (G_ij rn_i) rn_j over the materialised fp32 Gram matrix G = P P^T (gvit_bgemm) and
 * the norms of gvit_dense_rownorm.  idx / vals as gvit_knn_fwd emits them: descending similarity, ties -> lowest index. */
GVIT_API int gvit_knn_select(const float* G, int ldg, const float* rn, int B, int Np, int k, int32_t* idx, float* vals, void* stream);
/* G2 + G4: A~[b,i,:] = softmax_j(G[b,i,j] rn[b,i] rn[b,j]) as bf16; G fp32 (B,Np,ldg) - the Gram matrix P P^T of gvit_bgemm;
 * A~ (B,Np,ldA), ldA % 64 == 0, the pad columns [Np, ldA) are written as zeros (a K-major gvit_bgemm operand). */
GVIT_API int gvit_dense_softmax_fwd(const float* G, int ldg, const float* rn, int B, int Np, int ldA, void* A, void* stream);
/* Backward of G2 + G4: dG[b,i,j] = A~_ij (dA~_ij - sum_j' dA~_ij' A~_ij') rn_i rn_j as bf16 (B,Np,ldA), pads zero. */
GVIT_API int gvit_dense_softmax_bwd(const float* dA, int ldg, const void* A, int ldA, const float* rn, int B, int Np, void* dG,
                           void* stream);
/* dp[b,i,:] = T[b,i,:] + V[b,i,:] - rn_i^2 (p_i . V_i) p_i : T = A~^T dZ, V = (dG + dG^T) P, both (B,Np,D) contiguous bf16;
 * p / dp are strided patch-token views (dp is WRITTEN).  The last term is the backward of the L2 normalisation G1. */
GVIT_API int gvit_dense_combine_bwd(const void* T, const void* V, const void* p, int64_t batch_stride, int64_t row_stride,
                           const float* rn, int B, int Np, int D, void* dp, void* stream);

/* ---- a2: attention core, replaces /root/reference/src/models/vit.py:59-69 -------------------
 * qkv : the packed projection output of vit.py:59, (B,N,3,H,dh) contiguous - consumed in place, no
 *       permute; out : (B,N,H*dh) head-major (the layout vit.py:69's transpose+reshape produces);
 * lse : (B,H,N) fp32 log-sum-exp of the scaled scores, saved for the backward.  dh in {16,32,64} (fp32 arithmetic),
 * dh == 64 (bf16).  attn_drop is 0 in every shipped config (vit.py:127) and is not implemented. */
GVIT_API int gvit_attn_fwd(const void* qkv, int B, int N, int H, int dh, float scale, int dtype, void* out, float* lse,
                  void* stream);
/* delta_ws : (B,H,N) fp32 workspace; dqkv : (B,N,3,H,dh) written in full. */
GVIT_API int gvit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, int dh,
                  float scale, int dtype, float* delta_ws, void* dqkv, void* stream);

/* ---- a5: LayerNorm + residual edges, replace nn.LayerNorm / "+" at vit.py:103,108,116-119 ----
 * rows x D, eps as nn.LayerNorm (1e-5); statistics in fp32; mean/rstd (rows) saved for the backward.
 * dtype is the type of x / gamma / beta (and dx / the residual stream); y_dtype the type of y (and dy): equal to
 * dtype, or GVIT_BF16 with dtype GVIT_F32 - the fp32-stream / bf16-branch pairing of autocast, with the cast
 * folded into the kernel. */
GVIT_API int gvit_layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, int dtype,
                       int y_dtype, void* y, float* mean, float* rstd, void* stream);
/* dgamma/dbeta are fp32 (D); partial_ws is an fp32 workspace of 2*GVIT_LN_PARTIALS*D floats.  dx_add (nullable, type
 * of dx) is added to dx: the gradient that reaches x through the residual path of vit.py:117-118, folded into this
 * pass instead of a separate add kernel. */
enum { GVIT_LN_PARTIALS = 296 };
GVIT_API int gvit_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                       int64_t rows, int D, int dtype, int y_dtype, const void* dx_add, void* dx, float* dgamma,
                       float* dbeta, float* partial_ws, void* stream);
/* Column sums out[c] = sum_r x[r*D + c] (fp32 result): the bias gradient of the nn.Linear layers of the block
 * (vit.py:50,52,83,85) from the gradient of their output.  partial_ws: GVIT_COLSUM_CHUNKS * D floats.  Deterministic.
 * skip_period > 0 leaves the rows r with r % skip_period == 0 out of the sums - the CLS rows of a (B, 1+Np, D) token
 * tensor with skip_period = 1+Np: the graph projection only sees patch rows (section 9 G0), so its bias gradient is the
 * sum over patch rows and is exactly zero when they carry no gradient.  0 = every row. */
enum { GVIT_COLSUM_CHUNKS = 1024 };
GVIT_API int gvit_colsum(const void* x, int64_t rows, int D, int dtype, int skip_period, float* out, float* partial_ws, void* stream);

/* out = resid + dropout(y, p) with a Philox-4x32-7 keep mask generated from (seed, offset) - the proj_drop +
 * residual edge of vit.py:71,117 (also pos_drop, vit.py:212, with resid NULL).  p == 0 degenerates to an add.
 * dtype: type of resid / out (the residual stream); y_dtype: type of y (the branch) - same pairing rule as LayerNorm;
 * with resid NULL both must be equal.  keep_mask: n/8 bytes, ONE BIT per element (bit j of byte i = element 8i+j),
 * may be NULL when p == 0.  n % 8 == 0.
 * offset_dev (nullable): a device uint64 that is ADDED to `offset` when the kernel runs - a captured CUDA graph bakes
 * `seed` and `offset` into the launch, so the host advances *offset_dev (by >= 2^40) between replays to get fresh masks. */
GVIT_API int gvit_dropout_residual_fwd(const void* y, const void* resid, int64_t n, float p, uint64_t seed, uint64_t offset,
                              const uint64_t* offset_dev, int dtype, int y_dtype, void* out, uint8_t* keep_mask, void* stream);
/* dy = dout * keep / (1 - p);  dout has `dtype`, dy has `y_dtype`.
 * With colsum_out != NULL the same pass also writes colsum_out[c] = sum_r dy[r*D + c] (fp32, D values; the tensor is
 * read as n/D rows of D) - the bias gradient of the Linear whose output was dropped out (vit.py:70-71, 93-94);
 * partial_ws then holds GVIT_COLSUM_CHUNKS * D floats, and p == 0 (keep_mask NULL) is allowed: a cast + column sum.
 * skip_period: as for gvit_colsum (rows left out of the column sums; dy is written for every row). */
GVIT_API int gvit_dropout_bwd(const void* dout, const uint8_t* keep_mask, int64_t n, float p, int dtype, int y_dtype, void* dy,
                     int D, int skip_period, float* colsum_out, float* partial_ws, void* stream);

/* ---- Mlp activation edge: out = dropout(GELU(u), p), exact-erf GELU - nn.GELU + nn.Dropout at vit.py:84,92 in one
 * pass; the backward recomputes GELU' from the saved pre-activation u (no activation tensor is kept).
 * keep_mask as above (may be NULL when p == 0). */
GVIT_API int gvit_gelu_dropout_fwd(const void* u, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                          int dtype, void* out, uint8_t* keep_mask, void* stream);
/* colsum_out / partial_ws / D as for gvit_dropout_bwd: the bias gradient of fc1 (vit.py:90) in the same pass. */
GVIT_API int gvit_gelu_dropout_bwd(const void* dout, const void* u, const uint8_t* keep_mask, int64_t n, float p, int dtype,
                          void* du, int D, float* colsum_out, float* partial_ws, void* stream);

/* ---- f1: fc1 of the Mlp with its epilogue fused (vit.py:90-92): u = x W^T + bias (saved pre-activation), out =
 * dropout(gelu(u), p) with the keep-mask convention above - ONE persistent tcgen05 GEMM whose epilogue warps do
 * bias + GELU + Philox + both stores under the tensor pipe's shadow (accumulators double-buffered in TMEM).
 * x (M,K) and w (N,K) row-major bf16 (w is nn.Linear's weight), bias (N) bf16 or NULL, u / out (M,N) bf16,
 * keep_mask M*N/8 bytes (8-byte aligned; NULL when p == 0).  bf16 only, N % 256 == 0, K % 64 == 0: other shapes return
 * GVIT_ERR_UNSUPPORTED (compose a library GEMM with gvit_gelu_dropout_fwd); gvit_describe_path("fc1", dtype, N, K) tells. */
/* save_mode selects what the backward gets: 0 = the pre-activation u and the keep mask (what gvit_gelu_dropout_bwd and
 * gvit_linear_gelu_dropout_bwd(saved_mode 0) consume); 1 = `u` receives the BACKWARD FACTOR keep * gelu'(u) / (1 - p)
 * (bf16) instead, computed from the same Phi(u) / exp(-u^2/2) evaluation as the activation - the backward GEMM's epilogue
 * is then one multiply (keep_mask may be NULL).  u == NULL saves nothing (inference). */
GVIT_API int gvit_linear_gelu_dropout_fwd(const void* x, const void* w, const void* bias, int64_t M, int N, int K, float p, uint64_t seed,
                                 uint64_t offset, const uint64_t* offset_dev, int dtype, int save_mode, void* u, void* out,
                                 uint8_t* keep_mask, void* stream);

/* Same GEMM with the other epilogue of the block: out = resid + dropout(x W^T + bias, p) - proj + proj_drop + the residual
 * add of vit.py:70-71,117 (and fc2 + drop + residual, vit.py:93-94,118) in one kernel.  resid / out (M,N) of resid_dtype
 * (GVIT_BF16, or GVIT_F32 = the fp32 residual stream of torch.autocast: bf16 branch value added onto an fp32 x); same
 * shape limits and keep-mask convention (4-byte aligned).  Worth it while K is small (the main loop is L2-bound at
 * ~1.07 PF/s): the host side uses it for K <= 1024. */
GVIT_API int gvit_linear_dropout_residual_fwd(const void* x, const void* w, const void* bias, const void* resid, int64_t M, int N, int K, float p,
                                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int resid_dtype, void* out,
                                     uint8_t* keep_mask, void* stream);

/* Backward counterpart for the Mlp: the input gradient of fc2 fused with the backward of drop + GELU (vit.py:91-93):
 *   du = (dout W2) * keep / (1 - p) * gelu'(u),   colsum_out[n] = sum_rows du[:, n]   (fc1's bias gradient, fp32, deterministic)
 * dout (M,K) is the gradient of fc2's output after its own dropout backward, w2 (K,N) is fc2.weight as stored ((out, in)
 * row-major: no transpose copy), u (M,N) the saved pre-activation, keep_mask as written by the forward, du (M,N) bf16.
 * partial_ws: gvit_linear_gelu_dropout_bwd_ws_rows(M) * N floats.  bf16 only, N % 256 == 0, K % 64 == 0. */
GVIT_API int64_t gvit_linear_gelu_dropout_bwd_ws_rows(int64_t M);
GVIT_API int gvit_linear_gelu_dropout_bwd(const void* dout, const void* w2, const void* u, const uint8_t* keep_mask, int64_t M, int N, int K,
                                 float p, int dtype, int saved_mode, void* du, float* colsum_out, float* partial_ws, void* stream);
/* saved_mode 1: `u` is the backward factor written by gvit_linear_gelu_dropout_fwd(save_mode 1) and du = (dout W2) * u;
 * keep_mask is not read. */

/* ---- a1 / f1: the Linear GEMMs themselves (vit.py:59 qkv, :93 fc2, :28 patch projection, and the autograd of every
 * nn.Linear of the block): out (M,N) = op(a) op(b) [+ bias] as ONE persistent 2-SM tcgen05 kernel - 256 x 256 tiles per
 * CTA pair (`tcgen05.mma.cta_group::2`), each CTA staging its 128 rows of a and its half of the tile's b columns.
 *   forward          y  = x W^T + b :  a = x (M,K),   a_t 0;  b = W (N,K),  b_t 0;  out bf16
 *   input gradient   dx = dy W      :  a = dy (M,K'), a_t 0;  b = W (K',N), b_t 1;  out bf16          (K' = out features)
 *   weight gradient  dW = dy^T x    :  a = dy (K,M),  a_t 1;  b = x (K,N),  b_t 1;  out fp32          (K = all rows)
 * *_t == 0: the operand is stored [rows][K] (K contiguous); 1: stored [K][rows] (rows contiguous); *_rs = its row stride
 * in elements.  Operands bf16, bias (N) bf16 or NULL (bf16 output only), N % 256 == 0, 16-byte aligned rows.
 * An fp32 product with fewer 256 x 256 tiles than CTA pairs (the weight gradient) is split along K when `workspace` holds
 * gvit_linear_gemm_ws_bytes(M,N,K) bytes: the pieces of a tile leave fp32 partial tiles in the workspace and a second small
 * kernel adds them in a FIXED order (deterministic).  workspace may be NULL (whole tiles per pair, fewer SMs busy); it
 * must not be shared by launches that may run concurrently. */
GVIT_API int64_t gvit_linear_gemm_ws_bytes(int64_t M, int N, int K);
GVIT_API int gvit_linear_gemm(const void* a, int a_t, int64_t a_rs, const void* b, int b_t, int64_t b_rs, int64_t M, int N, int K,
                     const void* bias, int out_dtype, void* out, int64_t out_rs, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* ---- f4: token prologue, replaces PatchEmbed (/root/reference/src/models/vit.py:25-36) and the CLS / pos_embed /
 * pos_drop lines vit.py:207-212.  A kernel == stride convolution is a GEMM over non-overlapping patches:
 * gvit_patchify re-orders the image (B,C,H,W) into the patch matrix, laid out like the token tensor: out is
 * (B, 1+Np, C*P*P) with row 0 of every image ZERO (the CLS slot) and row 1+py*(W/P)+px holding patch (py,px) in
 * (c,i,j) feature order - the order of Conv2d's weight.view(D, C*P*P).  The projection (and its weight gradient) is then
 * a plain GEMM over all B*(1+Np) rows.  in_dtype: image dtype; out_dtype: patch dtype (the fp32 -> bf16 cast of
 * autocast is folded in).  P % 8 == 0. */
GVIT_API int gvit_patchify(const void* img, int B, int C, int H, int W, int P, int in_dtype, int out_dtype, void* out, void* stream);
/* out[b,0,:] = cls + pos[0];  out[b,n,:] = y[b,n,:] + bias + pos[n] for n >= 1 (row 0 of y is ignored);  then
 * dropout(p) with the keep-mask convention of gvit_dropout_residual_fwd (mask index = element index / 8).  y / out:
 * (B,N,D); y of `dtype`, out of `out_dtype` (= dtype, or GVIT_F32 over a bf16 y with fp32 parameters: the fp32 residual
 * stream torch.autocast produces at vit.py:207-211); bias (D, nullable), cls (D), pos (N,D) of `param_dtype` (GVIT_F32
 * master parameters may feed a bf16 stream).  The backward is gvit_dropout_bwd followed by gvit_colsum over the batch. */
GVIT_API int gvit_embed_assemble(const void* y, const void* bias, const void* cls, const void* pos, int B, int N, int D, float p,
                        uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int param_dtype, int out_dtype,
                        void* out, uint8_t* keep_mask, void* stream);

/* ---- f2: the optimiser side of Trainer.train_epoch (/root/reference/src/training/trainer.py:47-56,77-87,114-118) --------------
 * clip_grad_norm_(all params, max_norm) + LambdaLR (linear warm-up, then cosine, advanced per STEP) + AdamW over two
 * parameter groups, as three multi-tensor launches with everything (norm, clip coefficient, schedule factor, step count)
 * kept on the device - capturable in a CUDA graph, and the clip costs no extra pass over the gradients.
 * Device tables, one entry per tensor (n tensors, fp32): p / g / m / v = addresses of parameter, gradient (0 = no gradient:
 * the tensor is skipped), exp_avg, exp_avg_sq; numel; lr = its group's BASE learning rate; wd = its weight decay.
 * chunk_tensor / chunk_index: nchunks entries, chunk c covers elements [chunk_index*E, +E) of tensor chunk_tensor with
 * E = gvit_mt_chunk_elems().  step: device int64, completed steps (read as the scheduler epoch, then incremented).
 * tstep: n device int32, the number of updates each tensor has received (torch.optim keeps `step` per parameter: a tensor
 * without gradient does not advance); it drives AdamW's bias corrections.
 * sched: 3 device floats written by the call: total gradient norm, clip coefficient, schedule factor lambda(step).
 * max_norm <= 0 disables clipping; total_steps <= 0 keeps the learning rate constant.
 * partial_ws: nchunks floats. */
GVIT_API int gvit_mt_chunk_elems(void);
GVIT_API int gvit_mt_adamw_step(const int64_t* p, const int64_t* g, const int64_t* m, const int64_t* v, const int64_t* numel,
                       const float* lr, const float* wd, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t* tstep,
                       int ntensors, int nchunks, float max_norm, int64_t warmup_steps, int64_t total_steps, float beta1, float beta2,
                       float eps, int64_t* step, float* sched, float* partial_ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVIT_H_ */

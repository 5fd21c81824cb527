// kernels.cuh - internal launchers; entry.cu validates arguments and routes to these.
#pragma once
#include "common.cuh"

namespace gvit {

struct Tokens {            // strided view of the (B, Np, D) patch tokens inside a (B, 1+Np, D) tensor
  const void* ptr;
  int64_t batch_stride, row_stride;   // in elements
  int B, Np, D;
};

// ---- exact fp32-FMA kernels (templated on the storage type) : knn_simt.cu, graph_simt.cu, attn_simt.cu
int knn_fwd_simt(const Tokens& p, int k, int dtype, int32_t* idx, float* vals, float* rnorm, cudaStream_t st);
int graph_reverse(const int32_t* idx, int B, int Np, int k, int32_t* rev_ptr, int32_t* rev_src, cudaStream_t st);
int knn_bwd_simt(const Tokens& p, int k, int dtype, const int32_t* idx, const float* rnorm, const float* dvals,
                 const int32_t* rev_ptr, const int32_t* rev_src, void* dp, cudaStream_t st);
int agg_gather_fwd_simt(const Tokens& p, int k, int dtype, const int32_t* idx, const float* vals, float* w, void* z,
                        cudaStream_t st);
int agg_bwd_simt(const Tokens& p, int k, int dtype, const int32_t* idx, const float* w, const void* dz,
                 const int32_t* rev_ptr, const int32_t* rev_src, float* dvals, void* dp, cudaStream_t st);
int attn_fwd_simt(const void* qkv, int B, int N, int H, int dh, float scale, int dtype, void* out, float* lse,
                  cudaStream_t st);
int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H, int dh,
                  float scale, int dtype, float* delta_ws, void* dqkv, cudaStream_t st);

// ---- tcgen05 / TMEM / TMA kernels (bf16 storage, fp32 accumulate) : knn_tc.cu, agg_tc.cu, attn_tc.cu
bool knn_tc_supported(int Np, int D, int k);
int knn_fwd_tc(const Tokens& p, int k, int32_t* idx, float* vals, float* rnorm, cudaStream_t st);
bool agg_tc_supported(int Np, int D, int k);    // v2 kernel (agg_tc.cu): D up to 1024
bool agg3_tc_supported(int Np, int D, int k);   // v3 kernel (agg3_tc.cu): whole Z tile resident in TMEM, D <= 768
int agg3_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
                const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
                cudaStream_t st);
bool agg4_tc_supported(int Np, int D, int k);   // v4 kernel (agg4_tc.cu): one CTA pair per image (cta_group::2), D % 128 == 0, D <= 768
int agg4_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
                const void* bias, const void* resid, int resid_dtype, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
                cudaStream_t st);
int agg_fwd_tc(const void* h, int B, int Np, int D, int k, const int32_t* idx, const float* vals, const void* Wg,
               const void* bias, const void* resid, void* out, float* w_save, void* z_save, int64_t z_batch_stride,
               cudaStream_t st);
bool graph_bwd_tc_supported(int Np, int D, int k);
int graph_bwd_tc(const Tokens& p, int k, const int32_t* idx, const float* vals, const float* w, const float* rnorm,
                 const void* dz, int64_t dz_batch_stride, float* dvals, void* dp, cudaStream_t st);
bool graph_bwd_pair_supported(int Np, int D, int k);   // one CTA pair per image (graph_bwd_pair_tc.cu): 128 < Np <= 256, D % 128 == 0
int graph_bwd_pair_tc(const Tokens& p, int k, const int32_t* idx, const float* vals, const float* w, const float* rnorm,
                      const void* dz, int64_t dz_batch_stride, float* dvals, void* dp, cudaStream_t st);
bool attn_fwd_tc_supported(int N, int dh);
bool attn_bwd_tc_supported(int N, int dh);
int attn_fwd_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H,
                float scale, float* delta_ws, void* dqkv, cudaStream_t st);
int attn_fwd_long_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, cudaStream_t st);   // N > 256 (attn_fwd_long_tc.cu)
int attn_bwd_long_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int N, int H,
                     float scale, float* delta_ws, void* dqkv, cudaStream_t st);   // N > 256 (attn_long_tc.cu)

// ---- dense-adjacency graph layer : bgemm_tc.cu (batched tcgen05 GEMM), dense_graph.cu (row-wise stages)
struct BgemmProduct {       // one A_p B_p term.  *_t == 0: K-major (A stored [M][K], B stored [N][K]); 1: MN-major ([K][M], [K][N])
  const void* a; int64_t a_rs, a_bs; int a_t;
  const void* b; int64_t b_rs, b_bs; int b_t;
  int K;
};
int bgemm_tc(int batch, int M, int N, int nprod, const BgemmProduct* prods, const float* row_scale, int out_dtype, void* out,
             int64_t out_rs, int64_t out_bs, cudaStream_t st);
int dense_rownorm(const Tokens& t, float* rn, cudaStream_t st);
int knn_select(const float* G, int ldg, const float* rn, int B, int Np, int k, int32_t* idx, float* vals, cudaStream_t st);
int dense_softmax_fwd(const float* G, int ldg, const float* rn, int B, int Np, int ldA, void* A, cudaStream_t st);
int dense_softmax_bwd(const float* dA, int ldg, const void* A, int ldA, const float* rn, int B, int Np, void* dG, cudaStream_t st);
int dense_combine_bwd(const void* T, const void* V, const Tokens& t, const float* rn, void* dp, cudaStream_t st);

// ---- edges : edges.cu
int layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int D, float eps, int dtype,
                  int y_dtype, void* y, float* mean, float* rstd, cudaStream_t st);
int layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, int64_t rows,
                  int D, int dtype, int y_dtype, const void* dx_add, void* dx, float* dgamma, float* dbeta,
                  float* partial_ws, cudaStream_t st);
int colsum(const void* x, int64_t rows, int D, int dtype, int skip_period, float* out, float* partial_ws, cudaStream_t st);
int dropout_residual_fwd(const void* y, const void* resid, int64_t n, float p, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, int dtype, int y_dtype, void* out, uint8_t* keep_mask, cudaStream_t st);
int dropout_bwd(const void* dout, const uint8_t* keep_mask, int64_t n, float p, int dtype, int y_dtype, void* dy,
                int D, int skip_period, float* colsum_out, float* partial_ws, cudaStream_t st);
int gelu_dropout_fwd(const void* u, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype,
                     void* out, uint8_t* keep_mask, cudaStream_t st);
int gelu_dropout_bwd(const void* dout, const void* u, const uint8_t* keep_mask, int64_t n, float p, int dtype, void* du,
                     int D, float* colsum_out, float* partial_ws, cudaStream_t st);


// ---- fused fc1 : mlp_tc.cu
bool fc1_tc_supported(int64_t M, int N, int K);
int fc1_gelu_dropout_fwd_tc(const void* x, const void* w, const void* bias, int64_t M, int N, int K, float p, uint64_t seed,
                            uint64_t offset, const uint64_t* offset_dev, int save_mode, void* u, void* out, uint8_t* mask, cudaStream_t st);

int linear_dropout_residual_fwd_tc(const void* x, const void* w, const void* bias, const void* resid, int64_t M, int N, int K, float p,
                                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int resid_dtype, void* out, uint8_t* mask,
                                   cudaStream_t st);

int64_t fc2_bwd_partial_rows(int64_t M);
int linear_gelu_dropout_bwd_tc(const void* dout, const void* w2, const void* u, const uint8_t* mask, int64_t M, int N, int K, float p,
                               int saved_mode, void* du, float* colsum_out, float* partial_ws, cudaStream_t st);

// ---- 2-SM Linear GEMM : gemm2_tc.cu
bool gemm2_tc_supported(int64_t M, int N, int K);
int64_t gemm2_ws_bytes(int64_t M, int N, int K);
int gemm2_tc(const void* a, int a_t, int64_t a_rs, const void* b, int b_t, int64_t b_rs, int64_t M, int N, int K, const void* bias,
             int out_dtype, void* out, int64_t out_rs, float* ws, int64_t ws_bytes, cudaStream_t st);

// ---- optimiser step : optim.cu
int mt_chunk_elems();
int mt_adamw_step(const int64_t* p, const int64_t* g, const int64_t* m, const int64_t* v, const int64_t* numel, const float* lr,
                  const float* wd, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t* tstep, int ntensors, int nchunks,
                  float max_norm, int64_t warmup_steps, int64_t total_steps, float beta1, float beta2, float eps, int64_t* step,
                  float* sched, float* partial_ws, cudaStream_t st);

// ---- token prologue : embed.cu
int patchify(const void* img, int B, int C, int H, int W, int P, int in_dtype, int out_dtype, void* out, cudaStream_t st);
int embed_assemble(const void* y, const void* bias, const void* cls, const void* pos, int B, int N, int D, float p,
                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int param_dtype, int out_dtype, void* out,
                   uint8_t* keep_mask, cudaStream_t st);

}  // namespace gvit

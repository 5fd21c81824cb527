// api.cu - version, error reporting and path description for libgvit.so.
#include <stdarg.h>
#include <string.h>

#include "kernels.cuh"

namespace gvit {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(gvit_status s, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return static_cast<int>(s);
}

}  // namespace gvit

extern "C" {

int gvit_version(void) { return GVIT_ABI_VERSION; }

const char* gvit_last_error_string(void) { return gvit::g_err; }

int gvit_describe_path(const char* op, int dtype, int n_tokens, int dim, char* buf, int buf_len) {
  if (!op || !buf || buf_len <= 0) return gvit::fail(GVIT_ERR_SHAPE, "describe_path: null argument");
  const char* path = "unsupported";
  const bool bf16 = dtype == GVIT_BF16;
  if (dtype != GVIT_F32 && dtype != GVIT_BF16) return gvit::fail(GVIT_ERR_DTYPE, "describe_path: dtype %d", dtype);
  if (!strcmp(op, "knn")) path = (bf16 && gvit::knn_tc_supported(n_tokens, dim, 8)) ? "knn:tcgen05+tma"
                                 : (bf16 && n_tokens > 256 && n_tokens <= 1024 && dim % 64 == 0) ? "knn:tcgen05 gram + row select" : "knn:fp32-fma";
  else if (!strcmp(op, "agg")) path = (bf16 && (gvit::agg3_tc_supported(n_tokens, dim, 8) || gvit::agg_tc_supported(n_tokens, dim, 8))) ? "agg:tcgen05+tma" : "agg:gather-fma+library-gemm";
  else if (!strcmp(op, "agg_res32")) path = (bf16 && gvit::agg3_tc_supported(n_tokens, dim, 8)) ? "agg:tcgen05+tma fp32-stream" : "agg:host-side residual add";
  else if (!strcmp(op, "agg_dense")) path = (bf16 && n_tokens <= 1024 && dim % 64 == 0 && dim <= 1024) ? "agg_dense:tcgen05 batched GEMMs + row-wise softmax" : "agg_dense:ATen composition (fp32 parity path)";
  else if (!strcmp(op, "graph_bwd")) path = (bf16 && gvit::graph_bwd_tc_supported(n_tokens, dim, 8)) ? "graph_bwd:tcgen05+tma" : "graph_bwd:reverse-csr+gather";
  else if (!strcmp(op, "attn_fwd")) path = (bf16 && gvit::attn_fwd_tc_supported(n_tokens, dim)) ? "attn_fwd:tcgen05+tma" : "attn_fwd:fp32-fma";
  else if (!strcmp(op, "attn_bwd")) path = (bf16 && gvit::attn_bwd_tc_supported(n_tokens, dim)) ? "attn_bwd:tcgen05+tma" : "attn_bwd:fp32-fma";
  else if (!strcmp(op, "fc1")) path = (bf16 && gvit::fc1_tc_supported(1, n_tokens, dim)) ? "fc1:tcgen05+tma fused bias+gelu+dropout" : "fc1:library-gemm+edge";
  else if (!strcmp(op, "layernorm") || !strcmp(op, "dropout_residual")) path = "edge:vectorised";
  snprintf(buf, buf_len, "%s", path);
  return GVIT_OK;
}

}  // extern "C"

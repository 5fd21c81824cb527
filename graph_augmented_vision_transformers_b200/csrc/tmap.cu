// tmap.cu - host-side CUtensorMap construction.  cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint) once, so the library has no link-time dependency on libcuda.so and still
// loads on a machine without a GPU (the symbol-export test runs there).
#include <mutex>

#include "tc.cuh"

namespace gvit {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows) {
  EncodeTiledFn enc = resolve_encode();
  GVIT_REQUIRE(enc != nullptr, GVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  // The driver call needs a current context on THIS thread.  PyTorch's autograd worker thread has none until its first
  // runtime call that touches the device - if a libgvit backward is the first thing it runs, the encode failed with
  // CUDA_ERROR_INVALID_CONTEXT (201).  cudaSetDevice binds the primary context (once per thread).
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    GVIT_CHECK_CUDA(cudaGetDevice(&dev));
    GVIT_CHECK_CUDA(cudaSetDevice(dev));          // CUDA 12: initialises and binds the primary context; not a stream operation
    bound = true;
  }
  GVIT_REQUIRE(box_rows >= 1 && box_rows <= 256, GVIT_ERR_SHAPE, "TMA box rows %u out of range", box_rows);
  GVIT_REQUIRE(inner % 64 == 0, GVIT_ERR_SHAPE, "TMA inner extent %llu is not a multiple of 64", (unsigned long long)inner);
  GVIT_REQUIRE((row_stride_elems * 2) % 16 == 0 && (batch_stride_elems * 2) % 16 == 0 && aligned16(base), GVIT_ERR_ALIGN,
               "TMA strides/base must be 16-byte aligned");
  const cuuint64_t dims[3] = {inner, rows, batch};
  const cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {64, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GVIT_REQUIRE(r == CUDA_SUCCESS, GVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GVIT_OK;
}

}  // namespace gvit

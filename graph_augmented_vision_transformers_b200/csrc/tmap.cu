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

// esize 2: bf16, box {64, box_rows, 1}; esize 4: fp32, box {32, box_rows, 1} - 128-byte rows, 128-byte swizzle either way
static int make_tmap_3d(CUtensorMap* out, const void* base, int esize, uint64_t inner, uint64_t rows, uint64_t batch,
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
  const uint32_t box_inner = 128u / (uint32_t)esize;
  GVIT_REQUIRE(inner % box_inner == 0, GVIT_ERR_SHAPE, "TMA inner extent %llu is not a multiple of %u", (unsigned long long)inner, box_inner);
  GVIT_REQUIRE((row_stride_elems * esize) % 16 == 0 && (batch_stride_elems * esize) % 16 == 0 && aligned16(base), GVIT_ERR_ALIGN,
               "TMA strides/base must be 16-byte aligned");
  const cuuint64_t dims[3] = {inner, rows, batch};
  const cuuint64_t strides[2] = {row_stride_elems * (uint64_t)esize, batch_stride_elems * (uint64_t)esize};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {box_inner, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GVIT_REQUIRE(r == CUDA_SUCCESS, GVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GVIT_OK;
}


int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows) {
  return make_tmap_3d(out, base, 2, inner, rows, batch, row_stride_elems, batch_stride_elems, box_rows);
}

int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows) {
  return make_tmap_3d(out, base, 4, inner, rows, batch, row_stride_elems, batch_stride_elems, box_rows);
}

}  // namespace gvit

// tma_probe.cu - diagnostic only (not on the product path): streams a (B, rows, 64*slabs) bf16 tensor through shared
// memory with cp.async.bulk.tensor, S stages in flight per CTA, and nothing else.  Used to measure what TMA streaming
// bandwidth the tile shapes of the real kernels can reach (gvit_probe_tma, called from tools/tma_probe.py).
#include "kernels.cuh"
#include "tc.cuh"

namespace gvit {
namespace {
using namespace tc;

__global__ void __launch_bounds__(64, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int stages, int stage_bytes,
                                                          int box_rows, int slabs, int row_tiles, int images,
                                                          unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int per_image = slabs * row_tiles;
  const long long total = (long long)images * per_image;
  if (threadIdx.x == 0) {                       // producer
    int it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % stages;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
      mbar_expect_tx(&full[s], (uint32_t)(box_rows * 128));
      const int b = (int)(t / per_image), r = (int)(t % per_image);
      tma_load_3d(sm + (size_t)s * stage_bytes, &tm, (r % slabs) * 64, (r / slabs) * box_rows, b, &full[s]);
    }
  } else if (threadIdx.x == 32) {               // consumer: touch one word, release the stage
    unsigned long long acc = 0;
    int it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % stages;
      mbar_wait(&full[s], (it / stages) & 1);
      acc += *reinterpret_cast<volatile uint32_t*>(sm + (size_t)s * stage_bytes);
      mbar_arrive(&empty[s]);
    }
    if (acc == 0x1234567ull) *sink = acc;
  }
}
}  // namespace
}  // namespace gvit

extern "C" __attribute__((visibility("default"))) int gvit_probe_tma(const void* base, int B, int rows, int cols, int box_rows,
                                                                      int stages, int ctas_per_sm, void* sink, void* stream) {
  using namespace gvit;
  CUtensorMap tm;
  int rc = make_tmap_bf16_3d(&tm, base, cols, rows, B, cols, (uint64_t)rows * cols, box_rows);
  if (rc != GVIT_OK) return rc;
  const int stage_bytes = ((box_rows * 128 + 1023) / 1024) * 1024;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 2 * stages * 8;
  GVIT_CHECK_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int row_tiles = (rows + box_rows - 1) / box_rows;
  tma_probe_kernel<<<num_sms() * ctas_per_sm, 64, smem, static_cast<cudaStream_t>(stream)>>>(
      tm, stages, stage_bytes, box_rows, cols / 64, row_tiles, B, static_cast<unsigned long long*>(sink));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

// launch_overhead.cu - what an event pair around ONE launch of a (nearly) empty persistent-style kernel measures on B200:
// the floor under every libgvit kernel time in tools/kernel_bench.py / bench.py's kernel table.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/launch_overhead tools/launch_overhead.cu && tools/bin/launch_overhead
// Variants: shared-memory size (0 / 227 KB), cluster of 2, TMEM allocation (512 columns), grid 148 x 320 threads.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

__global__ void spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }

template <bool TMEM>
__global__ void __launch_bounds__(320, 1) probe(unsigned long long* span, int smem_bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned int tmem_slot;
  unsigned long long t;
  if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); span[2 * blockIdx.x] = t; }
  if (TMEM) {
    if (threadIdx.x < 32) {
      unsigned int a = (unsigned int)__cvta_generic_to_shared(&tmem_slot);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
  }
  if (smem_bytes > 0 && smem[threadIdx.x] == 123 && span == nullptr) printf("x");
  if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); span[2 * blockIdx.x + 1] = t; }
}

template <bool TMEM>
static void run(const char* name, int smem, int cluster, unsigned long long* span) {
  CK(cudaFuncSetAttribute(probe<TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  const int iters = 20;
  std::vector<cudaEvent_t> a(iters), b(iters);
  for (int i = 0; i < iters; ++i) { CK(cudaEventCreate(&a[i])); CK(cudaEventCreate(&b[i])); }
  for (int i = 0; i < 3; ++i) CK(cudaLaunchKernelEx(&cfg, probe<TMEM>, span, smem));
  CK(cudaDeviceSynchronize());
  spin<<<1, 1>>>(3000000);
  for (int i = 0; i < iters; ++i) { CK(cudaEventRecord(a[i])); CK(cudaLaunchKernelEx(&cfg, probe<TMEM>, span, smem)); CK(cudaEventRecord(b[i])); }
  CK(cudaDeviceSynchronize());
  std::vector<float> ms(iters);
  for (int i = 0; i < iters; ++i) CK(cudaEventElapsedTime(&ms[i], a[i], b[i]));
  std::sort(ms.begin(), ms.end());
  std::vector<unsigned long long> h(2 * 148);
  CK(cudaMemcpy(h.data(), span, h.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long b0 = ~0ull, e1 = 0;
  for (int i = 0; i < 148; ++i) { b0 = std::min(b0, h[2 * i]); e1 = std::max(e1, h[2 * i + 1]); }
  // back to back WITHOUT events in between: the per-launch cost a stream / graph of such kernels pays
  cudaEvent_t s, e; CK(cudaEventCreate(&s)); CK(cudaEventCreate(&e));
  spin<<<1, 1>>>(3000000);
  CK(cudaEventRecord(s));
  for (int i = 0; i < 50; ++i) CK(cudaLaunchKernelEx(&cfg, probe<TMEM>, span, smem));
  CK(cudaEventRecord(e));
  CK(cudaDeviceSynchronize());
  float tot; CK(cudaEventElapsedTime(&tot, s, e));
  printf("%-44s event pair: median %.2f us (min %.2f); CTA lifetimes span %.2f us; 50 back to back: %.2f us each\n", name,
         ms[iters / 2] * 1e3, ms[0] * 1e3, (e1 - b0) * 1e-3, tot * 1e3 / 50);
}

int main() {
  unsigned long long* span; CK(cudaMalloc(&span, 2 * 148 * 8));
  run<false>("148 x 320, no shared memory", 0, 1, span);
  run<false>("148 x 320, 227 KB shared memory", 227 * 1024, 1, span);
  run<false>("148 x 320, 227 KB, cluster 2", 227 * 1024, 2, span);
  run<true>("148 x 320, 226 KB, TMEM 512", 226 * 1024, 1, span);
  run<true>("148 x 320, 226 KB, cluster 2, TMEM 512", 226 * 1024, 2, span);
  return 0;
}

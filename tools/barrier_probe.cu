// Grid-barrier variants for the persistent step kernel (132 CTAs x 512 threads, clusters of 4): cycles per barrier.
#include <cstdio>
#include <cuda_runtime.h>
#include "../jamie_b200/csrc/ptx.cuh"
using namespace jb;
__device__ __forceinline__ void bar_flag(unsigned int* counter, unsigned int* flag, unsigned int target, unsigned int epoch) {
  fence_proxy_async_global();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    if (old == target - 1) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    } else {
      unsigned int v;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (static_cast<int>(v - epoch) < 0);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_hier(unsigned int* counter, unsigned int target) {
  fence_proxy_async_global();
  cluster_sync_all();
  if (cluster_ctarank() == 0 && threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (static_cast<int>(v - target) < 0);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  cluster_sync_all();
}
__global__ void __launch_bounds__(512, 1) probe(unsigned int* bar, long long* clk, int iters, int variant, float* sink) {
  unsigned int target = 0;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // a little work with skew, like a phase
    for (int k = 0; k < (blockIdx.x & 3) * 50; ++k) acc += __sinf(acc + k);
    if (variant == 0) { target += gridDim.x; grid_barrier(bar, target); }
    else if (variant == 1) { target += gridDim.x; bar_flag(bar, bar + 32, target, it + 1); }
    else { target += gridDim.x / 4; bar_hier(bar, target); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { sink[blockIdx.x] = acc; if (blockIdx.x == 0) clk[0] = t1 - t0; }
}
int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  unsigned int* bar; long long* clk; float* sink;
  cudaMalloc(&bar, 1024); cudaMalloc(&clk, 16); cudaMalloc(&sink, 4096);
  for (int variant = 0; variant < 3; ++variant)
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(bar, 0, 1024);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(132); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = 2;
      int iters = 500;
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, bar, clk, iters, variant, sink);
      cudaError_t e2 = cudaDeviceSynchronize();
      long long h = 0;
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      printf("variant %d: %s %s %.0f clk/iter (incl. ~150 x 3 sinf of skewed work)\n", variant, cudaGetErrorString(e), cudaGetErrorString(e2), h / 500.0);
    }
  return 0;
}

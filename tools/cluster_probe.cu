// Probe: how many 4-CTA clusters of a 1-CTA-per-SM kernel are co-resident on this GPU, and does a cooperative + cluster
// launch work; measures a cluster barrier + DSMEM exchange and the grid barrier under a cluster launch.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../jamie_b200/csrc/ptx.cuh"
using namespace jb;
__global__ void __launch_bounds__(512, 1) probe(unsigned int* bar, float* out, long long* clk, int iters) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank(), nr = cluster_nctarank();
  unsigned int target = 0;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    sm[tid] = static_cast<float>(blockIdx.x + it);
    cluster_sync_all();
    float s = 0.f;
    for (uint32_t r = 0; r < nr; ++r) {
      uint32_t a = mapa_shared(smem_u32(sm + tid), r);
      float v;
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a));
      s += v;
    }
    acc += s;
    cluster_sync_all();
  }
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) {
    target += gridDim.x;
    grid_barrier(bar, target);
  }
  long long t2 = clock64();
  if (tid == 0) { out[blockIdx.x] = acc; if (blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; } }
}
int main() {
  int smem = 225 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs = 1; cs <= 8; cs *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, probe, &cfg);
    printf("cluster size %d: max active clusters %d (%s) -> %d CTAs\n", cs, nc, cudaGetErrorString(e), nc * cs);
  }
  unsigned int* bar; float* out; long long* clk;
  cudaMalloc(&bar, 4); cudaMalloc(&out, 4 * 256); cudaMalloc(&clk, 16);
  for (int grid : {128, 132, 136, 140, 144, 148}) {
    for (int coop = 0; coop < 2; ++coop) {
      cudaMemset(bar, 0, 4);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
      int iters = 200;
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, bar, out, clk, iters);
      cudaError_t e2 = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
      printf("grid %d coop %d: launch %s, sync %s; cluster exchange %.0f clk/iter, grid barrier %.0f clk/iter\n", grid, coop,
             cudaGetErrorString(e), cudaGetErrorString(e2), h[0] / 200.0, h[1] / 200.0);
      if (e2 != cudaSuccess) return 1;
    }
  }
  return 0;
}

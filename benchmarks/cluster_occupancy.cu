// How many clusters of 2..8 (and 16, non-portable) CTAs with ~225 KB of shared memory each fit on the GPU at once?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/cluster_occupancy benchmarks/cluster_occupancy.cu
#include <cuda_runtime.h>

#include <cstdio>

__global__ void __launch_bounds__(224, 1) k(int* p) {
  extern __shared__ char smem[];
  if (p) p[threadIdx.x] = smem[threadIdx.x];
}

int main() {
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 12, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(224);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %3d (%3d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}

// How many clusters of each size are co-resident on this GPU with one ~208 KB CTA per SM?
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs %d\n", pr.name, pr.multiProcessorCount);
  for (int smem : {100 * 1024, 208 * 1024, 227 * 1024})
    for (int cs : {1, 2, 4, 6, 8, 10, 12, 14, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs, 1, 1); cfg.blockDim = dim3(320, 1, 1); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
      cudaGetLastError();
    }
  return 0;
}

// How many clusters of 1-CTA-per-SM kernels fit on the device?  (tools/micro)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / c * c); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %2d: max active clusters %d (%d CTAs) %s\n", c, n, n * c, cudaGetErrorString(e));
  }
  return 0;
}

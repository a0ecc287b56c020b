// How many 8-CTA clusters of a 256-thread, 255-register kernel can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1) k8(int* p) { if (p) p[0] = 1; }
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(256, 1) k4(int* p) { if (p) p[0] = 1; }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k2(int* p) { if (p) p[0] = 1; }
template <typename K> void q(K kern, int cs, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg);
  printf("cluster %d, dyn smem %zu KB: max active clusters %d (%d CTAs) %s\n", cs, smem >> 10, n, n * cs, cudaGetErrorString(e));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  q(k8, 8, 0); q(k8, 8, 120 << 10); q(k4, 4, 120 << 10); q(k2, 2, 120 << 10);
  return 0;
}

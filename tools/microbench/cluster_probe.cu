// Which cluster launches does this box accept?  (diagnosis of "cluster misconfiguration" for the CTA-pair GEMM)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) atomicAdd(out + r, 1);
}
static void tryit(dim3 grid, dim3 cl, int threads, size_t smem, bool pdl, int* d) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute a[2]; int n = 0;
  if (pdl) { a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  a[n].id = cudaLaunchAttributeClusterDimension; a[n].val.clusterDim.x = cl.x; a[n].val.clusterDim.y = cl.y; a[n].val.clusterDim.z = cl.z; ++n;
  cfg.attrs = a; cfg.numAttrs = n;
  int nc = -1;
  cudaError_t eo = cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
  cudaMemset(d, 0, 64);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("grid (%d,%d,%d) cluster (%d,%d,%d) threads %d smem %zu pdl %d: occupancy query %s -> %d clusters; launch %s / %s; ranks %d %d\n",
         grid.x, grid.y, grid.z, cl.x, cl.y, cl.z, threads, smem, (int)pdl, cudaGetErrorName(eo), nc, cudaGetErrorName(e), cudaGetErrorName(e2), h[0], h[1]);
  cudaGetLastError();
}
int main() {
  int* d; cudaMalloc(&d, 64);
  tryit(dim3(4, 1, 1), dim3(2, 1, 1), 128, 0, false, d);
  tryit(dim3(3, 2, 1), dim3(1, 2, 1), 128, 0, false, d);
  tryit(dim3(3, 2, 1), dim3(1, 2, 1), 320, 0, false, d);
  tryit(dim3(3, 2, 1), dim3(1, 2, 1), 320, 99584, false, d);
  tryit(dim3(3, 2, 1), dim3(1, 2, 1), 320, 197888, false, d);
  tryit(dim3(3, 2, 1), dim3(1, 2, 1), 320, 197888, true, d);
  tryit(dim3(8, 32, 1), dim3(1, 2, 1), 320, 99584, true, d);
  tryit(dim3(8, 32, 4), dim3(1, 2, 1), 320, 99584, true, d);
  return 0;
}

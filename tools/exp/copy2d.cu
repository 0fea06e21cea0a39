#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
int main(){
  const size_t n = 1011240;
  char *h; cudaMallocHost(&h, n*80); memset(h, 1, n*80);
  char *d80, *d32; cudaMalloc(&d80, n*80); cudaMalloc(&d32, n*32);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int rep=0; rep<3; ++rep){
    cudaEventRecord(a,s); cudaMemcpyAsync(d80,h,n*80,cudaMemcpyHostToDevice,s); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D contiguous 80B x n: %.3f ms (%.1f GB/s)\n", ms, n*80/ms/1e6);
    cudaEventRecord(a,s); cudaMemcpy2DAsync(d32,32,h,80,32,n,cudaMemcpyHostToDevice,s); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 2D width 32 of pitch 80: %.3f ms (%.1f GB/s payload)\n", ms, n*32/ms/1e6);
    cudaEventRecord(a,s); cudaMemcpyAsync(h,d80,n*80,cudaMemcpyDeviceToHost,s); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("D2H contiguous 80B x n: %.3f ms (%.1f GB/s)\n", ms, n*80/ms/1e6);
  }
  return 0;
}

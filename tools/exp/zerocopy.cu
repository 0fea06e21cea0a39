// Experiment: AoS->SoA straight out of mapped pinned host memory (zero copy), reading only the 48 needed bytes of
// each 80-byte record, versus cudaMemcpyAsync of the whole mirror followed by a device-side conversion.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
__global__ void k_zc(const float4* __restrict__ aos, float4* __restrict__ pos, float4* __restrict__ vel, int n, int full){
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i>=n) return;
  const float4* r = aos + 5*(size_t)i;
  float4 p = r[0], v = r[1], t = r[4];
  if (full) { float4 a = r[2], g = r[3]; p.x += a.x*0.f + g.x*0.f; }
  p.w = t.z; pos[i]=p; vel[i]=v;
}
int main(){
  const size_t n = 1011240;
  char *h; cudaHostAlloc(&h, n*80, cudaHostAllocMapped); memset(h, 1, n*80);
  float4 *dh; cudaHostGetDevicePointer(&dh, h, 0);
  char *d80; cudaMalloc(&d80, n*80); float4 *pos,*vel; cudaMalloc(&pos,n*16); cudaMalloc(&vel,n*16);
  cudaStream_t s; cudaStreamCreate(&s); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  for (int rep=0; rep<3; ++rep){
    cudaEventRecord(a,s); cudaMemcpyAsync(d80,h,n*80,cudaMemcpyHostToDevice,s); k_zc<<<(n+255)/256,256,0,s>>>((float4*)d80,pos,vel,n,0); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("memcpy 80B + convert: %.3f ms\n", ms);
    cudaEventRecord(a,s); k_zc<<<(n+255)/256,256,0,s>>>(dh,pos,vel,n,0); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("zero-copy 48 of 80 B: %.3f ms\n", ms);
    cudaEventRecord(a,s); k_zc<<<(n+255)/256,256,0,s>>>(dh,pos,vel,n,1); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("zero-copy 80 of 80 B: %.3f ms\n", ms);
  }
  return 0;
}

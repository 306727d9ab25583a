// Dependent-chain latencies of the fp64 operations on the LM pivot chain (one warp, clock64), and fp64 FMA throughput.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__global__ void lat(double* out, long long* cyc, double seed) {
  double v = seed + threadIdx.x * 1e-3;
  long long t0, t1;
  int k = 0;
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 256; ++i) v = fma(v, 1.0000001, 1e-9);
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) v = rsqrt(v) + 1.5;
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) v = __shfl_sync(0xffffffffu, v, (i * 7) & 31) ;
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) v = 1.0 / v + 1.5;
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) v = sqrt(v) + 1.5;
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) { float f = rsqrtf((float)v); v = (double)f + 1.5; }
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  asm volatile("" : "+d"(v)); t0 = clk(); asm volatile("" : "+d"(v));
#pragma unroll
  for (int i = 0; i < 64; ++i) { double x = (double)rsqrtf((float)v); x = x * fma(-0.5 * v * x, x, 1.5); x = x * fma(-0.5 * v * x, x, 1.5); v = x + 1.5; }
  asm volatile("" : "+d"(v)); t1 = clk(); asm volatile("" : "+d"(v)); cyc[k++] = (t1 - t0);
  out[threadIdx.x] = v;
}
__global__ void thr(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
    a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 8 * 1024 * 8); cudaMallocManaged(&cyc, 64 * 8);
  lat<<<1, 32>>>(out, cyc, 1.7); cudaDeviceSynchronize();
  lat<<<1, 32>>>(out, cyc, 1.7); cudaDeviceSynchronize();
  printf("dfma chain: %.1f cyc/op\nrsqrt(double)+add: %.1f\nshfl double: %.1f\n1/x+add: %.1f\nsqrt+add: %.1f\nrsqrtf via float + cvt + add: %.1f\nrsqrtf seed + 2 newton + add: %.1f\n",
         cyc[0] / 256.0, cyc[1] / 64.0, cyc[2] / 64.0, cyc[3] / 64.0, cyc[4] / 64.0, cyc[5] / 64.0, cyc[6] / 64.0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  thr<<<148 * 8, 1024>>>(out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); thr<<<148 * 8, 1024>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("fp64 FMA throughput: %.2f TFLOP/s\n", 2.0 * 8 * iters * 148.0 * 8 * 1024 / (ms * 1e-3) / 1e12);
  return 0;
}

// Latency of the 8x8 diagonal-block Cholesky factor on one warp (the LM pivot chain), variants.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CB = 8;
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }

// variant 0: as in wc_solve.cu (lanes = rows, shuffles)
__device__ __forceinline__ void factor_v0(double* A, int LD, int k0, double* rinv_out, int* s_fail) {
  const int lane = threadIdx.x & 31;
  double    a[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) a[c] = (lane < CB && c <= lane) ? A[(k0 + lane) * LD + k0 + c] : (c == lane ? 1.0 : 0.0);
  bool   bad  = false;
  double rinv = 1.0;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const double djj = __shfl_sync(0xffffffffu, a[j], j);
    if (!(djj > 0.0) || !isfinite(djj)) bad = true;
    const double rs = rsqrt(djj);
    if (lane == j) rinv = rs;
    if (lane >= j) a[j] = (lane == j) ? djj * rs : a[j] * rs;
#pragma unroll
    for (int k = j + 1; k < CB; ++k) {
      const double lkj = __shfl_sync(0xffffffffu, a[j], k);
      if (lane >= k) a[k] -= a[j] * lkj;
    }
  }
  if (bad && lane == 0) *s_fail = 1;
  if (lane < CB) {
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c <= lane) A[(k0 + lane) * LD + k0 + c] = a[c];
    rinv_out[lane] = rinv;
  }
}
// variant 1: every lane holds the WHOLE 8x8 block in registers and factors it redundantly: no shuffles at all
__device__ __forceinline__ double fast_rsqrt(double d) {
  double x = (double)rsqrtf((float)d);
  x = x * fma(-0.5 * d * x, x, 1.5);
  x = x * fma(-0.5 * d * x, x, 1.5);
  return x;
}
template <int RS>
__device__ __forceinline__ void factor_v1(double* A, int LD, int k0, double* rinv_out, int* s_fail) {
  const int lane = threadIdx.x & 31;
  double    a[CB][CB];
#pragma unroll
  for (int r = 0; r < CB; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a[r][c] = A[(k0 + r) * LD + k0 + c];
  bool bad = false;
  double ri[CB];
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const double djj = a[j][j];
    if (!(djj > 0.0) || !isfinite(djj)) bad = true;
    const double rs = RS ? fast_rsqrt(djj) : rsqrt(djj);
    ri[j] = rs;
    a[j][j] = djj * rs;
#pragma unroll
    for (int r = j + 1; r < CB; ++r) a[r][j] *= rs;
#pragma unroll
    for (int k = j + 1; k < CB; ++k)
#pragma unroll
      for (int r = k; r < CB; ++r) a[r][k] = fma(-a[r][j], a[k][j], a[r][k]);
  }
  if (bad && lane == 0) *s_fail = 1;
  if (lane < CB) {
#pragma unroll
    for (int r = 0; r < CB; ++r)
      if (r == lane) {
#pragma unroll
        for (int c = 0; c <= r; ++c) A[(k0 + r) * LD + k0 + c] = a[r][c];
        rinv_out[r] = ri[r];
      }
  }
}
__global__ void bench(double* gA, long long* cyc, int variant) {
  __shared__ double A[16 * 18];
  __shared__ double rinv[16];
  __shared__ int fail;
  const int LD = 18;
  for (int i = threadIdx.x; i < 16 * 18; i += 32) A[i] = gA[i];
  __syncwarp();
  long long t0 = clk();
  if (variant == 0) factor_v0(A, LD, 0, rinv, &fail);
  else if (variant == 1) factor_v1<0>(A, LD, 0, rinv, &fail);
  else factor_v1<1>(A, LD, 0, rinv, &fail);
  __syncwarp();
  long long t1 = clk();
  if (threadIdx.x == 0) cyc[variant] = t1 - t0;
  for (int i = threadIdx.x; i < 16 * 18; i += 32) gA[i + (variant + 1) * 16 * 18] = A[i];
}
// interference: warp 0 factors while the other warps run (mode 1) dependent-free DFMA streams, (mode 2) LDS.128 streams
__global__ void bench2(double* gA, long long* cyc, int variant, int mode, int nwarps_busy, int skip_smsp0) {
  __shared__ double A[16 * 18];
  __shared__ double rinv[16];
  __shared__ int fail;
  __shared__ double buf[4096];
  __shared__ volatile int stop;
  const int LD = 18, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 16 * 18; i += blockDim.x) A[i] = gA[i];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i * 1e-3;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    long long t0 = clk();
    if (variant == 0) factor_v0(A, LD, 0, rinv, &fail);
    else factor_v1<0>(A, LD, 0, rinv, &fail);
    __syncwarp();
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[0] = t1 - t0, stop = 1;
  } else if (warp <= nwarps_busy && !(skip_smsp0 && (warp & 3) == 0)) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double2 acc = make_double2(0, 0);
    int it = 0;
    while (!stop && it < 100000) {
      if (mode == 1) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9); }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) { double2 d = reinterpret_cast<double2*>(buf)[(threadIdx.x * 7 + u * 33 + it) & 2047]; acc.x += d.x; acc.y += d.y; }
      }
      ++it;
    }
    gA[2000 + threadIdx.x] = a0 + a1 + a2 + a3 + acc.x + acc.y;
  }
}
int main() {
  double h[16 * 18];
  // SPD 8x8: M M^T + 8 I
  double M[8][8];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) M[i][j] = sin(1.0 + i * 3.1 + j * 1.7);
  for (int i = 0; i < 16 * 18; ++i) h[i] = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double s = i == j ? 8.0 : 0.0; for (int k = 0; k < 8; ++k) s += M[i][k] * M[j][k]; h[i * 18 + j] = s; }
  double* d; long long* cyc;
  cudaMalloc(&d, 4 * 16 * 18 * 8); cudaMallocManaged(&cyc, 64);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 3; ++v) { bench<<<1, 32>>>(d, cyc, v); cudaDeviceSynchronize(); }
  double r[4][16 * 18];
  cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j <= i; ++j) { e1 = fmax(e1, fabs(r[1][i * 18 + j] - r[2][i * 18 + j])); e2 = fmax(e2, fabs(r[1][i * 18 + j] - r[3][i * 18 + j])); }
  printf("factor 8x8 cycles: v0 shuffles %lld | v1 all-in-registers %lld | v1 + fast rsqrt %lld | max diff v1 %.3g v2 %.3g (L00 %.6f)\n", cyc[0], cyc[1], cyc[2], e1, e2, r[1][0]);
  for (int variant = 0; variant < 2; ++variant)
    for (int mode = 1; mode <= 2; ++mode)
      for (int nb : {0, 3, 15})
        for (int skip : {0, 1}) {
          cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
          bench2<<<1, 512>>>(d, cyc, variant, mode, nb, skip); cudaDeviceSynchronize();
          cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
          bench2<<<1, 512>>>(d, cyc, variant, mode, nb, skip); cudaDeviceSynchronize();
          printf("variant %d interference %s busy warps %2d skip-smsp0 %d : factor %lld cycles\n", variant, mode == 1 ? "DFMA" : "LDS ", nb, skip, cyc[0]);
        }
  return 0;
}

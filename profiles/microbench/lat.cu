// Dependent-chain latencies of the primitives the KKT pivot warp is built from (one warp, clock64 around N repeats).
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void lat(double* out, long long* cyc, double seed, int iseed) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  double a = seed + lane, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = __fma_rn(a, b, c);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = __dmul_rn(a, b);
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // 64-bit shuffle chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (lane + iseed) & 31);
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // REDUX chain
  unsigned u = (unsigned)lane * 77u + iseed;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) u = __reduce_max_sync(0xffffffffu, u ^ lane) + lane;
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // LDS chain (pointer chase in shared)
  int idx = lane;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) idx = ((int)sm[idx & 2047] + idx + 1) & 2047;
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // F2F double->float->double chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = (double)((float)a) + 0.0;
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // rcp.approx.f64 + 2 Newton steps chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = __fma_rn(-a, r, 1.0); r = __fma_rn(r, e, r); e = __fma_rn(-a, r, 1.0); a = __fma_rn(r, e, r) + 1.5;
  }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // full division chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = 1.0 / a + 1.5;
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // 32-bit shuffle chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) u = __shfl_sync(0xffffffffu, u, (lane + iseed) & 31);
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) a = __dadd_rn(a, c);
  t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
  // STS + LDS round trip (same thread)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { sm[lane + 64] = a; __syncwarp(); a = sm[((lane + 1) & 31) + 64] + 1.0; __syncwarp(); }
  t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
  out[lane] = a + u + idx;
}
// bar.sync of 288 threads, repeated
__global__ void barlat(long long* cyc, double* out) {
  __shared__ double sm[4096];
  long long t0 = clock64();
  double a = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < N; ++i) { __syncthreads(); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // with 8 STS per thread before each barrier
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[(threadIdx.x + k * 288) & 4095] = a + k;
    __syncthreads();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<double2*>(sm)[(threadIdx.x + k * 288) & 2047] = make_double2(a, a + k);
    __syncthreads();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  out[threadIdx.x] = sm[threadIdx.x];
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 16 * 8);
  long long h[16];
  const char* names[] = {"DFMA", "DMUL", "SHFL64", "REDUX.MAX", "LDS chase(+I2F etc)", "F2F x2 + DADD", "rcp.approx+2NR (+DADD)", "1/x (+DADD)", "SHFL32", "DADD", "STS->LDS (+DADD, 2 syncwarp)"};
  for (int rep = 0; rep < 2; ++rep) { lat<<<1, 64>>>(out, cyc, 1.25, 3); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 11; ++i) printf("%-32s %.1f cycles\n", names[i], (double)h[i] / N);
  for (int rep = 0; rep < 2; ++rep) { barlat<<<1, 288>>>(cyc, out); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("bar.sync 288 thr: %.1f ; with 8 STS.64: %.1f ; with 4 STS.128: %.1f cycles\n", (double)h[0] / N, (double)h[1] / N, (double)h[2] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

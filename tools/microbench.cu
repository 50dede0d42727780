// Microbenchmarks used to size the rollout kernel (not part of the product library):
// dependent-issue latency of DFMA / DADD / DMUL / MUFU.RCP64H on one warp, and FP64 throughput
// versus warps per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-9;
  const double a = 0.999999999, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dadd(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-9;
  const double b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = x + b;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rcp(double* out, long long* cyc, int iters) {
  double x = 1.5 + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ILP k independent chains per warp, W warps per block on one SM: cycles per DFMA warp-instruction
template <int K>
__global__ void tput(double* out, long long* cyc, int iters) {
  double x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
  const double a = 0.999999999, b = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = fma(x[k], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// DFMA issue cost versus the number of active lanes of the warp (does a half-empty warp issue in
// one pass of the 16-lane FP64 unit?).  mask = active lanes; 4 independent chains, 1 warp.
__global__ void tput_lanes(double* out, long long* cyc, int iters, unsigned mask) {
  double x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
  const double a = 0.999999999, b = 1e-9;
  long long t0 = clock64();
  if ((mask >> (threadIdx.x & 31)) & 1u) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = fma(x[k], a, b);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x[0] + x[1] + x[2] + x[3];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  long long h;
  const int it = 4096;
  lat_dfma<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency   %.2f cycles\n", (double)h / (it * 16));
  lat_dadd<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DADD dependent latency   %.2f cycles\n", (double)h / (it * 16));
  lat_rcp<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent    %.2f cycles\n", (double)h / (it * 16));
  for (int warps = 1; warps <= 16; warps *= 2) {
    tput<1><<<1, 32 * warps>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double c1 = (double)h / (it * 8 * 1);
    tput<2><<<1, 32 * warps>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double c2 = (double)h / (it * 8 * 2);
    tput<4><<<1, 32 * warps>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double c4 = (double)h / (it * 8 * 4);
    printf("warps/SM %2d (per SMSP %.2f): cycles per DFMA per warp: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f\n",
           warps, warps / 4.0, c1, c2, c4);
  }
  const unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x55555555u, 0x000000ffu, 0x00ff00ffu, 0x0000000fu, 0x1u};
  for (unsigned m : masks) {
    tput_lanes<<<1, 32>>>(out, cyc, it, m); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("active-lane mask %08x: %.2f cycles per DFMA warp-instruction (1 warp, ILP4)\n", m, (double)h / (it * 8 * 4));
  }
  return 0;
}

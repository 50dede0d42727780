// Does the FP64 matrix instruction (DMMA, mma.sync m8n8k4 / m16n8k8 f64) run beside the FP64 pipe on B200, or through it?
// Cycles per warp-instruction per SM sub-partition, alone and interleaved with independent DFMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench3 microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: 8 DMMA m8n8k4 per iteration; 1: 8 DMMA + NF DFMA; 2: NF DFMA only; 3: 4 DMMA m16n8k8; 4: 4 DMMA m16n8k8 + NF DFMA
template <int MODE, int NF>
__global__ void k(double* out, long long* cyc, int iters) {
  double c[8][2], c4[4][4], x[8], a4[4], b2[2];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; x[i] = 1.0 + i + threadIdx.x * 1e-9; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { a4[i] = a + i; c4[i][0] = c4[i][1] = c4[i][2] = c4[i][3] = i; }
  b2[0] = b; b2[1] = -b;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i], a, b);
    }
    if (MODE == 3 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma1688(c4[i], a4, b2);
    }
    if (MODE == 1 || MODE == 2 || MODE == 4) {
#pragma unroll
      for (int u = 0; u < NF / 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], 0.999999999, 1e-9);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8);
  long long h;
  const int it = 2048;
  for (int warps = 1; warps <= 16; warps *= 4) {
    const double per = warps >= 4 ? warps / 4.0 : 1.0;
    printf("-- %d warp(s) in one CTA (%.2f per SMSP); cycles per iteration per SMSP-resident warp\n", warps, warps / 4.0);
#define RUN(M, NF, what) k<M, NF><<<1, 32 * warps>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("  %-48s %8.1f\n", what, (double)h / it / per);
    RUN(0, 0, "8 x DMMA m8n8k4 (4096 flops)")
    RUN(2, 64, "64 x DFMA (4096 flops)")
    RUN(1, 64, "8 x DMMA m8n8k4 + 64 x DFMA")
    RUN(3, 0, "4 x DMMA m16n8k8 (8192 flops)")
    RUN(2, 128, "128 x DFMA (8192 flops)")
    RUN(4, 128, "4 x DMMA m16n8k8 + 128 x DFMA")
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

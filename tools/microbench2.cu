// FP64 issue cost by operand form (register vs constant operands), 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double a_in, double b_in) {
  double x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = 1.0 + threadIdx.x * 1e-9 + i;
    y[i] = 0.999999 + threadIdx.x * 1e-12 + i * 1e-10;
    z[i] = 1e-9 * (i + 1) + threadIdx.x * 1e-15;
  }
  const double a = 0.999999999, b = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = fma(x[i], a, b);            // reg, const, const
        if (MODE == 1) x[i] = fma(x[i], y[i], z[i]);      // three distinct registers
        if (MODE == 2) x[i] = fma(y[i], z[i], x[i]);      // accumulate form
        if (MODE == 3) x[i] = fma(x[i], y[i], b);         // reg, reg, const
        if (MODE == 4) x[i] = x[i] + y[i];                // DADD reg reg
        if (MODE == 5) x[i] = x[i] * y[i];                // DMUL reg reg
        if (MODE == 6) x[i] = fma(x[i], a_in, b_in);      // kernel-parameter (constant bank) operands
        if (MODE == 8) x[i] = fma(x[i], y[0], z[i]);      // one register operand shared by consecutive DFMAs
        if (MODE == 9) x[i] = fma(y[0], z[0], x[i]);      // two shared register operands
        if (MODE == 10) x[i] = fma(y[i], y[0], x[i]);     // shared multiplier, accumulate form
        if (MODE == 7) x[i] = fma(x[i], y[(i + 1) & 7], z[(i + 3) & 7]);  // shuffled register operands
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8);
  long long h;
  const int it = 4096;
  const char* names[] = {"DFMA reg,const,const", "DFMA reg,reg,reg (x=x*y+z)", "DFMA reg,reg,reg (x=y*z+x)",
                         "DFMA reg,reg,const", "DADD reg,reg", "DMUL reg,reg", "DFMA reg,param,param",
                         "DFMA reg,reg,reg (rotated operands)",
                         "DFMA x=x*Y+z (Y shared by neighbours)", "DFMA x=Y*Z+x (Y,Z shared)", "DFMA x=y*Y+x (Y shared)"};
  for (int warps = 1; warps <= 16; warps *= 4) {
    printf("-- %d warp(s) in one CTA (%.2f per SMSP)\n", warps, warps / 4.0);
#define RUN(M) k<M><<<1, 32 * warps>>>(out, cyc, it, 0.999999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-40s %.2f cycles per warp-instruction per SMSP\n", names[M], (double)h / (it * 32) / (warps >= 4 ? warps / 4.0 : 1.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
  }
  return 0;
}

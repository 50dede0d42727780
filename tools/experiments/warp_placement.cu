// Where do one-warp CTAs land?  Each CTA records (%smid, %warpid) while all of them are resident (they spin on a
// clock), then the host prints how many SMs hold two warps in the same sub-partition (warpid % 4).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_placement warp_placement.cu ; ./warp_placement 512 32
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(int* smid, int* warpid, long long spin, int regs_hog) {
  __shared__ double pad[512];  // 4 KB like the lane-split kernel
  unsigned s, w;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
  pad[threadIdx.x] = s;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if ((threadIdx.x & 31) == 0) {
    const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    smid[i] = (int)s; warpid[i] = (int)w + (int)(pad[threadIdx.x] * 0 * regs_hog);
  }
}

int main(int argc, char** argv) {
  const int warps = argc > 1 ? atoi(argv[1]) : 512, block = argc > 2 ? atoi(argv[2]) : 32;
  const int wpb = block / 32, grid = warps / wpb;
  int *dsm, *dw;
  cudaMalloc(&dsm, warps * 4); cudaMalloc(&dw, warps * 4);
  k<<<grid, block>>>(dsm, dw, 2000000, 0);
  std::vector<int> sm(warps), w(warps);
  cudaMemcpy(sm.data(), dsm, warps * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(w.data(), dw, warps * 4, cudaMemcpyDeviceToHost);
  int cnt[256][4] = {}, per_sm[256] = {};
  for (int i = 0; i < warps; ++i) { cnt[sm[i]][w[i] & 3]++; per_sm[sm[i]]++; }
  int hist_sm[16] = {}, doubled = 0, sms_doubled = 0, used = 0;
  for (int s = 0; s < 256; ++s) {
    if (!per_sm[s]) continue;
    ++used; hist_sm[per_sm[s] < 15 ? per_sm[s] : 15]++;
    bool d = false;
    for (int q = 0; q < 4; ++q) if (cnt[s][q] > 1) { doubled += cnt[s][q] - 1; d = true; }
    sms_doubled += d;
  }
  printf("%d warps as %d CTAs of %d threads: %d SMs used; warps per SM histogram:", warps, grid, block, used);
  for (int i = 1; i < 16; ++i) if (hist_sm[i]) printf(" %d:%d", i, hist_sm[i]);
  printf("; SMs with two warps on one sub-partition: %d (%d surplus warps)\n", sms_doubled, doubled);
  printf("first CTAs (smid/warpid):");
  for (int i = 0; i < 12 && i < warps; ++i) printf(" %d/%d", sm[i], w[i]);
  printf("\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}

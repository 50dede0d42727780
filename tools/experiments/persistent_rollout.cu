// Experiment (not part of the product): persistent fixed-action rollout whose CTAs pull
// (64-env group, 64-step chunk) work items from an atomic ticket counter, state handed through global
// memory at chunk boundaries.  Question: how much of the gap between the chunked multi-stream schedule
// (7.37e10 env-steps/s) and two plain launches in flight (7.89e10) does it close for BASELINE config[1]?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o persistent_rollout persistent_rollout.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "../../safe-exploration-with-simulator-in-rl-algorithms_b200/csrc/dynamics.cuh"

using namespace swm;
constexpr int N = 3, NO = 2 * N + 2, NA = N - 1, BLOCK = 64;

struct Args {
  Phys P;
  int H, chunk, n_groups, n_chunks;
  long long B;
  const double* actions;
  double* state;      // [B, NO] hand-over
  double* acc;        // [B, 2]  sum of Gdot
  double* returns;    // [B]
  double* final_state;
  unsigned int* ticket;
  unsigned int* done;  // [n_groups] chunks completed
  int* error;
};

__device__ __forceinline__ void run_chunk(const Args& a, long long e, bool active, int t0, int t1, bool first,
                                          bool last) {
  double gdx, gdy, th[N], thd[N], sgx, sgy, ut[NA];
  const long long el = active ? e : a.B - 1;
  if (first) {
    gdx = gdy = 0.0; sgx = sgy = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { th[i] = 1.5707963267948966; thd[i] = 0.0; }
  } else {
    const double* sp = a.state + el * NO;
    gdx = __ldcg(sp); gdy = __ldcg(sp + 1);
#pragma unroll
    for (int i = 0; i < N; ++i) { th[i] = __ldcg(sp + 2 + 2 * i); thd[i] = __ldcg(sp + 3 + 2 * i); }
    sgx = __ldcg(a.acc + 2 * el); sgy = __ldcg(a.acc + 2 * el + 1);
  }
#pragma unroll
  for (int k = 0; k < NA; ++k) ut[k] = a.actions[el * NA + k] * a.P.u_scale;
  double sn[N], cs[N];
#pragma unroll
  for (int i = 0; i < N; ++i) sincos(th[i], &sn[i], &cs[i]);
  for (int t = t0; t < t1; ++t) {
    gym_step_tracked<N>(a.P, gdx, gdy, th, thd, sn, cs, ut, (t & 63) == 63);
    sgx += gdx; sgy += gdy;
  }
  if (!active) return;
  if (last) {
    a.returns[e] = fma(sgx, a.P.dirx, sgy * a.P.diry);
    double* o = a.final_state + e * NO;
    o[0] = gdx; o[1] = gdy;
#pragma unroll
    for (int i = 0; i < N; ++i) { o[2 + 2 * i] = th[i]; o[3 + 2 * i] = thd[i]; }
  } else {
    double* o = a.state + e * NO;
    __stcg(o, gdx); __stcg(o + 1, gdy);
#pragma unroll
    for (int i = 0; i < N; ++i) { __stcg(o + 2 + 2 * i, th[i]); __stcg(o + 3 + 2 * i, thd[i]); }
    __stcg(a.acc + 2 * e, sgx); __stcg(a.acc + 2 * e + 1, sgy);
  }
}

__global__ void __launch_bounds__(BLOCK) persistent_kernel(const Args a) {
  __shared__ unsigned int s_item;
  const int tid = threadIdx.x;
  const unsigned int total = (unsigned int)a.n_groups * a.n_chunks;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const unsigned int item = s_item;
    __syncthreads();
    if (item >= total) return;
    const int c = item / a.n_groups, g = item % a.n_groups;  // chunk-major ticket order
    if (c > 0) {
      if (tid == 0) {
        long long spins = 0;
        while (*((volatile unsigned int*)(a.done + g)) < (unsigned int)c) {
          __nanosleep(200);
          if (++spins > (1LL << 24)) { *a.error = 1; break; }  // bounded: never hang the GPU
        }
      }
      __syncthreads();
      __threadfence();
    }
    const long long e = (long long)g * BLOCK + tid;
    run_chunk(a, e, e < a.B, c * a.chunk, min(a.H, (c + 1) * a.chunk), c == 0, c == a.n_chunks - 1);
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(a.done + g, (unsigned int)(c + 1));
  }
}

__global__ void __launch_bounds__(BLOCK) plain_kernel(const Args a) {
  const long long e = (long long)blockIdx.x * BLOCK + threadIdx.x;
  run_chunk(a, e, e < a.B, 0, a.H, true, true);
}

int main(int argc, char** argv) {
  const long long B = 65536;
  const int H = 1000;
  Phys P;
  memset(&P, 0, sizeof(P));
  const double l = 1, m = 1, k = 10, h = 1e-3, n = N;
  P.l = l; P.m = m; P.k = k; P.h = h; P.max_u = 5; P.dirx = 1; P.diry = 0; P.inv_l = 1 / l; P.kappa = k * l / m;
  P.m2kappa = -2 * k * l / m; P.u_scale = 12 / (m * l * l); P.gdd_c = l / (2 * n); P.h_gdd_c = h * l / (2 * n);
  P.inv_n = 1 / n;
  std::vector<double> hact(B * NA);
  srand(1);
  for (auto& x : hact) x = -5.0 + 10.0 * (rand() / (double)RAND_MAX);
  Args a;
  a.P = P; a.H = H; a.B = B;
  double *act, *st, *acc, *ret, *fin, *ret2, *fin2;
  unsigned int *ticket, *done; int* err;
  cudaMalloc(&act, B * NA * 8); cudaMalloc(&st, B * NO * 8); cudaMalloc(&acc, B * 16); cudaMalloc(&ret, B * 8);
  cudaMalloc(&fin, B * NO * 8); cudaMalloc(&ret2, B * 8); cudaMalloc(&fin2, B * NO * 8);
  cudaMalloc(&ticket, 4); cudaMalloc(&done, (B / BLOCK) * 4); cudaMalloc(&err, 4);
  cudaMemcpy(act, hact.data(), B * NA * 8, cudaMemcpyHostToDevice);
  cudaMemset(err, 0, 4);
  a.actions = act; a.state = st; a.acc = acc; a.ticket = ticket; a.done = done; a.error = err;
  a.n_groups = (int)(B / BLOCK);
  int occ = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, persistent_kernel, BLOCK, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // plain reference
  a.returns = ret2; a.final_state = fin2; a.chunk = H; a.n_chunks = 1;
  for (int r = 0; r < 3; ++r) plain_kernel<<<(unsigned)(B / BLOCK), BLOCK>>>(a);
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) plain_kernel<<<(unsigned)(B / BLOCK), BLOCK>>>(a);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("plain launch: %.4f ms  %.3e env-steps/s (occupancy %d CTAs/SM, %d SMs)\n", ms / 10, B * H / (ms / 10) * 1e3, occ, sms);
  a.returns = ret; a.final_state = fin;
  for (int chunk : {64, 128}) {
    for (int per_sm : {8, 7, 6, 5, 4}) {
      if (per_sm < 1) continue;
      a.chunk = chunk; a.n_chunks = (H + chunk - 1) / chunk;
      const int grid = sms * per_sm;
      float best = 1e9;
      for (int r = 0; r < 6; ++r) {
        cudaMemsetAsync(ticket, 0, 4); cudaMemsetAsync(done, 0, (B / BLOCK) * 4);
        cudaEventRecord(e0);
        persistent_kernel<<<grid, BLOCK>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 1 && ms < best) best = ms;
      }
      std::vector<double> f1(B * NO), f2(B * NO);
      cudaMemcpy(f1.data(), fin, B * NO * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(f2.data(), fin2, B * NO * 8, cudaMemcpyDeviceToHost);
      int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
      printf("persistent chunk %3d, %d CTAs/SM (grid %4d): %.4f ms  %.3e env-steps/s  states bit-identical: %s  err=%d  cuda=%s\n",
             chunk, per_sm, grid, best, B * H / best * 1e3, memcmp(f1.data(), f2.data(), B * NO * 8) == 0 ? "yes" : "NO",
             herr, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

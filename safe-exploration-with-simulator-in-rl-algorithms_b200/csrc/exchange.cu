// Per-iteration record exchange of a sharded ARS iteration as ONE kernel launch per rank:
//   pack      this rank's record = [returns (2 N_local, means over R rollouts) | mask (N_local, optional) |
//             count, mean[F], M2[F]] from the rollout kernel's outputs (what engine.py used to do with
//             swm_reduce_returns + swm_stats_finalize + tensor copies),
//   exchange  store it into slot `rank` of EVERY rank's gather buffer over NVLink (peer pointers obtained
//             through CUDA IPC), publish an epoch flag on every rank, wait for all ranks' flags,
//   unpack    returns_all[2N], mask_all[N], records[world, 1+2F] for the redundant ranking / update / merge.
// It replaces the NCCL all-gather on the data path: a kernel is capturable in a CUDA graph next to eager
// NCCL traffic of the caller (timing barriers), which the captured collective was not (round 1).
// With world == 1 the same kernel only packs and unpacks.
//
// Synchronisation: flags are monotone epochs (one 64-bit word per source rank, in the DESTINATION's buffer);
// data slots are double-buffered by epoch parity, so a fast rank can never overwrite a record that a slow
// rank is still reading: writing epoch e+2 needs the slow rank's flag e+1, which it only publishes after its
// stream finished consuming epoch e.  A wait that exceeds ~10 s sets a sticky status word and falls
// through (the launch returns, the host sees swm_exchange_status != 0; later exchanges no longer wait)
// instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/swimmer_ars.h"
#include "errors.cuh"

namespace swm {

constexpr int kExBlock = 1024;
constexpr int kMaxWorld = 16;
constexpr long long kWaitCycles = 20000000000LL;  // ~10 s at 1.9 GHz

struct ExchangeDev {
  unsigned long long epoch;    // exchanges completed by this rank
  unsigned long long status;   // 0 ok, else 1 + rank that was not heard from in time
};

struct Exchange {
  int world, rank;
  size_t rec_doubles;          // capacity of one record
  size_t slot_bytes;           // record size rounded up to 128 B
  char* local;                 // cudaMalloc: [2][world][slot_bytes] data, then flags[world] (u64), then ExchangeDev
  size_t bytes;
  char* peers[kMaxWorld];      // device pointers to every rank's buffer (peers[rank] == local)
  bool opened[kMaxWorld];
  char** peers_dev;            // the same table in device memory
  int device;
};

struct PackArgs {
  // pack
  const double* returns_local;   // [2 N_local R]
  int n_local, R;
  const int* mask_local;         // [N_local] or NULL
  const double* stats_partial;   // [n_blocks, 2F] or NULL
  long long n_blocks;
  int F;
  double samples;
  const int* units;
  const double* pivot;
  // unpack
  double* returns_all;           // [2 N]
  int* mask_all;                 // [N] or NULL
  double* records;               // [world, 1+2F] or NULL
  double* record_out;            // optional: this rank's packed record (collective fallback)
  const double* gathered_in;     // optional: skip pack + exchange, unpack this [world, rec_len] buffer
  // exchange
  char** peers;                  // NULL when world == 1
  char* local;
  int world, rank;
  unsigned long long slot_bytes;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One CTA.  Record layout (doubles): [0, 2n) returns, [2n, 2n + m n) mask (m = mask ? 1 : 0), then count, mean[F], M2[F].
__global__ void __launch_bounds__(kExBlock) pack_exchange_kernel(const PackArgs a) {
  extern __shared__ double rec[];  // this rank's record
  const int tid = threadIdx.x;
  const int n2 = 2 * a.n_local;
  const int moff = n2, soff = n2 + (a.mask_local ? a.n_local : 0);
  const int rec_len = soff + (a.F > 0 ? 1 + 2 * a.F : 0);
  const double* gathered = rec;  // world == 1: unpack straight from shared memory
  size_t stride = 0;
  bool remote = false;           // gathered data written by other GPUs: read through L2
  if (a.gathered_in) {
    gathered = a.gathered_in;
    stride = (size_t)rec_len;
  } else {
  // ---- pack: per-policy mean return over its R rollouts (fixed order) ----
  for (int g = tid; g < n2; g += kExBlock) {
    double acc = 0.0;
    for (int r = 0; r < a.R; ++r) acc += a.returns_local[(long long)g * a.R + r];
    acc = a.R > 1 ? acc / a.R : acc;
    // a screened-out direction has no real-world return, whether its rollouts were skipped (the rollout kernel
    // already wrote NaN) or ran speculatively beside the simulator rollouts that screen them (engine.py)
    if (a.mask_local && a.mask_local[g >> 1] == 0) acc = __longlong_as_double(0x7ff8000000000000LL);
    rec[g] = acc;
  }
  if (a.mask_local)
    for (int k = tid; k < a.n_local; k += kExBlock) rec[moff + k] = a.mask_local[k] != 0 ? 1.0 : 0.0;
  if (a.F > 0) {
    // shifted sums -> (count, mean, M2): one warp per column, lanes stride over the per-block rows, then a
    // fixed-order butterfly (same arithmetic on every rank count: only this rank's rows are involved)
    double samples = a.samples;
    if (a.units) samples *= (double)(*a.units);
    const int warp = tid >> 5, lane = tid & 31, twoF = 2 * a.F;
    for (int c = warp; c < twoF; c += kExBlock / 32) {
      double acc = 0.0;
      if (a.stats_partial)
        for (long long r = lane; r < a.n_blocks; r += 32) acc += a.stats_partial[r * twoF + c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) rec[soff + 1 + c] = acc;
    }
    __syncthreads();
    if (tid < a.F) {
      // mean = pivot + S1/n,  M2 = S2 - S1^2/n; every thread touches only its own two entries
      const double s1 = rec[soff + 1 + tid], s2 = rec[soff + 1 + a.F + tid];
      const bool any = samples > 0.0 && a.stats_partial;
      rec[soff + 1 + tid] = any ? a.pivot[tid] + s1 / samples : 0.0;
      rec[soff + 1 + a.F + tid] = any ? fmax(s2 - s1 * (s1 / samples), 0.0) : 0.0;
    }
    if (tid == 0) rec[soff] = a.stats_partial ? samples : 0.0;
  }
  __syncthreads();
  if (a.record_out)
    for (int i = tid; i < rec_len; i += kExBlock) a.record_out[i] = rec[i];

  if (a.world > 1 && a.peers) {
    ExchangeDev* dev = reinterpret_cast<ExchangeDev*>(a.local + 2ull * a.world * a.slot_bytes + 8ull * a.world);
    const unsigned long long epoch = dev->epoch + 1;
    const unsigned long long par = epoch & 1ull;
    // ---- peer stores: slot [par][rank] of every rank's buffer ----
    for (int p = 0; p < a.world; ++p) {
      double* dst = reinterpret_cast<double*>(a.peers[p] + (par * a.world + a.rank) * a.slot_bytes);
      for (int i = tid; i < rec_len; i += kExBlock) dst[i] = rec[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) {
      unsigned long long* flag =
          reinterpret_cast<unsigned long long*>(a.peers[tid] + 2ull * a.world * a.slot_bytes) + a.rank;
      st_release_sys(flag, epoch);
    }
    // ---- wait until every rank's record of this epoch has landed here ----
    if (tid < a.world) {
      const unsigned long long* flag =
          reinterpret_cast<const unsigned long long*>(a.local + 2ull * a.world * a.slot_bytes) + tid;
      const long long t0 = clock64();
      const bool dead = dev->status != 0ull;  // a peer already timed out once: do not wait for it again
      while (ld_acquire_sys(flag) < epoch) {
        if (dead || clock64() - t0 > kWaitCycles) {
          atomicCAS(&dev->status, 0ull, 1ull + (unsigned long long)tid);
          break;
        }
        __nanosleep(100);
      }
    }
    __syncthreads();
    __threadfence_system();
    gathered = reinterpret_cast<const double*>(a.local + par * a.world * a.slot_bytes);
    stride = a.slot_bytes / sizeof(double);
    remote = true;
    if (tid == 0) dev->epoch = epoch;
  }
  }  // !gathered_in
  if (!a.returns_all) return;  // pack only
  if (!a.gathered_in && !remote && a.world > 1) return;  // packed for a collective: nothing gathered yet
  // ---- unpack (peer-written data: read through L2, never a stale L1 line) ----
  for (int r = 0; r < a.world; ++r) {
    const double* src = gathered + r * stride;
    for (int i = tid; i < n2; i += kExBlock)
      a.returns_all[(long long)r * n2 + i] = remote ? __ldcg(src + i) : src[i];
    if (a.mask_all && a.mask_local)
      for (int k = tid; k < a.n_local; k += kExBlock)
        a.mask_all[r * a.n_local + k] = (remote ? __ldcg(src + moff + k) : src[moff + k]) != 0.0 ? 1 : 0;
    if (a.records && a.F > 0)
      for (int i = tid; i < 1 + 2 * a.F; i += kExBlock)
        a.records[r * (1 + 2 * a.F) + i] = remote ? __ldcg(src + soff + i) : src[soff + i];
  }
}

}  // namespace swm

using namespace swm;

extern "C" int64_t swm_pack_record_doubles(int n_local, int has_mask, int n_features) {
  if (n_local < 1 || n_features < 0) return 0;
  return 2 * (int64_t)n_local + (has_mask ? n_local : 0) + (n_features > 0 ? 1 + 2 * (int64_t)n_features : 0);
}

extern "C" int swm_exchange_create(int world, int rank, int64_t record_doubles, swm_exchange_t** out) {
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || record_doubles < 1)
    return SWM_ERR_BAD_ARG;
  Exchange* ex = new Exchange();
  memset(ex, 0, sizeof(*ex));
  ex->world = world;
  ex->rank = rank;
  ex->rec_doubles = (size_t)record_doubles;
  ex->slot_bytes = (((size_t)record_doubles * sizeof(double)) + 127) / 128 * 128;
  ex->bytes = 2 * (size_t)world * ex->slot_bytes + 8 * (size_t)world + sizeof(ExchangeDev);
  if (cudaGetDevice(&ex->device) != cudaSuccess) { delete ex; return SWM_ERR_NO_DEVICE; }
  // a dedicated cudaMalloc allocation: CUDA IPC exports whole allocations, and torch's caching allocator
  // may hand out slices of larger (or virtual-memory) segments
  if (cudaMalloc(&ex->local, ex->bytes) != cudaSuccess) { delete ex; return swm::check_launch(); }
  if (cudaMemset(ex->local, 0, ex->bytes) != cudaSuccess ||
      cudaMalloc(&ex->peers_dev, sizeof(char*) * kMaxWorld) != cudaSuccess) {
    cudaFree(ex->local);
    delete ex;
    return swm::check_launch();
  }
  ex->peers[rank] = ex->local;
  cudaMemcpy(ex->peers_dev, ex->peers, sizeof(char*) * kMaxWorld, cudaMemcpyHostToDevice);
  cudaDeviceSynchronize();
  *out = reinterpret_cast<swm_exchange_t*>(ex);
  return SWM_OK;
}

extern "C" int swm_exchange_ipc_handle(swm_exchange_t* h, void* handle64) {
  Exchange* ex = reinterpret_cast<Exchange*>(h);
  if (!ex || !handle64) return SWM_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == SWM_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t mh;
  if (cudaIpcGetMemHandle(&mh, ex->local) != cudaSuccess) return swm::check_launch();
  memcpy(handle64, &mh, sizeof(mh));
  return SWM_OK;
}

extern "C" int swm_exchange_open_peers(swm_exchange_t* h, const void* handles) {
  Exchange* ex = reinterpret_cast<Exchange*>(h);
  if (!ex || !handles) return SWM_ERR_BAD_ARG;
  const char* hs = static_cast<const char*>(handles);
  for (int p = 0; p < ex->world; ++p) {
    if (p == ex->rank || ex->opened[p]) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, hs + (size_t)p * SWM_IPC_HANDLE_BYTES, sizeof(mh));
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return swm::check_launch();
    ex->peers[p] = static_cast<char*>(ptr);
    ex->opened[p] = true;
  }
  if (cudaMemcpy(ex->peers_dev, ex->peers, sizeof(char*) * kMaxWorld, cudaMemcpyHostToDevice) != cudaSuccess)
    return swm::check_launch();
  return SWM_OK;
}

extern "C" int swm_exchange_status(swm_exchange_t* h, uint64_t* epoch, uint64_t* status) {
  Exchange* ex = reinterpret_cast<Exchange*>(h);
  if (!ex) return SWM_ERR_BAD_ARG;
  ExchangeDev d;
  const char* p = ex->local + 2 * (size_t)ex->world * ex->slot_bytes + 8 * (size_t)ex->world;
  if (cudaMemcpy(&d, p, sizeof(d), cudaMemcpyDeviceToHost) != cudaSuccess) return swm::check_launch();
  if (epoch) *epoch = d.epoch;
  if (status) *status = d.status;
  return SWM_OK;
}

extern "C" int swm_exchange_destroy(swm_exchange_t* h) {
  Exchange* ex = reinterpret_cast<Exchange*>(h);
  if (!ex) return SWM_OK;
  cudaDeviceSynchronize();
  for (int p = 0; p < ex->world; ++p)
    if (ex->opened[p]) cudaIpcCloseMemHandle(ex->peers[p]);
  cudaFree(ex->peers_dev);
  cudaFree(ex->local);
  delete ex;
  return SWM_OK;
}

extern "C" int swm_ars_pack_exchange(swm_exchange_t* h, const swm_pack_t* p, void* stream) {
  Exchange* ex = reinterpret_cast<Exchange*>(h);
  if (!p || p->n_local < 1 || p->rollouts_per_policy < 1 || p->n_features < 0) return SWM_ERR_BAD_ARG;
  if (!p->gathered_in && !p->returns_local) return SWM_ERR_BAD_ARG;
  if (!p->returns_all && !p->record_out) return SWM_ERR_BAD_ARG;  // nothing to produce
  if (p->gathered_in && (ex || p->gathered_world < 1 || !p->returns_all)) return SWM_ERR_BAD_ARG;
  if (p->n_features > 0 && p->stats_partial && (!p->pivot || p->n_blocks < 1)) return SWM_ERR_BAD_ARG;
  const int world = p->gathered_in ? p->gathered_world : (ex ? ex->world : (p->gathered_world > 1 ? p->gathered_world : 1));
  const size_t rec_len = 2 * (size_t)p->n_local + (p->mask_local ? (size_t)p->n_local : 0) +
                         (p->n_features > 0 ? 1 + 2 * (size_t)p->n_features : 0);
  if (ex && rec_len > ex->rec_doubles) return SWM_ERR_BAD_ARG;
  if (ex && world > 1)
    for (int q = 0; q < world; ++q)
      if (!ex->peers[q]) return SWM_ERR_BAD_ARG;  // swm_exchange_open_peers was not called
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.returns_local = p->returns_local;
  a.n_local = p->n_local;
  a.R = p->rollouts_per_policy;
  a.mask_local = p->mask_local;
  a.stats_partial = p->stats_partial;
  a.n_blocks = p->n_blocks;
  a.F = p->n_features;
  a.samples = p->samples;
  a.units = p->units;
  a.pivot = p->pivot;
  a.returns_all = p->returns_all;
  a.mask_all = p->mask_all;
  a.records = p->records;
  a.record_out = p->record_out;
  a.gathered_in = p->gathered_in;
  a.world = world;
  a.rank = ex ? ex->rank : 0;
  a.peers = (ex && world > 1) ? ex->peers_dev : nullptr;
  a.local = ex ? ex->local : nullptr;
  a.slot_bytes = ex ? ex->slot_bytes : 0;
  const size_t smem = rec_len * sizeof(double);
  if (smem > 200 * 1024) return SWM_ERR_UNSUPPORTED;
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(pack_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SWM_ERR_CUDA;
  pack_exchange_kernel<<<1, kExBlock, smem, (cudaStream_t)stream>>>(a);
  return swm::check_launch();
}

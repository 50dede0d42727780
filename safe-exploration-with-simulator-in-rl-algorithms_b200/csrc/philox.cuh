// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11).
// Replaces the global MT19937 draws of the reference (ars/ars_agent.py:137, safe_ars/ars.py:84):
// perturbations are never stored or moved, every consumer regenerates them from
// (seed, iteration, direction, element).  Addressing is specified in include/swimmer_ars.h.
#pragma once
#include <stdint.h>

namespace swm {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two words (same construction as numpy's legacy random_sample).
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// Elements 2j and 2j+1 of delta_k.  dist 0: 2u-1, dist 1: u.
__host__ __device__ __forceinline__ void philox_delta_pair(uint64_t seed, uint32_t iteration,
                                                           uint32_t direction, uint32_t stream,
                                                           uint32_t j, int dist, double& d0,
                                                           double& d1) {
  uint32_t o[4];
  philox4x32_10(j, direction, iteration, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const double u0 = u53(o[0], o[1]), u1 = u53(o[2], o[3]);
  d0 = dist == 0 ? 2.0 * u0 - 1.0 : u0;
  d1 = dist == 0 ? 2.0 * u1 - 1.0 : u1;
}

}  // namespace swm

// Host-side dispatch of the lane-split rollout kernel for one segment count N (see lane_rollout.cuh).
// Each lane_nK.cu translation unit instantiates launch_lane_rollout_n<K>.
#pragma once
#include "lane2_rollout.cuh"
#include "launch.cuh"

namespace swm {

// lanes per environment of the lane-split kernel for an n-segment swimmer (LaneSplit<N>::L)
inline int lane_split_lanes(int n) { return (n + 1 <= 4) ? 4 : (n + 1 <= 8) ? 8 : 16; }

// what the lane-split kernel implements: gym dynamics, fixed actions or linear policies (V1 / V2, with or
// without moments); no per-step screening, no action clipping
inline bool lane_split_supported(const RolloutArgs& a, const RolloutFlags& f) {
  if (f.variant != SWM_DYN_GYM || f.screen || a.clip) return false;
  if (!f.linear) return !f.norm && !f.stats;
  return !(f.stats && !f.norm);
}

// two_warps: the warp-specialised form (lane2_rollout.cuh: main warp + factorisation warp per lane group)
template <int N> int launch_lane_rollout_n(const RolloutArgs& a, const RolloutFlags& f, bool two_warps, cudaStream_t st);

#ifdef SWM_INSTANTIATE_LANE_N

template <int N, bool LINEAR, bool NORM, bool STATS>
static int launch_lane_one(const RolloutArgs& a, bool two_warps, cudaStream_t st) {
  constexpr int G = LaneSplit<N>::G;
  const long long blocks = (a.B + G - 1) / G;
  if (blocks > 0x7fffffffLL) return SWM_ERR_BAD_ARG;
  if (two_warps) lane2_rollout_kernel<N, LINEAR, NORM, STATS><<<(unsigned)blocks, kLane2Block, 0, st>>>(a);
  else lane_rollout_kernel<N, LINEAR, NORM, STATS><<<(unsigned)blocks, kLaneBlock, 0, st>>>(a);
  return cudaPeekAtLastError() == cudaSuccess ? SWM_OK : SWM_ERR_CUDA;
}

template <int N>
int launch_lane_rollout_n(const RolloutArgs& a, const RolloutFlags& f, bool two_warps, cudaStream_t st) {
  if (!lane_split_supported(a, f)) return SWM_ERR_UNSUPPORTED;
  if (!f.linear) return launch_lane_one<N, false, false, false>(a, two_warps, st);
  if (f.norm && f.stats) return launch_lane_one<N, true, true, true>(a, two_warps, st);
  if (f.norm) return launch_lane_one<N, true, true, false>(a, two_warps, st);
  return launch_lane_one<N, true, false, false>(a, two_warps, st);
}

template int launch_lane_rollout_n<SWM_INSTANTIATE_LANE_N>(const RolloutArgs&, const RolloutFlags&, bool, cudaStream_t);

#endif  // SWM_INSTANTIATE_LANE_N

}  // namespace swm

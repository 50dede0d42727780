// Host-side dispatch of the lane-split rollout kernel for one segment count N (see lane_rollout.cuh).
// Each lane_nK.cu translation unit instantiates launch_lane_rollout_n<K>.
#pragma once
#include "lane2_rollout.cuh"
#include "launch.cuh"

namespace swm {

// lanes per environment of the lane-split kernel for an n-segment swimmer (LaneSplit<N>::L)
inline int lane_split_lanes(int n) { return (n + 1 <= 4) ? 4 : (n + 1 <= 8) ? 8 : 16; }

// what the lane-split kernels implement: gym dynamics, fixed actions or linear policies (V1 / V2, with or
// without moments), no action clipping; per-step screening only in the warp-specialised kernels (f_warps >= 1) and
// only for plain linear policies
inline bool lane_split_supported(const RolloutArgs& a, const RolloutFlags& f, int f_warps = 0) {
  if (f.variant != SWM_DYN_GYM || a.clip) return false;
  if (f.screen) return f_warps >= 1 && f.linear && !f.norm && !f.stats;
  if (!f.linear) return !f.norm && !f.stats;
  return !(f.stats && !f.norm);
}

// f_warps: 0 = one warp per lane group (lane_rollout.cuh); 1 / 2 = the warp-specialised form with that many
// operator warps beside the main warp (lane2_rollout.cuh)
template <int N> int launch_lane_rollout_n(const RolloutArgs& a, const RolloutFlags& f, int f_warps, cudaStream_t st);

#ifdef SWM_INSTANTIATE_LANE_N

template <int N, bool LINEAR, bool NORM, bool STATS>
static int launch_lane_one(const RolloutArgs& a, int f_warps, cudaStream_t st) {
  constexpr int G = LaneSplit<N>::G;
  const long long blocks = (a.B + G - 1) / G;
  if (blocks > 0x7fffffffLL) return SWM_ERR_BAD_ARG;
  if (f_warps == 2) lane2_rollout_kernel<N, LINEAR, NORM, STATS, 2><<<(unsigned)blocks, 96, 0, st>>>(a);
  else if (f_warps == 1) lane2_rollout_kernel<N, LINEAR, NORM, STATS, 1><<<(unsigned)blocks, 64, 0, st>>>(a);
  else lane_rollout_kernel<N, LINEAR, NORM, STATS><<<(unsigned)blocks, kLaneBlock, 0, st>>>(a);
  return cudaPeekAtLastError() == cudaSuccess ? SWM_OK : SWM_ERR_CUDA;
}

template <int N>
static int launch_lane_screen(const RolloutArgs& a, int f_warps, cudaStream_t st) {
  constexpr int G = LaneSplit<N>::G;
  const long long blocks = (a.B + G - 1) / G;
  if (blocks > 0x7fffffffLL) return SWM_ERR_BAD_ARG;
  if (f_warps == 2) lane2_rollout_kernel<N, true, false, false, 2, true><<<(unsigned)blocks, 96, 0, st>>>(a);
  else lane2_rollout_kernel<N, true, false, false, 1, true><<<(unsigned)blocks, 64, 0, st>>>(a);
  return cudaPeekAtLastError() == cudaSuccess ? SWM_OK : SWM_ERR_CUDA;
}

template <int N>
int launch_lane_rollout_n(const RolloutArgs& a, const RolloutFlags& f, int f_warps, cudaStream_t st) {
  if (!lane_split_supported(a, f, f_warps)) return SWM_ERR_UNSUPPORTED;
  if (f.screen) return launch_lane_screen<N>(a, f_warps, st);
  if (!f.linear) return launch_lane_one<N, false, false, false>(a, f_warps, st);
  if (f.norm && f.stats) return launch_lane_one<N, true, true, true>(a, f_warps, st);
  if (f.norm) return launch_lane_one<N, true, true, false>(a, f_warps, st);
  return launch_lane_one<N, true, false, false>(a, f_warps, st);
}

template int launch_lane_rollout_n<SWM_INSTANTIATE_LANE_N>(const RolloutArgs&, const RolloutFlags&, int, cudaStream_t);

#endif  // SWM_INSTANTIATE_LANE_N

}  // namespace swm

// lane-split rollout kernels for 3-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 3
#include "lane_launch.cuh"

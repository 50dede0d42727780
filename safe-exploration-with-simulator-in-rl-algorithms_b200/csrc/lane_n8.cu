// lane-split rollout kernels for 8-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 8
#include "lane_launch.cuh"

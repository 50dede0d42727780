// lane-split rollout kernels for 10-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 10
#include "lane_launch.cuh"

// lane-split rollout kernels for 6-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 6
#include "lane_launch.cuh"

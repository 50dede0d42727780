// lane-split rollout kernels for 7-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 7
#include "lane_launch.cuh"

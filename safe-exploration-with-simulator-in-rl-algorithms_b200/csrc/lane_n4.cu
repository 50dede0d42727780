// lane-split rollout kernels for 4-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 4
#include "lane_launch.cuh"

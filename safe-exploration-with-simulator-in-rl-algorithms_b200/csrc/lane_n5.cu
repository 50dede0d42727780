// lane-split rollout kernels for 5-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 5
#include "lane_launch.cuh"

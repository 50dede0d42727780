// lane-split rollout kernels for 9-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 9
#include "lane_launch.cuh"

// lane-split rollout kernels for 2-segment swimmers (see lane_rollout.cuh)
#define SWM_INSTANTIATE_LANE_N 2
#include "lane_launch.cuh"

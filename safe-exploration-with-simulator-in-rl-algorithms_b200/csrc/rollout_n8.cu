// Instantiates the step and rollout kernels for n = 8 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 8
#include "launch.cuh"

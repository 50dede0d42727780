// Instantiates the step and rollout kernels for n = 3 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 3
#include "launch.cuh"

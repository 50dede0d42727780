// Instantiates the step and rollout kernels for n = 2 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 2
#include "launch.cuh"

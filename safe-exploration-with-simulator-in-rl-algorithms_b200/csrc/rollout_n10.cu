// Instantiates the step and rollout kernels for n = 10 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 10
#include "launch.cuh"

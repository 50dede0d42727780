// Instantiates the step and rollout kernels for n = 5 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 5
#include "launch.cuh"

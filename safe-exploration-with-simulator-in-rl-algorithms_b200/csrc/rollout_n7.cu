// Instantiates the step and rollout kernels for n = 7 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 7
#include "launch.cuh"

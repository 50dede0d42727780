// Instantiates the step and rollout kernels for n = 6 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 6
#include "launch.cuh"

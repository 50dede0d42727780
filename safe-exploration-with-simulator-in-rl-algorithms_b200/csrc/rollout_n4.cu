// Instantiates the step and rollout kernels for n = 4 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 4
#include "launch.cuh"

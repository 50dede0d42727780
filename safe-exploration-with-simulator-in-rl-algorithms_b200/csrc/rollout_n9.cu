// Instantiates the step and rollout kernels for n = 9 segments (see launch.cuh).
#define SWM_INSTANTIATE_N 9
#include "launch.cuh"

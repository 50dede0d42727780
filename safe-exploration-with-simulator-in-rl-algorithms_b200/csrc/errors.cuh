// Launch-error bookkeeping shared by the translation units of libswimmer_ars.so.
#pragma once
namespace swm {
// SWM_OK, or SWM_ERR_CUDA after recording (and clearing) the pending CUDA error for swm_last_cuda_error().
int check_launch();
}  // namespace swm

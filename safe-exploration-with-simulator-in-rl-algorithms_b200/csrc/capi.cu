// extern "C" entry points of libswimmer_ars.so (see include/swimmer_ars.h): argument checks,
// host-side precomputation of the physical constants and dispatch on the segment count.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "errors.cuh"
#include "lane_launch.cuh"

using namespace swm;

namespace swm {
Phys make_phys_public(const swm_params_t* p);
}

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return SWM_ERR_CUDA;
}

int note_launch(int rc) {
  if (rc == SWM_ERR_CUDA) {
    cudaError_t e = cudaPeekAtLastError();
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
  }
  return rc;
}

}  // namespace

int swm::check_launch() {
  const cudaError_t e = cudaGetLastError();  // reads and clears
  if (e == cudaSuccess) return SWM_OK;
  return cuda_fail(e);
}

namespace {

bool params_ok(const swm_params_t* p) {
  return p && p->n >= SWM_MIN_SEGMENTS && p->n <= SWM_MAX_SEGMENTS && p->l_i > 0.0 && p->m_i > 0.0;
}

Phys make_phys(const swm_params_t* p) {
  Phys P;
  const double l = p->l_i, m = p->m_i, k = p->k, n = (double)p->n;
  P.l = l; P.m = m; P.k = k; P.h = p->h; P.max_u = p->max_u;
  P.dirx = p->direction[0]; P.diry = p->direction[1];
  P.inv_l = 1.0 / l;
  P.kappa = k * l / m;
  P.m2kappa = -2.0 * (k * l / m);
  P.u_scale = 12.0 / (m * l * l);
  P.gdd_c = l / (2.0 * n);
  P.h_gdd_c = p->h * (l / (2.0 * n));
  P.inv_n = 1.0 / n;
  P.kl = k * l;
  P.tau_c = k * (l * l * l) / 12.0;
  P.half_l = l / 2.0;
  P.I = m * (l * l) / 12.0;
  return P;
}

}  // namespace

// shared with rlglue_protocol.cu
Phys swm::make_phys_public(const swm_params_t* p) { return make_phys(p); }

namespace {

#define SWM_DISPATCH_N(n, CALL)                 \
  switch (n) {                                  \
    case 2: return note_launch(CALL(2));        \
    case 3: return note_launch(CALL(3));        \
    case 4: return note_launch(CALL(4));        \
    case 5: return note_launch(CALL(5));        \
    case 6: return note_launch(CALL(6));        \
    case 7: return note_launch(CALL(7));        \
    case 8: return note_launch(CALL(8));        \
    case 9: return note_launch(CALL(9));        \
    case 10: return note_launch(CALL(10));      \
    default: return SWM_ERR_UNSUPPORTED;        \
  }

int step_common(const swm_params_t* params, int variant, bool acc_only, const double* state_in,
                const double* action, double* out, double* reward, int64_t B, void* stream) {
  if (!params_ok(params) || B < 0) return SWM_ERR_BAD_ARG;
  if (variant != SWM_DYN_GYM && variant != SWM_DYN_RLGLUE) return SWM_ERR_BAD_ARG;
  if (B == 0) return SWM_OK;  // empty batch: nothing to do, pointers may be NULL
  if (!state_in || !action || !out) return SWM_ERR_BAD_ARG;
  const Phys P = make_phys(params);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K) launch_step_n<K>(P, variant, acc_only, state_in, action, out, reward, (long long)B, st)
  SWM_DISPATCH_N(params->n, CALL)
#undef CALL
}

}  // namespace

extern "C" int swm_abi_version(void) { return SWM_ABI_VERSION; }

extern "C" const char* swm_strerror(int status) {
  switch (status) {
    case SWM_OK: return "ok";
    case SWM_ERR_BAD_ARG: return "bad argument";
    case SWM_ERR_UNSUPPORTED: return "unsupported configuration";
    case SWM_ERR_CUDA: return "CUDA error (see swm_last_cuda_error)";
    case SWM_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

extern "C" const char* swm_last_cuda_error(void) { return g_cuda_err; }

extern "C" int swm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cuda_fail(e); return SWM_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return SWM_OK;
}

extern "C" int swm_step_batched(const swm_params_t* params, int variant, const double* state_in,
                                const double* action, double* state_out, double* reward,
                                int64_t B, void* stream) {
  return step_common(params, variant, false, state_in, action, state_out, reward, B, stream);
}

extern "C" int swm_step_batched_models(const swm_params_t* params, int n_models, int64_t envs_per_model,
                                       const double* state_in, const double* action, double* state_out,
                                       double* reward, void* stream) {
  if (!params || n_models < 1 || n_models > SWM_MAX_MODELS_PER_STEP || envs_per_model < 0) return SWM_ERR_BAD_ARG;
  PhysSet set;
  for (int i = 0; i < n_models; ++i) {
    if (!params_ok(&params[i]) || params[i].n != params[0].n) return SWM_ERR_BAD_ARG;
    set.p[i] = make_phys(&params[i]);
  }
  for (int i = n_models; i < SWM_MAX_MODELS_PER_STEP; ++i) set.p[i] = set.p[0];
  const long long B = (long long)n_models * envs_per_model;
  if (B == 0) return SWM_OK;
  if (!state_in || !action || !state_out) return SWM_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K) launch_step_models_n<K>(set, (long long)envs_per_model, state_in, action, state_out, reward, B, st)
  SWM_DISPATCH_N(params[0].n, CALL)
#undef CALL
}

extern "C" int swm_accelerations_batched(const swm_params_t* params, int variant,
                                         const double* state, const double* action, double* acc,
                                         int64_t B, void* stream) {
  return step_common(params, variant, true, state, action, acc, nullptr, B, stream);
}

namespace {

int sm_count_cached() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// Validates cfg and fills the kernel arguments; returns SWM_OK or an error.
int build_rollout_args(const swm_params_t* params, const swm_rollout_t* cfg, RolloutArgs& a, RolloutFlags& f) {
  if (!params_ok(params) || !cfg) return SWM_ERR_BAD_ARG;
  if (cfg->B < 0 || cfg->H < 0 || cfg->rollouts_per_policy < 1) return SWM_ERR_BAD_ARG;
  if (cfg->variant != SWM_DYN_GYM && cfg->variant != SWM_DYN_RLGLUE) return SWM_ERR_BAD_ARG;
  if (cfg->kernel < SWM_KERNEL_AUTO || cfg->kernel > SWM_KERNEL_LANES) return SWM_ERR_BAD_ARG;
  if (cfg->B > (int64_t)kRolloutBlock * 0x7fffffffLL) return SWM_ERR_BAD_ARG;
  memset(&a, 0, sizeof(a));
  f.variant = cfg->variant;
  f.norm = cfg->normalize != 0;
  f.stats = cfg->stats_partial != nullptr;
  f.screen = cfg->screen.enabled != 0;
  f.group_w = (cfg->rollouts_per_policy % 32) == 0;
  switch (cfg->policy_mode) {
    case SWM_POLICY_FIXED_ACTION:
      f.linear = false;
      break;
    case SWM_POLICY_EXPLICIT:
    case SWM_POLICY_PHILOX:
    case SWM_POLICY_DELTAS:
      if (cfg->B % cfg->rollouts_per_policy != 0) return SWM_ERR_BAD_ARG;
      if (cfg->policy_mode != SWM_POLICY_EXPLICIT && ((cfg->B / cfg->rollouts_per_policy) % 2) != 0)
        return SWM_ERR_BAD_ARG;  // +delta / -delta pairs
      if (cfg->policy_mode == SWM_POLICY_EXPLICIT && cfg->dir_mask) return SWM_ERR_BAD_ARG;
      f.linear = true;
      break;
    default:
      return SWM_ERR_BAD_ARG;
  }
  if (!f.linear && cfg->dir_mask) return SWM_ERR_BAD_ARG;
  if (cfg->init_state && cfg->init_state_count < 1) return SWM_ERR_BAD_ARG;
  if (f.screen) {
    if (!params_ok(&cfg->screen.sim) || cfg->screen.sim.n != params->n) return SWM_ERR_BAD_ARG;
    a.sim = make_phys(&cfg->screen.sim);
  }
  a.real = make_phys(params);
  a.H = cfg->H;
  a.R = cfg->rollouts_per_policy;
  a.policy_mode = cfg->policy_mode;
  a.clip = cfg->clip_actions;
  a.B = cfg->B;
  a.actions = cfg->actions;
  a.policies = cfg->policies;
  a.deltas = cfg->deltas;
  a.dir_mask = cfg->dir_mask;
  a.init_perturb = cfg->init_perturb;
  a.nu = cfg->nu;
  a.seed = cfg->philox.seed;
  a.iteration = cfg->philox.iteration;
  a.iter_dev = cfg->philox.iteration_dev;
  a.dir0 = cfg->philox.dir0;
  a.dist = cfg->philox.dist;
  a.mean = cfg->mean;
  a.inv_sigma = cfg->inv_sigma;
  a.init_state = cfg->init_state;
  a.init_count = cfg->init_state ? cfg->init_state_count : 1;
  a.returns = cfg->returns;
  a.final_state = cfg->final_state;
  a.trajectory = cfg->trajectory;
  a.stats_partial = cfg->stats_partial;
  a.stats_pivot = cfg->stats_pivot;
  a.sim_thresh = cfg->screen.sim_thresh;
  a.real_thresh = cfg->screen.real_thresh;
  a.violations = cfg->screen.violations;
  a.frozen_at = cfg->screen.frozen_at;
  a.accumulate = cfg->accumulate_returns;
  return SWM_OK;
}

// AUTO: the lane-split kernel wins while its warps (B * L / 32) still find a sub-partition of their own,
// or at most share one with a second warp: measured crossover, tools/lane_split_sweep.py.  It is tuned for
// chains of up to 7 segments (8 lanes per environment).
constexpr double kLaneSplitMaxWarpsPerSmsp = 2.0;

int choose_kernel(int n, const swm_rollout_t* cfg, const RolloutArgs& a, const RolloutFlags& f) {
  const bool ok = lane_split_supported(a, f);
  if (cfg->kernel == SWM_KERNEL_LANES) return ok ? SWM_KERNEL_LANES : SWM_ERR_UNSUPPORTED;
  if (cfg->kernel == SWM_KERNEL_THREAD || !ok || n > 7) return SWM_KERNEL_THREAD;
  const int per_warp = 32 / lane_split_lanes(n);
  const double warps = (double)((cfg->B + per_warp - 1) / per_warp);
  return warps <= kLaneSplitMaxWarpsPerSmsp * 4.0 * sm_count_cached() ? SWM_KERNEL_LANES : SWM_KERNEL_THREAD;
}

}  // namespace

extern "C" int swm_rollout_kernel_choice(const swm_params_t* params, const swm_rollout_t* cfg) {
  RolloutArgs a;
  RolloutFlags f;
  const int rc = build_rollout_args(params, cfg, a, f);
  if (rc != SWM_OK) return rc;
  return choose_kernel(params->n, cfg, a, f);
}

extern "C" int64_t swm_rollout_stats_blocks(const swm_params_t* params, const swm_rollout_t* cfg) {
  if (!cfg || cfg->B < 1) return 0;
  // the row count only depends on the kernel choice, which never looks at the output pointers
  swm_rollout_t probe = *cfg;
  static double dummy;
  probe.stats_partial = &dummy;
  if (swm_rollout_kernel_choice(params, &probe) == SWM_KERNEL_LANES) {
    const int per_warp = 32 / lane_split_lanes(params->n);
    return (cfg->B + per_warp - 1) / per_warp;
  }
  return (cfg->B + kRolloutBlock - 1) / kRolloutBlock;
}

extern "C" int swm_rollout(const swm_params_t* params, const swm_rollout_t* cfg, void* stream) {
  RolloutArgs a;
  RolloutFlags f;
  if (params_ok(params) && cfg && cfg->B == 0 && cfg->H >= 0 && cfg->rollouts_per_policy >= 1 &&
      (cfg->variant == SWM_DYN_GYM || cfg->variant == SWM_DYN_RLGLUE))
    return SWM_OK;  // empty batch: nothing to do, pointers may be NULL
  const int rc = build_rollout_args(params, cfg, a, f);
  if (rc != SWM_OK) return rc;
  if (!cfg->returns) return SWM_ERR_BAD_ARG;
  if (!f.linear && !cfg->actions) return SWM_ERR_BAD_ARG;
  if (f.linear && !cfg->policies) return SWM_ERR_BAD_ARG;
  if (cfg->policy_mode == SWM_POLICY_DELTAS && !cfg->deltas) return SWM_ERR_BAD_ARG;
  if (f.norm && (!cfg->mean || !cfg->inv_sigma)) return SWM_ERR_BAD_ARG;
  if (f.stats && !cfg->stats_pivot) return SWM_ERR_BAD_ARG;
  const int kernel = choose_kernel(params->n, cfg, a, f);
  if (kernel < 0) return kernel;
  cudaStream_t st = (cudaStream_t)stream;
  if (kernel == SWM_KERNEL_LANES) {
#define CALL(K) launch_lane_rollout_n<K>(a, f, st)
    SWM_DISPATCH_N(params->n, CALL)
#undef CALL
  }
#define CALL(K) launch_rollout_n<K>(a, f, st)
  SWM_DISPATCH_N(params->n, CALL)
#undef CALL
}

// Layout self-check for language bindings: sizes of the ABI structs as this build sees them.
extern "C" int swm_abi_struct_sizes(int* params, int* philox, int* screen, int* rollout) {
  if (params) *params = (int)sizeof(swm_params_t);
  if (philox) *philox = (int)sizeof(swm_philox_t);
  if (screen) *screen = (int)sizeof(swm_screen_t);
  if (rollout) *rollout = (int)sizeof(swm_rollout_t);
  return SWM_OK;
}

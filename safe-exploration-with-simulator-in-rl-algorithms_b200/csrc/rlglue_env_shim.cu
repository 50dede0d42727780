// RL-Glue environment plugin ABI (env_init / env_start / env_step / env_cleanup / env_message)
// backed by the sm_100a kernels: the B200-native counterpart of
// rlglue/environment/SwimmerEnvironment.cpp:14-97.  Same symbols, same ownership rules (the callee
// owns every returned pointer; pointers stay valid until the next call; arrays are allocated in
// env_init and released in env_cleanup), same messages, same start state (all entries 0.001,
// cpp:39-42), same parameter-file format (cpp:297-326).  One environment = a batch of 1 through
// swm_step_batched(variant = SWM_DYN_RLGLUE); the batched entry points are the fast path, this shim
// exists so that an RL-Glue experiment can link the new library unchanged.
// Differences, on purpose: no per-step printing of the state (cpp:61), and an out-of-range action
// returns the previous observation with terminal = 1 instead of assert-aborting (cpp:56-58) when
// SWM_RLGLUE_NO_ABORT is set in the environment (default: abort like the reference).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/swimmer_rlglue_env.h"

namespace {

observation_t this_observation;
observation_t saved_observation;
reward_observation_terminal_t this_reward_observation;
swm_params_t g_params = {3, 0, 1.0, 1.0, 10.0, 0.01, 5.0, {1.0, 0.0}};
double* d_state = nullptr;   // [2n+2]
double* d_action = nullptr;  // [n-1]
double* d_reward = nullptr;  // [1]
std::string g_task_spec, g_param_msg;

void die(const char* what) {
  fprintf(stderr, "swimmer rlglue shim: %s (%s)\n", what, swm_last_cuda_error());
  abort();
}

void set_parameters(const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) { fprintf(stderr, "Unable to open file for setting environment parameters\n"); return; }
  char name[128];
  char line[512];
  while (fgets(line, sizeof(line), f)) {
    double a = 0, b = 0;
    const int got = sscanf(line, "%127s %lf %lf", name, &a, &b);
    if (got < 2) continue;
    if (!strcmp(name, "n_seg")) g_params.n = (int)a;
    else if (!strcmp(name, "max_u")) g_params.max_u = a;
    else if (!strcmp(name, "l_i")) g_params.l_i = a;
    else if (!strcmp(name, "k")) g_params.k = a;
    else if (!strcmp(name, "m_i")) g_params.m_i = a;
    else if (!strcmp(name, "h_global")) g_params.h = a;
    else if (!strcmp(name, "direction") && got == 3) { g_params.direction[0] = a; g_params.direction[1] = b; }
  }
  fclose(f);
}

void copy_obs(observation_t& dst, const observation_t& src) {
  dst.numInts = src.numInts; dst.numDoubles = src.numDoubles; dst.numChars = src.numChars;
  for (unsigned i = 0; i < src.numDoubles; ++i) dst.doubleArray[i] = src.doubleArray[i];
}

}  // namespace

extern "C" {

void allocateRLStruct(rl_abstract_type_t* dst, unsigned int numInts, unsigned int numDoubles,
                      unsigned int numChars) {
  dst->numInts = numInts; dst->numDoubles = numDoubles; dst->numChars = numChars;
  dst->intArray = numInts ? (int*)calloc(numInts, sizeof(int)) : nullptr;
  dst->doubleArray = numDoubles ? (double*)calloc(numDoubles, sizeof(double)) : nullptr;
  dst->charArray = numChars ? (char*)calloc(numChars + 1, 1) : nullptr;
}

void clearRLStruct(rl_abstract_type_t* dst) {
  free(dst->intArray); free(dst->doubleArray); free(dst->charArray);
  dst->intArray = nullptr; dst->doubleArray = nullptr; dst->charArray = nullptr;
  dst->numInts = dst->numDoubles = dst->numChars = 0;
}

/* Programmatic alternative to the "set parameters" message (the reference has file globals). */
int swm_rlglue_set_params(const swm_params_t* p) {
  if (!p || p->n < SWM_MIN_SEGMENTS || p->n > SWM_MAX_SEGMENTS) return SWM_ERR_BAD_ARG;
  g_params = *p;
  return SWM_OK;
}

const char* env_init(void) {
  const int n_obs = 2 + 2 * g_params.n, n_action = g_params.n - 1;
  allocateRLStruct(&this_observation, 0, n_obs, 0);
  allocateRLStruct(&saved_observation, 0, n_obs, 0);
  this_reward_observation.observation = &this_observation;
  this_reward_observation.reward = 0;
  this_reward_observation.terminal = 0;
  if (cudaMalloc(&d_state, sizeof(double) * n_obs) != cudaSuccess ||
      cudaMalloc(&d_action, sizeof(double) * (n_action > 0 ? n_action : 1)) != cudaSuccess ||
      cudaMalloc(&d_reward, sizeof(double)) != cudaSuccess)
    die("cudaMalloc failed");
  g_task_spec = "VERSION RL-Glue-3.0 PROBLEMTYPE continuing DISCOUNTFACTOR 0.9 OBSERVATIONS DOUBLES (" +
                std::to_string(n_obs) + " UNSPEC UNSPEC) ACTIONS DOUBLES (" + std::to_string(n_action) +
                " " + std::to_string(-g_params.max_u) + " " + std::to_string(g_params.max_u) +
                ") REWARDS (UNSPEC UNSPEC) EXTRA SwimmerEnvironment(B200) swimmer_ars";
  return g_task_spec.c_str();
}

const observation_t* env_start(void) {
  for (unsigned i = 0; i < this_observation.numDoubles; ++i) this_observation.doubleArray[i] = 0.001;
  copy_obs(saved_observation, this_observation);
  return &this_observation;
}

const reward_observation_terminal_t* env_step(const action_t* this_action) {
  const int n_obs = 2 + 2 * g_params.n, n_action = g_params.n - 1;
  bool valid = (int)this_action->numDoubles == n_action;
  for (int i = 0; valid && i < n_action; ++i)
    valid = fabs(this_action->doubleArray[i]) <= g_params.max_u;
  if (!valid) {
    if (!getenv("SWM_RLGLUE_NO_ABORT")) { fprintf(stderr, "env_step: invalid action\n"); abort(); }
    this_reward_observation.terminal = 1;
    return &this_reward_observation;
  }
  if (cudaMemcpy(d_state, this_observation.doubleArray, sizeof(double) * n_obs, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_action, this_action->doubleArray, sizeof(double) * n_action, cudaMemcpyHostToDevice) != cudaSuccess)
    die("H2D copy failed");
  if (swm_step_batched(&g_params, SWM_DYN_RLGLUE, d_state, d_action, d_state, d_reward, 1, nullptr) != SWM_OK)
    die("swm_step_batched failed");
  double reward = 0.0;
  if (cudaMemcpy(this_observation.doubleArray, d_state, sizeof(double) * n_obs, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(&reward, d_reward, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
    die("D2H copy failed");
  this_reward_observation.observation = &this_observation;
  this_reward_observation.reward = reward;
  this_reward_observation.terminal = 0;
  return &this_reward_observation;
}

void env_cleanup(void) {
  clearRLStruct(&this_observation);
  clearRLStruct(&saved_observation);
  cudaFree(d_state); cudaFree(d_action); cudaFree(d_reward);
  d_state = d_action = d_reward = nullptr;
}

const char* env_message(const char* message) {
  if (strcmp(message, "what is your name?") == 0) return "My name is swimmer_environment, B200 edition!";
  if (strcmp(message, "save state") == 0) {
    copy_obs(saved_observation, this_observation);
    return "saved_observation has the value of this_observation";
  }
  if (strcmp(message, "load state") == 0) {
    copy_obs(this_observation, saved_observation);
    return "this_observation has the value of saved_observation";
  }
  if (strncmp(message, "set parameters", 14) == 0) {
    const char* path = message[14] == ' ' ? message + 15 : getenv("SWM_PARAMETERS_FILE");
    set_parameters(path ? path : "../parameters.txt");
    g_param_msg = "Environment parameters are: n_seg=" + std::to_string(g_params.n) +
                  "; max_u=" + std::to_string(g_params.max_u) + "; l_i=" + std::to_string(g_params.l_i) +
                  "; k=" + std::to_string(g_params.k) + "; m_i=" + std::to_string(g_params.m_i) +
                  "; h_global=" + std::to_string(g_params.h);
    return g_param_msg.c_str();
  }
  return "SwimmerEnvironment(B200) does not respond to that message.";
}

}  // extern "C"

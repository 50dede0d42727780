// TEST INFRASTRUCTURE ONLY.  Thin extern "C" access to the *unmodified* reference
// rlglue/environment/SwimmerEnvironment.cpp (compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/).  Used to validate the restated rlglue variant and as
// a timed CPU reference.  Nothing here restates the reference's arithmetic: it only sets
// the reference's globals (cpp:3-9) and calls updateState (cpp:102) /
// compute_accelerations (cpp:139), never env_step, which prints every step (cpp:61).
#include "SwimmerEnvironment.h"
#include <cstdlib>

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

void ref_set_params(int n, double max_u_, double l, double k_, double m, double h,
                    double dx, double dy) {
  n_seg = (size_t)n; max_u = max_u_; l_i = l; k = k_; m_i = m; h_global = h;
  direction = Vector2d(dx, dy);
}

// one semi-implicit step in place; state[2n+2], action[n-1]
void ref_update_state(double* state, const double* action) {
  observation_t obs; obs.numInts = 0; obs.numChars = 0; obs.intArray = nullptr; obs.charArray = nullptr;
  obs.numDoubles = (unsigned)(2 + 2 * n_seg); obs.doubleArray = state;
  action_t act; act.numInts = 0; act.numChars = 0; act.intArray = nullptr; act.charArray = nullptr;
  act.numDoubles = (unsigned)(n_seg - 1); act.doubleArray = const_cast<double*>(action);
  updateState(obs, &act);
}

// steps x updateState from `state`, fixed action; returns sum of rewards (calculate_reward cpp:273)
double ref_rollout_fixed(double* state, const double* action, int steps) {
  observation_t obs; obs.numInts = 0; obs.numChars = 0; obs.intArray = nullptr; obs.charArray = nullptr;
  obs.numDoubles = (unsigned)(2 + 2 * n_seg); obs.doubleArray = state;
  double total = 0.0;
  for (int t = 0; t < steps; ++t) { ref_update_state(state, action); total += calculate_reward(obs); }
  return total;
}

void ref_compute_accelerations(const double* state, const double* action, double* gdd, double* thdd) {
  std::vector<double> torque(action, action + (n_seg - 1));
  Vector2d G_dot(state[0], state[1]);
  std::vector<double> theta, theta_dot;
  for (size_t i = 0; i < n_seg; i++) { theta.push_back(state[2 + 2 * i]); theta_dot.push_back(state[3 + 2 * i]); }
  Vector2d G_dotdot(0., 0.);
  std::vector<double> theta_dotdot;
  compute_accelerations(torque, G_dot, theta, theta_dot, G_dotdot, theta_dotdot);
  gdd[0] = G_dotdot(0); gdd[1] = G_dotdot(1);
  for (size_t i = 0; i < n_seg; i++) thdd[i] = theta_dotdot[i];
}

}  // extern "C"

"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (C restatement of the
reference hot path) and of oracle/_ref/libref_swimmer.so (the unmodified reference C++
swimmer, when it was built).  Used by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs -- never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)

GYM, RLGLUE = 0, 1


class Params(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("l_i", ctypes.c_double), ("m_i", ctypes.c_double),
                ("k", ctypes.c_double), ("h", ctypes.c_double), ("max_u", ctypes.c_double),
                ("direction", ctypes.c_double * 2)]


def make_params(n=3, l_i=1.0, m_i=1.0, k=10.0, h=0.001, max_u=5.0, direction=(1.0, 0.0)):
    p = Params()
    p.n, p.l_i, p.m_i, p.k, p.h, p.max_u = n, l_i, m_i, k, h, max_u
    p.direction[0], p.direction[1] = direction
    return p


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "swimmer_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref():
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        pp = ctypes.POINTER(Params)
        L.orc_gym_accelerations.argtypes = [pp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_rlglue_accelerations.argtypes = [pp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_step.argtypes = [pp, ctypes.c_int, _dp, _dp]
        L.orc_step.restype = ctypes.c_double
        L.orc_rollout.argtypes = [pp, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, _dp,
                                  ctypes.c_int, _dp, _dp]
        L.orc_rollout.restype = ctypes.c_double
        L.orc_rollout_safe_step.argtypes = [pp, pp, _dp, _dp, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_double, _dp, _dp, _ip, _ip]
        L.orc_rollout_safe_step.restype = ctypes.c_double
        L.orc_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.orc_philox_delta.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _dp]
        L.orc_sort_directions.argtypes = [_dp, ctypes.c_int, _ip]
        L.orc_update_policy.argtypes = [_dp, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_int]
        L.orc_update_policy.restype = ctypes.c_double
        L.orc_mean_var.argtypes = [_dp, ctypes.c_long, ctypes.c_int, _dp, _dp]
        L.orc_threshold_alpha.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int]
        L.orc_threshold_alpha.restype = ctypes.c_double
        L.orc_rollout_fixed_batch.argtypes = [pp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_long,
                                              ctypes.c_long, _dp, _dp]
        L.orc_rollout_policy_batch.argtypes = [pp, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int,
                                               ctypes.c_long, ctypes.c_long, _dp]
        _lib = L
    return _lib


# ---- convenience wrappers (numpy in / numpy out) -------------------------------------------

def accelerations(p, variant, state, action):
    n = p.n
    state = np.ascontiguousarray(state, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    gd = np.ascontiguousarray(state[:2])
    th = np.ascontiguousarray(state[2::2])
    thd = np.ascontiguousarray(state[3::2])
    gdd, thdd = np.zeros(2), np.zeros(n)
    f = lib().orc_gym_accelerations if variant == GYM else lib().orc_rlglue_accelerations
    rc = f(ctypes.byref(p), _d(action), _d(gd), _d(th), _d(thd), _d(gdd), _d(thdd))
    assert rc == 0, rc
    return gdd, thdd


def step(p, variant, state, action):
    st = np.array(state, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    r = lib().orc_step(ctypes.byref(p), variant, _d(st), _d(action))
    return st, r


def step_batch(p, variant, states, actions):
    states = np.array(states, dtype=np.float64)
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    rewards = np.zeros(len(states))
    for i in range(len(states)):
        rewards[i] = lib().orc_step(ctypes.byref(p), variant, _d(states[i]), _d(actions[i]))
    return states, rewards


def rollout(p, variant, H, action=None, policy=None, mean=None, inv_sigma=None, clip=False,
            init_state=None, want_traj=False):
    no = 2 * p.n + 2
    if policy is not None:
        W = np.ascontiguousarray(policy, dtype=np.float64).reshape(-1)
        mode = 1
    else:
        W = np.ascontiguousarray(action, dtype=np.float64)
        mode = 0
    mean = None if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
    inv_sigma = None if inv_sigma is None else np.ascontiguousarray(inv_sigma, dtype=np.float64)
    init = None if init_state is None else np.ascontiguousarray(init_state, dtype=np.float64)
    traj = np.zeros((H, no)) if want_traj else None
    final = np.zeros(no)
    r = lib().orc_rollout(ctypes.byref(p), variant, mode, _d(W), _d(mean), _d(inv_sigma),
                          int(clip), _d(init), H, _d(traj), _d(final))
    return r, final, traj


def rollout_safe_step(real_p, sim_p, policy, H, sim_thresh, real_thresh, init_state=None,
                      want_traj=False):
    no = 2 * real_p.n + 2
    W = np.ascontiguousarray(policy, dtype=np.float64).reshape(-1)
    init = None if init_state is None else np.ascontiguousarray(init_state, dtype=np.float64)
    traj = np.zeros((H, no)) if want_traj else None
    final = np.zeros(no)
    viol, frozen = ctypes.c_int(0), ctypes.c_int(0)
    r = lib().orc_rollout_safe_step(ctypes.byref(real_p), ctypes.byref(sim_p), _d(W), _d(init), H,
                                    sim_thresh, real_thresh, _d(traj), _d(final),
                                    ctypes.byref(viol), ctypes.byref(frozen))
    return r, final, traj, viol.value, frozen.value


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def philox_delta(seed, iteration, direction, count, dist=0, stream=0):
    out = np.zeros(count)
    lib().orc_philox_delta(seed, iteration, direction, stream, dist, count, _d(out))
    return out


def sort_directions(returns):
    returns = np.ascontiguousarray(returns, dtype=np.float64)
    N = len(returns) // 2
    order = np.zeros(N, dtype=np.int32)
    lib().orc_sort_directions(_d(returns), N, order.ctypes.data_as(_ip))
    return order


def update_policy(W, deltas, returns, b, alpha, semantics):
    W = np.array(W, dtype=np.float64)
    flat = W.reshape(-1)
    deltas = np.ascontiguousarray(deltas, dtype=np.float64).reshape(len(deltas), -1)
    returns = np.ascontiguousarray(returns, dtype=np.float64)
    sigma = lib().orc_update_policy(_d(flat), flat.size, _d(deltas), _d(returns), len(deltas), b,
                                    alpha, semantics)
    return flat.reshape(W.shape), sigma


def mean_var(states):
    states = np.ascontiguousarray(states, dtype=np.float64)
    mean, var = np.zeros(states.shape[1]), np.zeros(states.shape[1])
    lib().orc_mean_var(_d(states), states.shape[0], states.shape[1], _d(mean), _d(var))
    return mean, var


def threshold_alpha(K, A, B, H):
    return lib().orc_threshold_alpha(K, A, B, H)


def rollout_fixed_batch(p, variant, actions, H, lo=0, hi=None, want_final=True):
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    B = len(actions)
    hi = B if hi is None else hi
    returns = np.zeros(B)
    final = np.zeros((B, 2 * p.n + 2)) if want_final else None
    lib().orc_rollout_fixed_batch(ctypes.byref(p), variant, _d(actions), H, lo, hi, _d(returns),
                                  _d(final))
    return returns, final


def rollout_policy_batch(p, variant, policies, H, mean=None, inv_sigma=None, lo=0, hi=None):
    policies = np.ascontiguousarray(policies, dtype=np.float64)
    B = len(policies)
    hi = B if hi is None else hi
    returns = np.zeros(B)
    mean = None if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
    inv_sigma = None if inv_sigma is None else np.ascontiguousarray(inv_sigma, dtype=np.float64)
    lib().orc_rollout_policy_batch(ctypes.byref(p), variant, _d(policies.reshape(B, -1)), _d(mean),
                                   _d(inv_sigma), H, lo, hi, _d(returns))
    return returns


# ---- the unmodified reference C++ swimmer (oracle/_ref) ------------------------------------
_ref = None


def ref_cpp():
    """ctypes handle of oracle/_ref/libref_swimmer.so or None if it was never built."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libref_swimmer.so")
        if not os.path.exists(so):
            return None
        R = ctypes.CDLL(so)
        R.ref_set_params.argtypes = [ctypes.c_int] + [ctypes.c_double] * 7
        R.ref_update_state.argtypes = [_dp, _dp]
        R.ref_compute_accelerations.argtypes = [_dp, _dp, _dp, _dp]
        R.ref_rollout_fixed.argtypes = [_dp, _dp, ctypes.c_int]
        R.ref_rollout_fixed.restype = ctypes.c_double
        _ref = R
    return _ref


def ref_cpp_set_params(p):
    ref_cpp().ref_set_params(p.n, p.max_u, p.l_i, p.k, p.m_i, p.h, p.direction[0], p.direction[1])


def ref_cpp_step(state, action):
    st = np.array(state, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    ref_cpp().ref_update_state(_d(st), _d(action))
    return st


def ref_cpp_accelerations(state, action, n):
    state = np.ascontiguousarray(state, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    gdd, thdd = np.zeros(2), np.zeros(n)
    ref_cpp().ref_compute_accelerations(_d(state), _d(action), _d(gdd), _d(thdd))
    return gdd, thdd

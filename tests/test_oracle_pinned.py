"""Pins the CPU oracle (oracle/swimmer_oracle.c) against fixtures generated from the UNMODIFIED
reference (oracle/make_golden.py) and against the reference's own golden text files.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err


@pytest.mark.parametrize("n", [2, 3, 5, 10])
def test_gym_step_matches_python_reference(O, n):
    g = golden("gym_step.npz")
    for t in range(len(g[f"n{n}_state"])):
        l, m, k, h = g[f"n{n}_params"][t]
        p = O.make_params(n=n, l_i=l, m_i=m, k=k, h=h, direction=(1., 0.) if t % 2 == 0 else (0.6, -0.8))
        nxt, r = O.step(p, O.GYM, g[f"n{n}_state"][t], g[f"n{n}_action"][t])
        assert rel_err(nxt, g[f"n{n}_next"][t]) < 1e-13
        assert abs(r - g[f"n{n}_reward"][t]) < 1e-13 * max(1, abs(r))
        gdd, thdd = O.accelerations(p, O.GYM, g[f"n{n}_state"][t], g[f"n{n}_action"][t])
        assert rel_err(np.concatenate([gdd, thdd]), g[f"n{n}_acc"][t]) < 1e-12


def test_gym_single_step_kat(O):
    # default physics, n=3, reset state, action [2.5, 2.5]
    p = O.make_params(n=3)
    st = np.array([0, 0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0.])
    nxt, _ = O.step(p, O.GYM, st, [2.5, 2.5])
    want = [-8.32667268468868e-19, -4.622231866529367e-35, 1.5707963267948966, -0.015000000000000006,
            1.5707963267948966, 3.4416913763379856e-18, 1.5707963267948966, 0.014999999999999996]
    assert rel_err(nxt, want) < 1e-15


def test_coulom_barycentre_acceleration(O):
    # rlglue/test/acceleration-compare.txt:4-6: Coulom's program, same state: Gdd_x = 0.284343
    gold = json.load(open(os.path.join(GOLDEN, "rlglue_golden.json")))
    p = O.make_params(n=3)
    gdd, _ = O.accelerations(p, O.GYM, gold["state"], gold["torque"])
    assert abs(gdd[0] - gold["coulom_barycenter_acc_6digits"][0]) < 5e-6


@pytest.mark.parametrize("n", [3, 5, 10])
def test_gym_rollouts_match_python_reference(O, n):
    g = golden("gym_rollout.npz")
    H = int(g[f"n{n}_H"])
    p = O.make_params(n=n, l_i=.8, m_i=1.2, k=10.2)
    for i in range(3):
        r, fin, traj = O.rollout(p, O.GYM, H, policy=g[f"n{n}_W"][i], want_traj=True)
        assert abs(r - g[f"n{n}_v1_return"][i]) < 1e-10 * max(1, abs(r))
        assert rel_err(fin, g[f"n{n}_v1_final"][i]) < 1e-10
        assert rel_err(traj[49::50], g[f"n{n}_v1_traj50"][i]) < 1e-10
        r, fin, traj = O.rollout(p, O.GYM, H, policy=g[f"n{n}_W"][i], mean=g[f"n{n}_mean"],
                                 inv_sigma=g[f"n{n}_var"] ** -0.5, want_traj=True)
        assert abs(r - g[f"n{n}_v2_return"][i]) < 1e-10 * max(1, abs(r))
        assert rel_err(fin, g[f"n{n}_v2_final"][i]) < 1e-10


def test_rlglue_golden_text(O):
    # rlglue/test/acceleration-compare.txt:102-103 and swimmer-compare.txt:100 (6 printed digits)
    gold = json.load(open(os.path.join(GOLDEN, "rlglue_golden.json")))
    p = O.make_params(n=3, h=gold["params"]["h_inferred"])
    gdd, thdd = O.accelerations(p, O.RLGLUE, gold["state"], gold["torque"])
    np.testing.assert_allclose(gdd, gold["G_dotdot_6digits"], rtol=5e-6)
    np.testing.assert_allclose(thdd, gold["theta_dotdot_6digits"], rtol=5e-6)
    nxt, _ = O.step(p, O.RLGLUE, gold["state"], gold["torque"])
    np.testing.assert_allclose(nxt, gold["state_after_update_6digits"], rtol=5e-6)


@pytest.mark.parametrize("n", [2, 3, 5, 10])
def test_rlglue_step_matches_compiled_reference_fixture(O, n):
    g = golden("rlglue_step.npz")
    for t in range(len(g[f"n{n}_state"])):
        l, m, k, h = g[f"n{n}_params"][t]
        p = O.make_params(n=n, l_i=l, m_i=m, k=k, h=h)
        nxt, _ = O.step(p, O.RLGLUE, g[f"n{n}_state"][t], g[f"n{n}_action"][t])
        assert rel_err(nxt, g[f"n{n}_next"][t]) < 1e-11
        gdd, thdd = O.accelerations(p, O.RLGLUE, g[f"n{n}_state"][t], g[f"n{n}_action"][t])
        assert rel_err(np.concatenate([gdd, thdd]), g[f"n{n}_acc"][t]) < 1e-9


def test_rlglue_rollout_fixture(O):
    g = golden("rlglue_step.npz")
    p = O.make_params(n=3, h=0.01)
    r, fin, _ = O.rollout(p, O.RLGLUE, 500, action=g["roll_action"])
    assert rel_err(fin, g["roll_final"]) < 1e-9
    assert abs(r - g["roll_return"]) < 1e-9 * max(1, abs(r))


def test_rlglue_live_compiled_reference(O):
    """When oracle/_ref was built (it ships to the GPU box), compare live on fresh random states."""
    if O.ref_cpp() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    from conftest import rand_states
    for n in (2, 3, 4, 7):
        p = O.make_params(n=n, l_i=1.1, m_i=.9, k=9., h=0.004)
        O.ref_cpp_set_params(p)
        for st in rand_states(rng, n, 10):
            a = rng.uniform(-5, 5, n - 1)
            assert rel_err(O.step(p, O.RLGLUE, st, a)[0], O.ref_cpp_step(st, a)) < 1e-11


def test_philox_known_answers(O):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    d = O.philox_delta(seed=7, iteration=3, direction=11, count=16)
    assert np.all(d >= -1) and np.all(d < 1)
    u = O.philox_delta(seed=7, iteration=3, direction=11, count=16, dist=1)
    np.testing.assert_array_equal(d, 2 * u - 1)
    # odd count: last element of a pair is dropped, not shifted
    np.testing.assert_array_equal(O.philox_delta(7, 3, 11, 15), d[:15])


def test_topb_matches_reference_sort(O):
    g = golden("topb.npz")
    for N in (1, 2, 8, 33, 256, 1024):
        np.testing.assert_array_equal(O.sort_directions(g[f"returns_{N}"]), g[f"order_{N}"])
    np.testing.assert_array_equal(O.sort_directions(g["returns_ties"]), g["order_ties"])


def test_ars_agent_update_matches_reference(O):
    g = golden("ars_agent.npz")
    # config 1 (N=8=b): replay the recorded perturbations and returns through the oracle update
    W = np.zeros((2, 8))
    for it in range(3):
        d = 2 * g["c1_rand"][it] - 1
        W, _ = O.update_policy(W, d, g["c1_returns"][it], b=8, alpha=0.0075, semantics=0)
        assert rel_err(W, g["c1_policies"][it]) < 1e-12
    # and the returns themselves: rollouts of W +- nu*delta
    p = O.make_params(n=3)
    d0 = 2 * g["c1_rand"][0] - 1
    for i in range(8):
        for s, sign in enumerate((+1, -1)):
            r, _, _ = O.rollout(p, O.GYM, 1000, policy=np.zeros((2, 8)) + sign * 0.01 * d0[i])
            assert abs(r - g["c1_returns"][0][2 * i + s]) < 1e-10 * max(1, abs(r))
    # first returns of config-1 iteration 0 with seed 0
    np.testing.assert_allclose(g["c1_returns"][0][:4], [-0.11878042019245531, 0.46055543070364563,
                                                        -3.532406000157358, -5.630701221472039], rtol=1e-12)


def test_v2_statistics_match_reference(O):
    g = golden("ars_agent.npz")
    p = O.make_params(n=3, l_i=.8, m_i=1.2, k=10.2)
    W = np.zeros((2, 8))
    mean, inv_sigma = np.zeros(8), np.ones(8)
    states = []
    for it in range(3):
        d = 2 * g["v2_rand"][it] - 1
        rets = []
        for i in range(4):
            for sign in (+1, -1):
                r, _, traj = O.rollout(p, O.GYM, 250, policy=W + sign * 0.01 * d[i], mean=mean,
                                       inv_sigma=inv_sigma, want_traj=True)
                rets.append(r); states.append(traj)
        np.testing.assert_allclose(rets, g["v2_returns"][it], rtol=1e-9, atol=1e-12)
        W, _ = O.update_policy(W, d, np.array(rets), b=4, alpha=0.0075, semantics=0)
        assert rel_err(W, g["v2_policies"][it]) < 1e-9
        mean, var = O.mean_var(np.concatenate(states))
        np.testing.assert_allclose(mean, g["v2_means"][it], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(var, g["v2_vars"][it], rtol=1e-10)
        inv_sigma = var ** -0.5


def test_basic_ars_train_matches_reference(O):
    g = golden("safe_ars.npz")
    p = O.make_params(n=3)
    W = np.zeros((2, 8))
    for it in range(3):
        d = 2 * g["basic_rand"][it] - 1
        rets = []
        for i in range(4):
            for sign in (+1, -1):
                rets.append(O.rollout(p, O.GYM, 200, policy=W + sign * 0.05 * d[i])[0])
        assert abs(np.mean(rets) - g["basic_curve"][it]) < 1e-12
        W, _ = O.update_policy(W, d, np.array(rets), b=2, alpha=0.02, semantics=1)
    assert rel_err(W, g["basic_policy"]) < 1e-10


def test_safe_step_rollouts_match_reference(O):
    g = golden("safe_ars.npz")
    m, l, k = g["sim_mlk"]
    real, sim = O.make_params(n=3), O.make_params(n=3, m_i=m, l_i=l, k=k)
    for i in range(6):
        r, fin, _, viol, frozen = O.rollout_safe_step(real, sim, g["safe_W"][i], 400,
                                                      g["safe_sim_thresh"][i], g["safe_real_thresh"][i])
        assert abs(r - g["safe_returns"][i]) < 1e-10 * max(1, abs(r))
        assert rel_err(fin, g["safe_finals"][i]) < 1e-10
        assert frozen == g["safe_frozen_from"][i]


def test_threshold_alpha(O):
    for K, A, B, H, want in golden("misc.npz")["threshold_alpha"]:
        assert abs(O.threshold_alpha(K, A, B, int(H)) - want) <= 1e-12 * abs(want)


def _rlglue_par(g, tag):
    n, N, b, H, alpha, nu, max_u, l_i, k, m_i, h = g[tag + "_par"]
    return dict(n_seg=int(n), N=int(N), b=int(b), H=int(H), alpha=alpha, nu=nu, max_u=max_u, l_i=l_i, k=k, m_i=m_i,
                h_global=h, direction=(1.0, 0.0))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rlglue_agent_protocol_restatement_matches_unmodified_agent(O, tag):
    """tests/golden/rlglue_agent.npz = the UNMODIFIED rlglue/agent/SwimmerAgent.py driven against the UNMODIFIED
    compiled C++ swimmer the way SwimmerExperiment.cpp:65-100 does (oracle/rlglue_protocol.py).  The restated
    step-level loop over the C port reproduces its evaluation returns and policies, which pins the agent state
    machine (:79-130), the clip (:181-201), the U[0,1) perturbations as drawn (:203-212) and the index-order /
    sample-stdev update (:214-241)."""
    from oracle import rlglue_protocol as RP
    g = golden("rlglue_agent.npz")
    par = _rlglue_par(g, tag)
    n_it = len(g[tag + "_results"])
    res, pol = RP.restated_protocol(par, n_it, g[tag + "_deltas"])
    np.testing.assert_allclose(res, g[tag + "_results"], rtol=1e-11)
    np.testing.assert_allclose(pol, g[tag + "_policies"], rtol=1e-9, atol=1e-13)
    # the oracle's update rule "semantics 2" (first b directions, sample stdev, divisor b) against the policy
    # steps the unmodified agent took from its own reward tables
    for it in range(n_it):
        W, _ = O.update_policy(g[tag + "_policies"][it], g[tag + "_deltas"][it], g[tag + "_rewards"][it],
                               b=par["b"], alpha=par["alpha"], semantics=2)
        np.testing.assert_allclose(W, g[tag + "_policies"][it + 1], rtol=1e-9, atol=1e-13)
    # the reference's literal quirk: in iteration 0 the last slot of the reward table holds one step's reward
    assert abs(g[tag + "_rewards"][0][-1]) < 2e-3 < abs(g[tag + "_rewards"][0][0])

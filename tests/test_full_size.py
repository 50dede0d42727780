"""BASELINE.json configurations at their full sizes: the CUDA path against the oracle on sampled
environments, plus size-independent properties (mirror symmetry, shard-invariance, moments of all
visited states).  Tolerance: H=1000 returns within 1e-6 relative (north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RET_TOL = 1e-6


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _close(got, want, tol=RET_TOL, floor=1e-3):
    return abs(got - want) <= tol * max(floor, abs(want))


def test_config2_sampled_against_oracle_and_mirror_symmetry(S, O):
    """config[1]: n=3, 65,536 envs, fixed random actions, 1,000 steps."""
    n, B, H = 3, 65536, 1000
    ps, po = S.make_params(n=n), O.make_params(n=n)
    rng = np.random.default_rng(0)
    ac = rng.uniform(-5, 5, (B, n - 1))
    res = S.ops.rollout(ps, H, actions=_cuda(ac), want_final=True)
    ret, fin = res.returns.cpu().numpy(), res.final_state.cpu().numpy()
    idx = np.concatenate([[0, 1, 63, 64, B - 1], rng.integers(0, B, 43)])
    want_r, want_f = O.rollout_fixed_batch(po, O.GYM, ac[idx], H)
    for i, e in enumerate(idx):
        assert _close(ret[e], want_r[i]), (e, ret[e], want_r[i])
    assert np.max(np.abs(fin[idx] - want_f)) < 1e-8 * max(1.0, np.max(np.abs(want_f)))
    # Reflecting the swimmer about its initial axis (theta -> pi - theta, u -> -u) flips Gdot_x: from
    # reset, R(-u) = -R(u) up to round-off of pi/2 (the reference has the same symmetry).
    mir = S.ops.rollout(ps, H, actions=_cuda(-ac[:4096])).returns.cpu().numpy()
    scale = np.maximum(1e-3, np.abs(ret[:4096]))
    assert np.max(np.abs(mir + ret[:4096]) / scale) < 1e-7


def test_config3_full_iteration_v2(S, O):
    """config[2]: n=5, ARS V2, 1,024 directions, H=1000: one engine iteration from a non-trivial
    policy and normalisation; sampled returns, the top-b order, the update and the V2 statistics
    against the oracle."""
    n, N, H, nu, alpha, seed = 5, 1024, 1000, 0.01, 0.0075, 11
    ps, po = S.make_params(n=n), O.make_params(n=n)
    no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
    rng = np.random.default_rng(5)
    W0 = rng.uniform(-1, 1, ws) * 0.2
    eng = S.ArsEngine(ps, N=N, b=N, alpha=alpha, nu=nu, H=H, v2=True, semantics=S.ARS_AGENT, seed=seed,
                      initial_policy=W0, distributed=False)
    mean0 = rng.normal(size=no) * 0.05
    mean0[2::2] += np.pi / 2
    sig0 = rng.uniform(0.5, 2.0, no)
    eng.mean.copy_(_cuda(mean0))
    eng.inv_sigma.copy_(_cuda(1.0 / sig0))
    ret = eng.run_iteration(update=False).cpu().numpy().copy()
    assert ret.shape == (2 * N,) and np.isfinite(ret).all()
    deltas = S.ops.philox_deltas(seed, 0, 0, N, ws).cpu().numpy()
    ks = np.concatenate([[0, 1, N - 1], rng.integers(0, N, 9)])
    for k in ks:
        np.testing.assert_array_equal(deltas[k], O.philox_delta(seed, 0, int(k), ws))
        for j, sign in enumerate((+1, -1)):
            want = O.rollout(po, O.GYM, H, policy=W0 + sign * nu * deltas[k], mean=mean0, inv_sigma=1.0 / sig0)[0]
            assert _close(ret[2 * k + j], want), (k, sign, ret[2 * k + j], want)
    # ranking is bit-exact given identical rewards; the update follows ars_agent.py:110-130
    eng.apply_update()
    order = eng.order.cpu().numpy()
    np.testing.assert_array_equal(order, O.sort_directions(ret))
    W1, _ = O.update_policy(W0, deltas, ret, b=N, alpha=alpha, semantics=0)
    np.testing.assert_allclose(eng.W.cpu().numpy(), W1, rtol=1e-9, atol=1e-12)
    # V2 statistics over all 2,048,000 visited states: against the oracle on a 64-rollout slice run
    # through the same kernel path (moments are sums, so slices add up), and against the oracle's own
    # trajectories for that slice
    cnt = float(eng.stats.cpu()[0])
    assert cnt == 2.0 * N * H
    sl = S.ops.rollout(ps, H, B=64, base_policy=_cuda(W0), nu=nu, seed=seed, iteration=0, dir0=0,
                       mean=_cuda(mean0), inv_sigma=_cuda(1.0 / sig0), stats_pivot=eng.pivot)
    rec = S.ops.stats_finalize(sl.stats_partial, sl.samples, eng.pivot).cpu().numpy()
    trajs = [O.rollout(po, O.GYM, H, policy=W0 + s * nu * deltas[k], mean=mean0, inv_sigma=1.0 / sig0,
                       want_traj=True)[2] for k in range(32) for s in (+1, -1)]
    m, v = O.mean_var(np.concatenate(trajs))
    np.testing.assert_allclose(rec[1:1 + no], m, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rec[1 + no:] / (rec[0] - 1), v, rtol=1e-8)


def test_config4_reward_constraint_screening(S, O):
    """config[3]: n=3, 256 directions, real (l=.8, m=1.2, k=10.2) vs simulator = real + eps*u/|u|:
    phase 1 simulator rollouts of all +-delta, mask, phase 2 real rollouts of the survivors."""
    n, N, H, nu, seed, eps = 3, 256, 1000, 0.05, 3, 1e-3
    real = dict(l_i=0.8, m_i=1.2, k=10.2)
    u = np.array([1.0, -2.0, 0.5])
    u = eps * u / np.linalg.norm(u)
    sim = dict(m_i=real["m_i"] + u[0], l_i=real["l_i"] + u[1], k=real["k"] + u[2])
    ps, psim = S.make_params(n=n, **real), S.make_params(n=n, **sim)
    po, posim = O.make_params(n=n, **real), O.make_params(n=n, **sim)
    ws = (n - 1) * (2 * n + 2)
    rng = np.random.default_rng(8)
    W0 = rng.uniform(-1, 1, ws) * 0.5
    probe = S.ArsEngine(psim, N=N, b=N, alpha=0.0075, nu=nu, H=H, seed=seed, initial_policy=W0, distributed=False)
    pr = probe.run_iteration(update=False).cpu().numpy()
    thr = float(np.median(np.minimum(pr[0::2], pr[1::2])))  # screens out about half of the directions
    eng = S.ArsEngine(ps, N=N, b=N, alpha=0.0075, nu=nu, H=H, semantics=S.ARS_AGENT, seed=seed,
                      initial_policy=W0, distributed=False, sim_params=psim, sim_threshold=thr)
    ret = eng.run_iteration(update=False).cpu().numpy().copy()
    mask = eng.mask.cpu().numpy().astype(bool)
    sim_ret = eng.sim_returns.cpu().numpy().reshape(N, 2)
    np.testing.assert_array_equal(mask, (sim_ret[:, 0] > thr) & (sim_ret[:, 1] > thr))  # ars_agent.py:150-157
    assert 0 < mask.sum() < N, "threshold should screen out some but not all directions"
    assert np.isnan(ret.reshape(N, 2)[~mask]).all() and np.isfinite(ret.reshape(N, 2)[mask]).all()
    deltas = S.ops.philox_deltas(seed, 0, 0, N, ws).cpu().numpy()
    for k in np.concatenate([[0, N - 1], rng.integers(0, N, 8)]):
        for j, sign in enumerate((+1, -1)):
            pol = W0 + sign * nu * deltas[k]
            assert _close(sim_ret[k, j], O.rollout(posim, O.GYM, H, policy=pol)[0])
            if mask[k]:
                assert _close(ret[2 * k + j], O.rollout(po, O.GYM, H, policy=pol)[0])
    # survivors-only update: equals the oracle's update over the compacted triples
    eng.apply_update()
    keep = np.nonzero(mask)[0]
    rk = ret.reshape(N, 2)[keep].reshape(-1)
    W1, _ = O.update_policy(W0, deltas[keep], rk, b=N, alpha=0.0075, semantics=0)
    np.testing.assert_allclose(eng.W.cpu().numpy(), W1, rtol=1e-9, atol=1e-12)


def test_config5_per_gpu_share_grouped_rollouts(S, O):
    """config[4]: n=10, 128 rollouts per policy from perturbed starts; the per-GPU share at 8 GPUs is
    512 directions x 2 x 128 = 131,072 envs.  Sampled envs against the oracle; sharding the same
    directions over two launches (dir0 offset) reproduces the one-launch result bit for bit."""
    n, D, R, H, nu, seed, it = 10, 512, 128, 1000, 0.01, 21, 2
    ps, po = S.make_params(n=n), O.make_params(n=n)
    no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
    rng = np.random.default_rng(3)
    W0 = rng.uniform(-1, 1, ws) * 0.05
    B = 2 * D * R
    kw = dict(base_policy=_cuda(W0), nu=nu, seed=seed, iteration=it, rollouts_per_policy=R, init_perturb=1e-2)
    full = S.ops.rollout(ps, H, B=B, dir0=0, **kw).returns
    half = B // 2
    lo = S.ops.rollout(ps, H, B=half, dir0=0, **kw).returns
    hi = S.ops.rollout(ps, H, B=half, dir0=D // 2, **kw).returns
    assert torch.equal(full[:half], lo) and torch.equal(full[half:], hi)
    ret = full.cpu().numpy().reshape(D, 2, R)
    assert np.isfinite(ret).all()
    start = np.zeros(no)
    start[2::2] = np.pi / 2
    for k, j, r in ((0, 0, 0), (0, 1, 5), (D - 1, 1, R - 1), (D // 2, 0, 64), (17, 1, 33)):
        d = O.philox_delta(seed, it, k, ws)
        pert = O.philox_delta(seed, it, r, no, dist=1, stream=1)  # U[0,1) start perturbation of rollout r
        want = O.rollout(po, O.GYM, H, policy=W0 + (1, -1)[j] * nu * d, init_state=start + 1e-2 * pert)[0]
        assert _close(ret[k, j, r], want), (k, j, r, ret[k, j, r], want)
    red = S.ops.reduce_returns(full, R).cpu().numpy()
    np.testing.assert_allclose(red, ret.reshape(2 * D, R).mean(1), rtol=1e-13)


def test_config5_full_million_envs_engine_iteration(S, O):
    """config[4] at its FULL size on one GPU: 4,096 directions x 2 x 128 rollouts = 1,048,576 ten-segment
    environments, one ARS V2 engine iteration (graph-free): sampled environments against the oracle, the
    per-direction means, the count of the V2 statistics, ranking and update against the oracle's update on the
    engine's own returns."""
    n, N, R, H, nu, alpha, seed = 10, 4096, 128, 1000, 0.01, 0.0075, 4
    ps, po = S.make_params(n=n), O.make_params(n=n)
    no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
    rng = np.random.default_rng(6)
    W0 = rng.uniform(-1, 1, ws) * 0.05
    eng = S.ArsEngine(ps, N=N, b=N, alpha=alpha, nu=nu, H=H, v2=True, semantics=S.ARS_AGENT, seed=seed,
                      rollouts_per_direction=R, init_perturb=1e-2, initial_policy=W0, distributed=False)
    ret = eng.run_iteration(update=False).cpu().numpy().copy()      # [2N] per-direction means
    per_env = eng.last.returns.cpu().numpy().reshape(N, 2, R)
    assert per_env.size == 1048576 and np.isfinite(per_env).all()
    np.testing.assert_allclose(ret, per_env.reshape(2 * N, R).mean(1), rtol=1e-13)
    start = np.zeros(no)
    start[2::2] = np.pi / 2
    for k, j, r in ((0, 0, 0), (N - 1, 1, R - 1), (2047, 0, 64), (1234, 1, 17), (4000, 0, 99)):
        d = O.philox_delta(seed, 0, k, ws)
        pert = O.philox_delta(seed, 0, r, no, dist=1, stream=1)
        want = O.rollout(po, O.GYM, H, policy=W0 + (1, -1)[j] * nu * d, init_state=start + 1e-2 * pert)[0]
        assert _close(per_env[k, j, r], want), (k, j, r, per_env[k, j, r], want)
    eng.apply_update()
    assert float(eng.stats.cpu()[0]) == 2.0 * N * R * H
    np.testing.assert_array_equal(eng.order.cpu().numpy(), O.sort_directions(ret))
    deltas = S.ops.philox_deltas(seed, 0, 0, N, ws).cpu().numpy()
    W1, _ = O.update_policy(W0, deltas, ret, b=N, alpha=alpha, semantics=0)
    np.testing.assert_allclose(eng.W.cpu().numpy(), W1, rtol=1e-9, atol=1e-12)
    assert bool(torch.isfinite(eng.inv_sigma).all())

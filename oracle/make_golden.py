"""TEST INFRASTRUCTURE ONLY: generates tests/golden/*.npz|json by running the UNMODIFIED
reference (Python files through oracle/ref_harness.py, C++ swimmer through oracle/_ref) in
the build container.  The fixtures are committed; this script is the record of how they were
made.  Run:  python oracle/make_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as O  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rand_state(rng, n, scale=3.0):
    th = rng.uniform(-4, 4, n)
    thd = rng.normal(size=n) * scale
    return np.concatenate([rng.normal(size=2), np.stack([th, thd], 1).ravel()])


class RandRecorder:
    """Wraps np.random.rand to record every draw the reference makes (ars_agent.py:137,
    safe_ars/ars.py:84) so tests can replay the same perturbations through the kernels."""

    def __init__(self):
        self.draws = []
        self._orig = np.random.rand

    def __enter__(self):
        def rec(*shape):
            x = self._orig(*shape)
            self.draws.append(np.array(x, copy=True))
            return x
        np.random.rand = rec
        return self

    def __exit__(self, *a):
        np.random.rand = self._orig


def gym_steps(ref):
    rng = np.random.default_rng(20260101)
    out = {}
    for n in (2, 3, 5, 10):
        T = 24
        par = np.zeros((T, 4))
        st = np.zeros((T, 2 * n + 2)); ac = np.zeros((T, n - 1))
        nx = np.zeros_like(st); rw = np.zeros(T); acc = np.zeros((T, n + 2))
        for t in range(T):
            l, m, k, h = rng.uniform(.5, 1.5), rng.uniform(.5, 1.5), rng.uniform(5, 15), 10 ** rng.uniform(-3.3, -2)
            if t < 4:
                l, m, k, h = 1., 1., 10., 1e-3
            env = ref.SwimmerEnv(n=n, l_i=l, m_i=m, k=k, h=h, direction=[1., 0.] if t % 2 == 0 else [0.6, -0.8])
            s = rand_state(rng, n) if t else np.array(env.reset())
            a = rng.uniform(-5, 5, n - 1)
            if t == 0:
                a = np.full(n - 1, 2.5)  # SURVEY App. A single-step KAT
            env.set_state(s.tolist())
            g, thdd = env.compute_accelerations(a, env.G_dot, env.theta, env.theta_dot)
            ob, r, done, info = env.step(a)
            par[t] = (l, m, k, h); st[t] = s; ac[t] = a; nx[t] = ob; rw[t] = r
            acc[t] = np.concatenate([g, thdd])
        out.update({f"n{n}_params": par, f"n{n}_state": st, f"n{n}_action": ac, f"n{n}_next": nx,
                    f"n{n}_reward": rw, f"n{n}_acc": acc})
    np.savez(os.path.join(OUT, "gym_step.npz"), **out)


def gym_rollouts(ref):
    rng = np.random.default_rng(7)
    out = {}
    for n, H in ((3, 1000), (5, 1000), (10, 300)):
        ep = ref.EnvParam("x", n=n, H=H, l_i=.8, m_i=1.2, h=1e-3, k=10.2, epsilon=0)
        E = ref.Environment(ep)
        no, na = 2 * n + 2, n - 1
        W = rng.uniform(-1, 1, (3, na, no)) * 0.4
        mean = rng.normal(size=no) * 0.2
        var = rng.uniform(0.3, 3.0, no)
        r1 = np.zeros(3); r2 = np.zeros(3)
        f1 = np.zeros((3, no)); f2 = np.zeros((3, no))
        tr1 = np.zeros((3, H // 50, no)); tr2 = np.zeros((3, H // 50, no))
        for i in range(3):
            R, S = E.rollout(W[i])
            r1[i] = R; f1[i] = S[-1]; tr1[i] = np.array(S)[49::50]
            R, S = E.rollout(W[i], covariance=np.diag(var), mean=mean)
            r2[i] = R; f2[i] = S[-1]; tr2[i] = np.array(S)[49::50]
        out.update({f"n{n}_W": W, f"n{n}_mean": mean, f"n{n}_var": var, f"n{n}_H": H,
                    f"n{n}_v1_return": r1, f"n{n}_v1_final": f1, f"n{n}_v1_traj50": tr1,
                    f"n{n}_v2_return": r2, f"n{n}_v2_final": f2, f"n{n}_v2_traj50": tr2})
    np.savez(os.path.join(OUT, "gym_rollout.npz"), **out)


def ars_agent_runs(ref):
    out = {}
    # config 1 of BASELINE.json: n=3, V1, N=8, b=8, H=1000, alpha=.0075, nu=.01, seed 0
    ep = ref.EnvParam("LeonSwimmer-RealWorld", n=3, H=1000, l_i=1., m_i=1., h=1e-3, k=10., epsilon=0)
    ap = ref.ARSParam("RLControl", V1=True, n_iter=2, H=1000, N=8, b=8, alpha=0.0075, nu=0.01,
                      safe=False, threshold=0, initial_w="Zero")
    ag = ref.ARSAgent(ep, ap, seed=0)
    with RandRecorder() as rec:
        rets, pols = [], []
        for it in range(3):
            rets.append(ag.runOneIteration()); pols.append(ag.policy.copy())
    out["c1_deltas"] = 2 * np.array(rec.draws).reshape(3, 8, 2, 8) - 1
    out["c1_rand"] = np.array(rec.draws).reshape(3, 8, 2, 8)
    out["c1_returns"] = np.array(rets); out["c1_policies"] = np.array(pols)
    # V2, n=3, N=4, b=4, H=250, 3 iterations
    ep = ref.EnvParam("x", n=3, H=250, l_i=.8, m_i=1.2, h=1e-3, k=10.2, epsilon=0)
    ap = ref.ARSParam("x", V1=False, n_iter=2, H=250, N=4, b=4, alpha=0.0075, nu=0.01,
                      safe=False, threshold=0, initial_w="Zero")
    ag = ref.ARSAgent(ep, ap, seed=3)
    with RandRecorder() as rec:
        rets, pols, means, covs = [], [], [], []
        for it in range(3):
            rets.append(ag.runOneIteration()); pols.append(ag.policy.copy())
            means.append(ag.mean.copy()); covs.append(np.diag(ag.covariance).copy())
    out["v2_rand"] = np.array(rec.draws).reshape(3, 4, 2, 8)
    out["v2_returns"] = np.array(rets); out["v2_policies"] = np.array(pols)
    out["v2_means"] = np.array(means); out["v2_vars"] = np.array(covs)
    # runTraining curve (V1, N=2, b=2, H=200, n_iter=4, seed 11): length n_iter+1
    ep = ref.EnvParam("x", n=3, H=200, l_i=1., m_i=1., h=1e-3, k=10., epsilon=0)
    ap = ref.ARSParam("x", V1=True, n_iter=4, H=200, N=2, b=2, alpha=0.02, nu=0.05,
                      safe=False, threshold=0, initial_w="Zero")
    ag = ref.ARSAgent(ep, ap, seed=11)
    with RandRecorder() as rec:
        curve = ag.runTraining()
    out["rt_rand"] = np.array(rec.draws).reshape(5, 2, 2, 8)
    out["rt_curve"] = curve; out["rt_policy"] = ag.policy.copy()
    np.savez(os.path.join(OUT, "ars_agent.npz"), **out)


def ars_agent_safe(ref):
    """Reward-constraint safe mode (ars_agent.py:144-157), N=1 (the only N the reference's scripts
    use, and the only N for which its bookkeeping is well defined -- SURVEY App. D-2)."""
    out = {}
    rng = np.random.default_rng(5)
    W0 = rng.uniform(-1, 1, (2, 8)) * 0.3
    with tempfile.TemporaryDirectory() as td:
        wpath = os.path.join(td, "w0.npy"); np.save(wpath, W0)
        dpath = os.path.join(td, "db.npz")
        np.savez(dpath, policies=[W0], trajectories=[np.zeros((50, 8))])
        ep = ref.EnvParam("real", n=3, H=300, l_i=.8, m_i=1.2, h=1e-3, k=10.2, epsilon=0.001)
        E = ref.Environment(ep)
        base_ret, _ = E.rollout(W0)
        # a threshold that screens out about half of the perturbations (|r+- - base| ~ 0.04..0.17)
        thr = base_ret - 0.09 - ref.Threshold(K=1, A=0.1, B=0.001).compute_alpha(300) * 0.001
        ap = ref.ARSParam("x", V1=True, n_iter=7, H=300, N=1, b=1, alpha=0.0075, nu=0.01,
                          safe=True, threshold=thr, initial_w=wpath)
        st = ref.Threshold(K=1, A=0.1, B=0.001)
        np.random.seed(99)  # the eps-perturbation is drawn before the agent seeds (ars_agent.py:52 vs :95)
        ag = ref.ARSAgent(ep, ap, data_path=dpath, seed=4, approx_error=0.001, sim_thresh=st)
        est = ag.estimated_param
        out["sim_params"] = np.array([est.l_i, est.m_i, est.k])
        out["real_params"] = np.array([.8, 1.2, 10.2])  # (aliased and mutated by the reference!)
        out["sim_threshold"] = ag.sim_threshold; out["threshold"] = thr; out["W0"] = W0
        out["alpha_H"] = st.compute_alpha(300)
        # NOTE reference quirk D-5: estimated_param aliases real_env_param, so after construction the
        # "real world" Environment was already built with the unperturbed values while
        # estimated_param carries the perturbed ones.
        with RandRecorder() as rec:
            rets, pols = [], []
            for it in range(8):
                r = ag.runOneIteration()
                rets.append(np.array(r + [np.nan] * (2 - len(r)))); pols.append(ag.policy.copy())
        out["rand"] = np.array(rec.draws).reshape(8, 2, 8)
        out["returns"] = np.array(rets); out["policies"] = np.array(pols)
        # simulator returns for each iteration's +/- policies (recomputed, for the screening mask)
    np.savez(os.path.join(OUT, "ars_agent_safe.npz"), **out)


def safe_ars_runs(ref):
    out = {}
    n = 3
    real = ref.SwimmerEnv("RealWorld", n=n, m_i=1., l_i=1., k=10.)
    np.random.seed(0)
    d = np.random.rand(3)
    th_sim = np.array([1., 1., 10.]) + d / np.linalg.norm(d) * 0.05
    sim = ref.SwimmerEnv("Simulator", n=n, m_i=th_sim[0], l_i=th_sim[1], k=th_sim[2])
    cost = lambda x: np.max([abs(x[3 + 2 * i]) for i in range(n)])  # safe_ars/experiment.py:44
    out["sim_mlk"] = th_sim
    # Basic_ARS.train
    ag = ref.Basic_ARS()
    np.random.seed(1)
    with RandRecorder() as rec:
        curve, states = ag.train(3, real, 4, 2, 0.02, 0.05, 200)
    out["basic_rand"] = np.array(rec.draws).reshape(3, 4, 2, 8)
    out["basic_curve"] = curve; out["basic_policy"] = ag.policy.copy()
    out["basic_states_last"] = states[-1][-1]
    # Safe_ARS.rollout on fixed policies with thresholds that freeze some rollouts mid-way
    rng = np.random.default_rng(3)
    Ws = rng.uniform(-1, 1, (6, 2, 8)) * 1.5
    thrs = np.array([6.0, 12.0, 6.0, 8.5, 30.0, 100.0])  # last one never freezes
    rets = np.zeros(6); finals = np.zeros((6, 8)); frozen = np.zeros(6, dtype=np.int64)
    import io, contextlib
    for i in range(6):
        sag = ref.Safe_ARS(cost, thrs[i], thrs[i] - 0.2, sim)
        with contextlib.redirect_stdout(io.StringIO()):
            R, S = sag.rollout(real, Ws[i], 400)
        S = np.array(S)
        rets[i] = R; finals[i] = S[-1]
        same = np.all(S[1:] == S[:-1], axis=1)
        fr = 400
        for t in range(len(S) - 1, 0, -1):
            if same[t - 1]:
                fr = t
            else:
                break
        frozen[i] = fr
    out["safe_W"] = Ws; out["safe_returns"] = rets; out["safe_finals"] = finals
    out["safe_frozen_from"] = frozen; out["safe_real_thresh"] = thrs; out["safe_sim_thresh"] = thrs - 0.2
    # Safe_ARS.train (2 iterations)
    sag = ref.Safe_ARS(cost, 3.0, 2.0, sim)
    np.random.seed(2)
    with RandRecorder() as rec, contextlib.redirect_stdout(io.StringIO()):
        curve, states = sag.train(2, real, 3, 2, 0.02, 0.3, 150)
    out["safetrain_rand"] = np.array(rec.draws).reshape(2, 3, 2, 8)
    out["safetrain_curve"] = curve; out["safetrain_policy"] = sag.policy.copy()
    np.savez(os.path.join(OUT, "safe_ars.npz"), **out)


def topb(ref):
    rng = np.random.default_rng(12)
    ag = ref.Basic_ARS()
    out = {}
    for N in (1, 2, 8, 33, 256, 1024):
        r = rng.normal(size=2 * N) * 50
        out[f"returns_{N}"] = r
        out[f"order_{N}"] = np.array(ag.sort_directions([None] * N, r.tolist()), dtype=np.int64)
    # ties and NaNs: the declared rule is np.argsort(kind='stable')[::-1]
    r = np.array([1., 0., 1., -1., 0.5, 1., 1., 1., np.nan, 0., 2., 2., 0., np.nan])
    mx = [max(r[2 * i], r[2 * i + 1]) for i in range(len(r) // 2)]
    out["returns_ties"] = r
    out["order_ties"] = np.argsort(mx, kind="stable")[::-1].astype(np.int64)
    np.savez(os.path.join(OUT, "topb.npz"), **out)


def rlglue(ref):
    """Golden numbers printed in the reference's own test logs + full-precision outputs of the
    compiled reference C++ (oracle/_ref) on random states."""
    gold = {
        "source": "rlglue/test/acceleration-compare.txt:24,102-103; rlglue/test/swimmer-compare.txt:100",
        "params": {"n": 3, "l_i": 1.0, "m_i": 1.0, "k": 10.0, "max_u": 5.0, "h_inferred": 0.003},
        "state": [-0.0453422, 1.33766e-11, -1.35003, -1.4868, 1.5708, -1.88179e-15, -1.79156, 1.4868],
        "torque": [2.5, 2.5],
        "G_dotdot_6digits": [-2.05622, -0.0173465],
        "theta_dotdot_6digits": [3.97741, 14.2667, 19.7878],
        "state_after_update_6digits": [-0.0515109, -5.20396e-05, -1.35445, -1.47487, 1.57093,
                                       0.0428002, -1.78692, 1.54616],
        "coulom_barycenter_acc_6digits": [0.284343, -8.38483e-11],
    }
    with open(os.path.join(OUT, "rlglue_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    O.build_ref()
    assert O.ref_cpp() is not None
    rng = np.random.default_rng(77)
    out = {}
    for n in (2, 3, 5, 10):
        T = 16
        par = np.zeros((T, 4)); st = np.zeros((T, 2 * n + 2)); ac = np.zeros((T, n - 1))
        nx = np.zeros_like(st); acc = np.zeros((T, n + 2))
        for t in range(T):
            l, m, k, h = rng.uniform(.5, 1.5), rng.uniform(.5, 1.5), rng.uniform(5, 15), 10 ** rng.uniform(-3, -2)
            if t < 3:
                l, m, k, h = 1., 1., 10., 0.01
            p = O.make_params(n=n, l_i=l, m_i=m, k=k, h=h)
            O.ref_cpp_set_params(p)
            s = rand_state(rng, n) if t else np.full(2 * n + 2, 0.001)  # env_start state cpp:39-42
            a = rng.uniform(-5, 5, n - 1)
            g, thdd = O.ref_cpp_accelerations(s, a, n)
            par[t] = (l, m, k, h); st[t] = s; ac[t] = a; nx[t] = O.ref_cpp_step(s, a)
            acc[t] = np.concatenate([g, thdd])
        out.update({f"n{n}_params": par, f"n{n}_state": st, f"n{n}_action": ac, f"n{n}_next": nx,
                    f"n{n}_acc": acc})
    # 500-step fixed-torque trajectory from the env_start state, n=3, parameters.txt values
    p = O.make_params(n=3, l_i=1., m_i=1., k=10., h=0.01)
    O.ref_cpp_set_params(p)
    s = np.full(8, 0.001); a = np.array([1.5, -2.0]); tot = 0.0
    for t in range(500):
        s = O.ref_cpp_step(s, a); tot += s[0]
    out["roll_action"] = a; out["roll_final"] = s; out["roll_return"] = tot
    np.savez(os.path.join(OUT, "rlglue_step.npz"), **out)


def rlglue_agent(ref):
    """The UNMODIFIED rlglue/agent/SwimmerAgent.py driven against the UNMODIFIED compiled C++ swimmer the way
    rlglue/experiment/SwimmerExperiment.cpp:65-100 drives them (oracle/rlglue_protocol.py emulates the RL-Glue
    runtime in process).  Case a: today's rlglue/parameters.txt (N = b = 1, H = 1000); case b: several
    directions, b < N, five segments."""
    from oracle import rlglue_protocol as RP
    out = {}
    base = dict(n_seg=3, direction=(1.0, 0.0), h_global=0.01, N=1, b=1, H=1000, alpha=0.02, nu=0.02, max_u=5.,
                l_i=1., k=10., m_i=1.)
    for tag, par, n_it, seed in (("a", base, 6, 5), ("b", dict(base, n_seg=5, N=3, b=2, H=150, alpha=0.05, nu=0.3), 4, 9)):
        r = RP.run_reference_protocol(par, n_it, seed)
        out[tag + "_par"] = np.array([par["n_seg"], par["N"], par["b"], par["H"], par["alpha"], par["nu"], par["max_u"],
                                      par["l_i"], par["k"], par["m_i"], par["h_global"]])
        for k, v in r.items():
            out[tag + "_" + k] = v
    np.savez(os.path.join(OUT, "rlglue_agent.npz"), **out)


def misc(ref):
    out = {}
    vals = []
    for K, A, B, H in ((1, 0.1, 0.001, 1000), (1, 1, 0.5, 10), (2, 0.3, 0.9, 1000), (1, 0.7, 1e-6, 300)):
        vals.append((K, A, B, H, ref.Threshold(K=K, A=A, B=B).compute_alpha(H)))
    out["threshold_alpha"] = np.array(vals)
    np.savez(os.path.join(OUT, "misc.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_harness.load()
    gym_steps(ref)
    gym_rollouts(ref)
    topb(ref)
    misc(ref)
    rlglue(ref)
    rlglue_agent(ref)
    safe_ars_runs(ref)
    ars_agent_safe(ref)
    ars_agent_runs(ref)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

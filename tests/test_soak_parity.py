"""Randomised soak: many short rollouts with random segment counts, physics, policy scales, normalisation
and start states against the CPU oracle (returns 1e-6 relative, final states 1e-8 where the dynamics stay in
explicit Euler's stable regime)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _soak(S, O, cases, seed):
    rng = np.random.default_rng(seed)
    worst_r, worst_f, skipped = 0.0, 0.0, 0
    for case in range(cases):
        n = int(rng.integers(2, 11))
        kw = dict(n=n, l_i=float(rng.uniform(0.4, 2.5)), m_i=float(rng.uniform(0.3, 4.0)), k=float(rng.uniform(0.0, 40.0)),
                  h=float(10 ** rng.uniform(-3.5, -2.5)))
        ps, po = S.make_params(**kw), O.make_params(**kw)
        no = 2 * n + 2
        H = int(rng.integers(1, 260))
        P = 6
        Ws = rng.uniform(-1, 1, (P, n - 1, no)) * 10 ** rng.uniform(-2, 0.3)
        v2 = bool(rng.integers(0, 2))
        mean = rng.normal(size=no) * 0.1 if v2 else None
        inv = rng.uniform(0.3, 3.0, no) if v2 else None
        init = rng.normal(size=no) * rng.uniform(0, 3) if rng.integers(0, 2) else None
        res = S.ops.rollout(ps, H, policies=torch.as_tensor(Ws).cuda(),
                            mean=None if mean is None else torch.as_tensor(mean).cuda(),
                            inv_sigma=None if inv is None else torch.as_tensor(inv).cuda(),
                            init_state=None if init is None else torch.as_tensor(init[None]).cuda(), want_final=True)
        gr, gf = res.returns.cpu().numpy(), res.final_state.cpu().numpy()
        for q in range(P):
            r, f, _ = O.rollout(po, O.GYM, H, policy=Ws[q], mean=mean, inv_sigma=inv, init_state=init)
            if not np.isfinite(f).all() or np.abs(f[3::2]).max() > 60.0:
                skipped += 1  # explicit Euler has left its stable regime: round-off is amplified exponentially
                assert np.isfinite(gf[q]).all() == np.isfinite(f).all() or not np.isfinite(f).all()
                continue
            er = abs(gr[q] - r) / max(1e-3, abs(r))
            ef = np.max(np.abs(gf[q] - f)) / max(1.0, np.max(np.abs(f)))
            worst_r, worst_f = max(worst_r, er), max(worst_f, ef)
            assert er < 1e-6 and ef < 1e-8, (case, kw, H, v2, er, ef)
    return cases, skipped, worst_r, worst_f


@pytest.mark.parametrize("seed", [0, 1])
def test_randomised_rollout_soak(S, O, seed):
    cases, skipped, worst_r, worst_f = _soak(S, O, 150, seed)
    assert skipped < cases * 6 * 0.6  # most cases stay in the stable regime and are really compared
    assert worst_r < 1e-6 and worst_f < 1e-8

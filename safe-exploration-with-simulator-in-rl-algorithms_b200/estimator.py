"""Parameter estimation objective with the reference's interface (ars/estimator.py:16-120).

`Estimator.I(x)` -- the one-step-prediction error of a candidate (m_i, l_i, k) against recorded
real trajectories -- is H-1 *independent* single steps per trajectory, so it maps onto one
batched policy product + one batched step (swm_policy_actions, swm_step_batched) per trajectory
instead of H-1 Python iterations.  CMA-ES itself is a third-party host optimiser
(`cma`, un-vendored and unpinned in the reference): it stays on the host and is only imported
when `estimate_real_env_param` is called.
"""
import dataclasses

import numpy as np
import torch

from . import _lib, ops
from .parameters import EnvParam


class Estimator:
    def __init__(self, database, guess_param, capacity, unknowns=("m_i", "l_i", "k"), device=None):
        assert database.size > 0, "Database is empty"
        assert len(database.trajectories[0]) == guess_param.H, "Rollouts are not the same"
        self.guess_param = guess_param
        self.unknowns = unknowns
        self.database = database
        self.subset = np.random.randint(0, self.database.size, capacity)
        self.iter = 0
        _lib.require_cuda()
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self._cache = {}

    def _traj(self, k):
        if k not in self._cache:
            t = torch.as_tensor(np.asarray(self.database.trajectories[k], dtype=np.float64))
            p = torch.as_tensor(np.asarray(self.database.policies[k], dtype=np.float64))
            self._cache[k] = (t.to(self.device).contiguous(), p.to(self.device).contiguous())
        return self._cache[k]

    def convert_to_env_param(self, x):
        d = dataclasses.asdict(self.guess_param)
        for i, name in enumerate(self.unknowns):
            d[name] = x[i]
        return EnvParam(**d)

    def I(self, x):
        """sum_k sum_t || step_x(s_t, W_k s_t) - s_{t+1} ||_2 (ars/estimator.py:36-62)."""
        p = self.convert_to_env_param(x)
        params = _lib.make_params(n=p.n, l_i=p.l_i, m_i=p.m_i, k=p.k, h=p.h)
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        for k in self.subset:
            traj, policy = self._traj(int(k))
            s = traj[:-1].contiguous()
            act = ops.policy_actions(params, s, policy.reshape(1, -1), rollouts_per_policy=s.shape[0])
            nxt, _ = ops.step_batched(params, s, act, want_reward=False)
            total = total + torch.linalg.norm(nxt - traj[1:], dim=1).sum()
        return float(total.cpu())

    def I_population(self, xs):
        """Objective for a whole CMA-ES generation (list of candidates) -> list of floats: for every
        recorded trajectory ONE launch steps all candidates' models from the same states
        (swm_step_batched_models); the actions W_k s_t do not depend on the candidate."""
        cands = [self.convert_to_env_param(x) for x in xs]
        plist = [_lib.make_params(n=p.n, l_i=p.l_i, m_i=p.m_i, k=p.k, h=p.h) for p in cands]
        M = len(plist)
        total = torch.zeros(M, dtype=torch.float64, device=self.device)
        for k in self.subset:
            traj, policy = self._traj(int(k))
            s = traj[:-1].contiguous()
            act = ops.policy_actions(plist[0], s, policy.reshape(1, -1), rollouts_per_policy=s.shape[0])
            nxt, _ = ops.step_batched_models(plist, s.unsqueeze(0).expand(M, -1, -1).contiguous(),
                                             act.unsqueeze(0).expand(M, -1, -1).contiguous())
            total = total + torch.linalg.norm(nxt - traj[1:].unsqueeze(0), dim=2).sum(dim=1)
        return total.cpu().tolist()

    def estimate_real_env_param(self):
        """CMA-ES over I (ars/estimator.py:89-110); needs the third-party `cma` package."""
        try:
            import cma
        except ImportError as e:
            raise ImportError("estimate_real_env_param needs the `cma` package (host-side CMA-ES, "
                              "ars/estimator.py:103); it is not part of this library") from e
        d = dataclasses.asdict(self.guess_param)
        x0 = np.array([d[u] for u in self.unknowns], dtype=np.float64)
        es = cma.CMAEvolutionStrategy(x0, 1).optimize(self.I)
        est_x, _, _ = es.best.get()
        return self.convert_to_env_param(est_x)

"""Rollout driver with the reference's interface (ars/environment.py:10-60): a linear policy
(ARS V1) or a normalised linear policy (ARS V2) run for H steps on a SwimmerEnv.  The H-step
loop, the policy product and the physics are one fused kernel launch."""
import numpy as np
import torch

from . import _lib, ops
from .swimmer_env import SwimmerEnv


def _inv_sigma(covariance):
    """diag(cov) ** (-1/2) as in ars/environment.py:32; accepts a matrix or its diagonal."""
    c = np.asarray(covariance, dtype=np.float64)
    d = np.diag(c) if c.ndim == 2 else c
    return d ** (-1 / 2)


class Environment:
    def __init__(self, env_param, device=None, variant="gym"):
        self.env_param = env_param
        self.env = SwimmerEnv(envName=env_param.name, n=env_param.n, l_i=env_param.l_i,
                              m_i=env_param.m_i, h=env_param.h, k=env_param.k, variant=variant,
                              device=device)

    def _t(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.env._dev())

    def select_action(self, policy, observation, covariance=None, mean=None):
        """W obs (V1) or (W diag(cov)^-1/2)(obs - mean) (V2) -- ars/environment.py:19-35."""
        obs = self._t(observation).reshape(1, -1)
        pol = self._t(policy).reshape(1, -1)
        v2 = covariance is not None and mean is not None
        act = ops.policy_actions(self.env.params(), obs, pol,
                                 mean=self._t(mean) if v2 else None,
                                 inv_sigma=self._t(_inv_sigma(covariance)) if v2 else None)
        return act[0].cpu().numpy()

    def rollout(self, policy, covariance=None, mean=None):
        """-> (total_reward, saved_states[H][2n+2]) like ars/environment.py:37-57."""
        res = self.rollout_batched(np.asarray(policy)[None], covariance, mean, want_trajectory=True)
        total = float(res.returns.cpu()[0])
        states = res.trajectory[:, 0, :].cpu().numpy().tolist()
        return total, states

    def rollout_batched(self, policies, covariance=None, mean=None, want_trajectory=False,
                        want_final=False, rollouts_per_policy=1, init_state=None):
        """policies[P, n-1, 2n+2] -> ops.RolloutResult (device tensors), one env per policy."""
        v2 = covariance is not None and mean is not None
        return ops.rollout(self.env.params(), self.env_param.H, variant=self.env.variant,
                           policies=self._t(policies),
                           mean=self._t(mean) if v2 else None,
                           inv_sigma=self._t(_inv_sigma(covariance)) if v2 else None,
                           rollouts_per_policy=rollouts_per_policy, init_state=init_state,
                           want_trajectory=want_trajectory, want_final=want_final)

    def close(self):
        self.env.close()

"""ARS V1 with the RL-Glue agent's semantics, at rollout granularity.

The reference spreads this loop over four processes talking over sockets: the experiment driver
(rlglue/experiment/SwimmerExperiment.cpp:65-100) issues 2*N*H training `RL_step`s per iteration,
re-loading the saved start state every H steps, then freezes the agent and runs one H-step evaluation
rollout whose return is the line written to plot/results.txt; the agent
(rlglue/agent/SwimmerAgent.py:79-241) draws delta ~ U[0,1) (:203-212; the docstring says normal),
clips actions to +-max_u (:181-201), does NOT sort the directions (:214-221 is a TODO returning the
identity), uses the first b of them with the *sample* standard deviation (:223-241), and the
environment (rlglue/environment/SwimmerEnvironment.cpp) integrates semi-implicitly from the start
state 0.001 (:39-42).

Here one iteration is: one fused launch of the 2N training rollouts (dynamics variant RLGLUE, clipped
actions, U[0,1) Philox perturbations), the index-order / sample-std update, and one evaluation
rollout of the updated policy.  Parity note: the RL-Glue runtime and its Python codec are not
available, so the four-process reference cannot be run; the arithmetic pieces (dynamics,
clipping, update rule) are pinned by the parity tests (CPU restatement and the compiled reference C++), the loop
itself -- including the reference's step-level bookkeeping of which observation selects the first
action of a rollout -- is restated, not pinned.
"""
import numpy as np
import torch

from . import _lib, ops
from ._lib import ARS_RLGLUE, DELTA_01, RLGLUE
from .engine import ArsEngine


def read_parameters(path):
    """rlglue/parameters.txt: whitespace-separated `key value...` lines, parsed independently by the
    environment (SwimmerEnvironment.cpp:297-326), the agent (SwimmerAgent.py:243-257) and the
    experiment (SwimmerExperiment.cpp:48-60)."""
    out = {}
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            vals = [float(x) for x in tok[1:]]
            out[tok[0]] = vals[0] if len(vals) == 1 else vals
    return out


class RlglueArsExperiment:
    def __init__(self, n_seg=3, direction=(1.0, 0.0), h_global=0.01, N=1, b=1, H=1000, alpha=0.02, nu=0.02,
                 max_u=5.0, l_i=1.0, k=10.0, m_i=1.0, seed=0, device=None):
        self.params = _lib.make_params(n=int(n_seg), l_i=l_i, m_i=m_i, k=k, h=h_global, max_u=max_u,
                                       direction=direction)
        self.N, self.b, self.H = int(N), int(b), int(H)
        self.engine = ArsEngine(self.params, N=self.N, b=self.b, alpha=alpha, nu=nu, H=self.H,
                                semantics=ARS_RLGLUE, variant=RLGLUE, delta_dist=DELTA_01, clip_actions=True,
                                seed=seed, distributed=False, device=device)
        self.results = []

    @classmethod
    def from_parameters_file(cls, path, **overrides):
        p = read_parameters(path)
        for key in ("n_seg", "N", "b", "H"):
            if key in p:
                p[key] = int(p[key])
        p.update(overrides)
        return cls(**p)

    @property
    def policy(self):
        return self.engine.policy_numpy()

    def run_one_training_iteration(self):
        """2N training rollouts + update, then the frozen evaluation rollout (SwimmerExperiment.cpp:65-84).
        Returns the evaluation return as a device scalar tensor; nothing synchronises."""
        eng = self.engine
        eng.run_iteration()
        ev = ops.rollout(self.params, self.H, variant=RLGLUE, policies=eng.W.reshape(1, -1),
                         clip_actions=True)
        return ev.returns

    def run_training(self, n_it):
        """-> np.ndarray[n_it] of evaluation returns, the numbers of rlglue/plot/results.txt."""
        evs = [self.run_one_training_iteration().clone() for _ in range(int(n_it))]
        out = torch.cat(evs).cpu().numpy() if evs else np.zeros(0)
        self.results += out.tolist()
        return out

    def write_results(self, path):
        """Same line format as SwimmerExperiment.cpp:83."""
        with open(path, "w") as f:
            for i, r in enumerate(self.results):
                f.write("Reward for one rollout with policy at iteration %d: %g\n" % (i, r))

"""ARS V1 with the RL-Glue agent's semantics.

The reference spreads this loop over four processes talking over sockets: the experiment driver
(rlglue/experiment/SwimmerExperiment.cpp:65-100) issues 2*N*H training `RL_step`s per iteration,
re-loading the saved start state every H steps, then freezes the agent and runs one H-step evaluation
rollout whose return is the line written to plot/results.txt; the agent
(rlglue/agent/SwimmerAgent.py:79-241) draws delta ~ U[0,1) (:203-212; the docstring says normal),
clips actions to +-max_u (:181-201), does NOT sort the directions (:214-221 is a TODO returning the
identity), uses the first b of them with the *sample* standard deviation (:223-241), and the
environment (rlglue/environment/SwimmerEnvironment.cpp) integrates semi-implicitly from the start
state 0.001 (:39-42).

Two modes:

  protocol="batched" (default)  the intended algorithm at rollout granularity: one fused launch of the 2N
      independent training rollouts (dynamics variant RLGLUE, clipped actions, U[0,1) Philox perturbations),
      the index-order / sample-std update, and one evaluation rollout of the updated policy;
  protocol="reference"          the reference's literal step-level state machine (csrc/rlglue_protocol.cu):
      the first environment step after every "load state" still applies the action chosen at the end of the
      previous rollout, the agent acts on the stored initial observation at the first step of a rollout,
      every return is shifted by one step and the last slot of the reward table is filled one iteration
      late.  Consecutive rollouts are coupled, so an experiment is sequential (one thread per replica).
      Pinned to the unmodified agent + the compiled reference C++ environment driven the way
      SwimmerExperiment.cpp does (tests/golden/rlglue_agent.npz): evaluation returns and policies agree to
      1e-9 relative.
"""
import ctypes

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
                 max_u=5.0, l_i=1.0, k=10.0, m_i=1.0, seed=0, device=None, protocol="batched", replicas=1):
        if protocol not in ("batched", "reference"):
            raise ValueError("protocol must be 'batched' or 'reference'")
        self.params = _lib.make_params(n=int(n_seg), l_i=l_i, m_i=m_i, k=k, h=h_global, max_u=max_u,
                                       direction=direction)
        self.N, self.b, self.H = int(N), int(b), int(H)
        self.alpha, self.nu, self.seed, self.protocol = float(alpha), float(nu), int(seed), protocol
        self.results = []
        if protocol == "batched":
            self.engine = ArsEngine(self.params, N=self.N, b=self.b, alpha=alpha, nu=nu, H=self.H,
                                    semantics=ARS_RLGLUE, variant=RLGLUE, delta_dist=DELTA_01, clip_actions=True,
                                    seed=seed, distributed=False, device=device)
            return
        # step-level reference protocol: `replicas` independent experiments (seed + replica), one thread each
        _lib.require_cuda()
        self.engine = None
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.replicas = int(replicas)
        n = int(n_seg)
        self.ws = (n - 1) * (2 * n + 2)
        self._state = torch.zeros(self.replicas, _lib.lib().swm_rlglue_protocol_state_doubles(n),
                                  dtype=torch.float64, device=self.device)
        self.iterations_done = 0
        self.reward_tables = []

    def run_reference_protocol(self, n_it, deltas=None):
        """n_it more iterations of the reference's step-level loop for every replica.  deltas (optional):
        [iterations_done + n_it (+1), N, n-1, 2n+2] U[0,1) draws replacing Philox (replays the agent's
        np.random.rand).  -> (results[replicas, n_it], reward_tables[replicas, n_it, 2N]) device tensors."""
        if self.protocol != "reference":
            raise ValueError("construct with protocol='reference'")
        cfg = _lib.SwmRlglueProtocol()
        cfg.N, cfg.b, cfg.H, cfg.n_it = self.N, self.b, self.H, int(n_it)
        cfg.alpha, cfg.nu, cfg.seed, cfg.iteration0 = self.alpha, self.nu, self.seed & (2 ** 64 - 1), 0
        if deltas is not None:
            deltas = torch.as_tensor(np.ascontiguousarray(deltas, dtype=np.float64)).to(self.device).reshape(-1, self.N, self.ws)
            if deltas.shape[0] < self.iterations_done + int(n_it):
                raise ValueError("deltas must cover every iteration since the start of the experiment")
            cfg.deltas = deltas.data_ptr()
        res = torch.zeros(self.replicas, int(n_it), dtype=torch.float64, device=self.device)
        tab = torch.zeros(self.replicas, int(n_it), 2 * self.N, dtype=torch.float64, device=self.device)
        cfg.state, cfg.results, cfg.table, cfg.replicas = self._state.data_ptr(), res.data_ptr(), tab.data_ptr(), self.replicas
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().swm_rlglue_protocol(ctypes.byref(self.params), ctypes.byref(cfg), _lib.stream_ptr()))
        self.iterations_done += int(n_it)
        self.reward_tables.append(tab)
        return res, tab

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
        if self.protocol == "reference":   # replica 0
            n = self.params.n
            return self._state[0, :self.ws].cpu().numpy().reshape(n - 1, 2 * n + 2).copy()
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
        if self.protocol == "reference":
            out = self.run_reference_protocol(n_it)[0][0].cpu().numpy()
            self.results += out.tolist()
            return out
        evs = [self.run_one_training_iteration().clone() for _ in range(int(n_it))]
        out = torch.cat(evs).cpu().numpy() if evs else np.zeros(0)
        self.results += out.tolist()
        return out

    def write_results(self, path):
        """Same line format as SwimmerExperiment.cpp:83."""
        with open(path, "w") as f:
            for i, r in enumerate(self.results):
                f.write("Reward for one rollout with policy at iteration %d: %g\n" % (i, r))

"""Seed fan-out: the reference runs one `ARSAgent` Ray actor per random seed
(ars/experiment.py:64-72, `ray.init(num_cpus=8)` in ars/plot_graph.py:11) and stacks the learning
curves into `r_graphs[n_seed, n_iter + 1]`.  Here the seeds share one GPU: every seed owns an
`ArsEngine` and a CUDA stream, the iterations of all seeds are enqueued without ever synchronising
with the host, and the small per-seed kernels (a 16-rollout iteration of BASELINE config[0] is a
single CTA) overlap on the device.  Seeds are independent replicas: different Philox keys, no
exchange between them.

`SeedFanout` is the batched engine-level entry point; `Experiment` keeps the reference's
constructor / `plot(n_seed, agent_param)` call shape and `.npy` output, without the matplotlib
figures (plotting is out of scope).  The scripts ars/plot_graph.py and
ars/safe_exploration.py run on it with only their imports changed (examples/).
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import ARS_AGENT
from .engine import ArsEngine


class SeedFanout:
    def __init__(self, params, seeds, *, device=None, max_iterations=4096, use_graph=True, **engine_kwargs):
        """`seeds`: iterable of ints; `engine_kwargs`: everything `ArsEngine` takes except `seed`
        (N, b, alpha, nu, H, v2, semantics, sim_params, ...).  Always single-process
        (`distributed=False`): the seeds are the parallel dimension.  Every engine records its
        learning curve on the device (`max_iterations` entries) and, with `use_graph`, replays one
        captured CUDA graph per iteration, so a round over all seeds costs one launch per seed."""
        _lib.require_cuda()
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self.seeds = [int(s) for s in seeds]
        if not self.seeds:
            raise ValueError("at least one seed")
        engine_kwargs = dict(engine_kwargs)
        engine_kwargs["distributed"] = False
        engine_kwargs["use_graph"] = use_graph
        engine_kwargs["curve_capacity"] = int(max_iterations)
        self.max_iterations = int(max_iterations)
        self.streams, self.engines = [], []
        for s in self.seeds:
            st = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(st):
                eng = ArsEngine(params, seed=s, device=self.device, **engine_kwargs)
            self.streams.append(st)
            self.engines.append(eng)
        self._curves = None

    def run(self, n_iter, include_initial=True):
        """Enqueues `n_iter (+1)` iterations of every seed round-robin and returns the learning
        curves `[n_seeds, n_iter + 1]`: entry j = mean of the returns of iteration j over the
        directions that were rolled out, or the previous entry when every direction was screened
        out (ars_agent.py:195-201).  One host synchronisation, at the end."""
        total = n_iter + (1 if include_initial else 0)
        start = self.engines[0].iteration
        if start + total > self.max_iterations:
            raise ValueError("max_iterations=%d is too small for %d more iterations" % (self.max_iterations, total))
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        for j in range(total):
            for eng, st in zip(self.engines, self.streams):
                with torch.cuda.stream(st):
                    eng.run_iteration()
        for st in self.streams:
            cur.wait_stream(st)
        out = torch.stack([e.curve[start:start + total] for e in self.engines]).cpu().numpy()
        for i, row in enumerate(out):  # all directions screened out -> previous value
            for j in range(total):
                if np.isnan(row[j]):
                    row[j] = row[j - 1] if j > 0 else (self._curves[i][-1] if self._curves is not None else np.nan)
        self._curves = out
        return out

    def policies(self):
        return np.stack([e.policy_numpy() for e in self.engines])


class Experiment:
    """ars/experiment.py:12-122 without the figures: `plot(n_seed, agent_param)` trains `n_seed`
    agents (seeds 0..n_seed-1, like `seed=i` at experiment.py:65) on one GPU and returns
    `r_graphs[n_seed, n_iter + 1]`; the array is saved under `results_path/array/` with the
    reference's file name when `results_path` is given."""

    def __init__(self, real_env_param, results_path=None, data_path=None, save_data_path=None,
                 save_policy_path=None, guess_param=None, approx_error=None, sim_thresh=None, device=None):
        self.real_env_param = real_env_param
        self.results_path = results_path
        self.data_path, self.save_data_path, self.save_policy_path = data_path, save_data_path, save_policy_path
        self.guess_param, self.approx_error, self.sim_thresh = guess_param, approx_error, sim_thresh
        self.device = device

    def plot(self, n_seed, agent_param, plot_mean=True, delta_source=None):
        """`delta_source`: "philox" = seed fan-out on CUDA streams (default for plain training), "numpy" = one
        `ARSAgent` per seed consuming numpy's global stream exactly like the reference (same seed, same
        curve).  Safe exploration and trajectory recording (`save_data_path`) always go through
        `ARSAgent`, which owns the database / estimator / simulator-threshold logic of
        ars_agent.py:37-71."""
        p = self.real_env_param
        via_agents = (agent_param.safe or self.save_data_path is not None or delta_source == "numpy")
        if via_agents:
            from .ars_agent import ARSAgent
            curves, last = [], None
            for i in range(n_seed):  # experiment.py:64-72, one actor per seed
                last = ARSAgent(p, agent_param, seed=i, data_path=self.data_path, guess_param=self.guess_param,
                                approx_error=self.approx_error, sim_thresh=self.sim_thresh,
                                delta_source=delta_source or "numpy", device=self.device)
                curves.append(last.runTraining(save_data_path=self.save_data_path,
                                               save_policy_path=self.save_policy_path))
            r_graphs = np.array(curves)
            self._save_array(agent_param, r_graphs)
            return r_graphs
        if agent_param.initial_w == "Zero":
            W0 = None
        else:
            W0 = np.load(agent_param.initial_w)
        fan = SeedFanout(
            _lib.make_params(n=p.n, l_i=p.l_i, m_i=p.m_i, k=p.k, h=p.h), range(n_seed), device=self.device,
            N=agent_param.N, b=agent_param.b, alpha=agent_param.alpha, nu=agent_param.nu, H=agent_param.H,
            v2=not agent_param.V1, semantics=ARS_AGENT, initial_policy=W0)
        r_graphs = fan.run(agent_param.n_iter)
        if self.save_policy_path is not None:
            np.save(self.save_policy_path, fan.policies()[-1])  # the reference's actors overwrite one file
        self._save_array(agent_param, r_graphs)
        return r_graphs

    def _save_array(self, agent_param, r_graphs):
        """`np.save` of the curves under results_path/array/ with the reference's file name
        (experiment.py:90-93)."""
        if self.results_path is None:
            return
        p = self.real_env_param
        ars = (f"{agent_param.name}, ARS_{'V1' if agent_param.V1 else 'V2'}"
               f"{'-t' if agent_param.b < agent_param.N else ''}, n_directions={agent_param.N}, "
               f"deltas_used={agent_param.b}, step_size={agent_param.alpha}, delta_std={agent_param.nu}")
        env = (f"{p.name}, n_segments={p.n}, m_i={round(p.m_i, 2)}, l_i={round(p.l_i, 2)}, "
               f"epsilon={round(p.epsilon, 4)}, deltaT={p.h}")
        d = os.path.join(self.results_path, "array")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, f"{env.replace(', ', '-')}-{ars.replace(', ', '-')}"), r_graphs)

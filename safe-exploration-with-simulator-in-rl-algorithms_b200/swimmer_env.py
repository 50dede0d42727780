"""SwimmerEnv: the gym.Env-shaped plugin surface of the reference
(envs/gym_swimmer/swimmer/remy_swimmer_env.py:13-251) on top of the CUDA library, plus the
batched entry points the reference lacks.

Single-environment calls (`reset/step/set_state/...`) keep the reference's conventions: Python
lists out of reset/step, `(ob, reward, False, {})` from step, no action clipping, mutable host
attributes `G_dot / theta / theta_dot`.  Each such call is one B=1 kernel launch -- it is the
compatibility path.  The fast paths are `step_batched`, `rollout_batched` and the ARS classes.
"""
import math

import numpy as np
import torch

from . import _lib, ops
from ._lib import GYM, RLGLUE


class Box:
    """The two attributes callers read from gym.spaces.Box (`.shape`, bounds)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box(%s, %s, %s)" % (self.low, self.high, self.shape)


_VARIANTS = {"gym": GYM, "rlglue": RLGLUE, GYM: GYM, RLGLUE: RLGLUE}


class SwimmerEnv:
    metadata = {"render.modes": ["human"]}

    def __init__(self, envName="LeonSwimmer-v0", direction=[1., 0.], n=3, max_u=5., l_i=1., k=10.,
                 m_i=1., h=0.001, variant="gym", device=None):
        self.direction = np.array(direction, dtype=np.float64)
        self.n, self.max_u, self.l_i, self.k, self.m_i, self.h = n, max_u, l_i, k, m_i, h
        self.envName = envName
        self.variant = _VARIANTS[variant]
        self.device = torch.device(device) if device is not None else None
        inf = 1000
        self.observation_space = Box(-inf, inf, (2 * n + 2,))
        self.action_space = Box(-max_u, max_u, (n - 1,))
        _lib.make_params(n=n)  # validates n early
        self._batch = None     # device state of the batched interface

    # ---- plumbing ----
    def _dev(self):
        _lib.require_cuda()
        return self.device or torch.device("cuda", torch.cuda.current_device())

    def params(self):
        """C-ABI parameter struct from the *current* attribute values (they are mutable in the
        reference, which reads self.l_i etc. on every step)."""
        return _lib.make_params(n=self.n, l_i=self.l_i, m_i=self.m_i, k=self.k, h=self.h,
                                max_u=self.max_u, direction=self.direction)

    def _pack(self, G_dot, theta, theta_dot):
        s = np.empty(2 * self.n + 2, dtype=np.float64)
        s[:2] = np.asarray(G_dot, dtype=np.float64)
        s[2::2] = np.asarray(theta, dtype=np.float64)
        s[3::2] = np.asarray(theta_dot, dtype=np.float64)
        return s

    def _up(self, a, shape):
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(shape))
        return t.to(self._dev())

    # ---- gym surface (remy_swimmer_env.py:41-67, 216-251) ----
    def reset(self):
        if self.variant == GYM:
            self.G_dot = np.full(2, 0.)
            self.theta = np.full(self.n, math.pi / 2)
            self.theta_dot = np.full(self.n, 0.)
        else:  # env_start of SwimmerEnvironment.cpp:39-42
            self.G_dot = np.full(2, 0.001)
            self.theta = np.full(self.n, 0.001)
            self.theta_dot = np.full(self.n, 0.001)
        return self.get_state()

    def step(self, action):
        self.G_dot, self.theta, self.theta_dot = self.next_observation(
            action, self.G_dot, self.theta, self.theta_dot)
        return self.get_state(), self.get_reward(), self.check_terminal(), {}

    def next_observation(self, torque, G_dot, theta, theta_dot):
        """One integration step as a pure function of its arguments (remy_swimmer_env.py:69-93)."""
        st = self._up(self._pack(G_dot, theta, theta_dot), (1, -1))
        ac = self._up(torque, (1, self.n - 1))
        nxt, _ = ops.step_batched(self.params(), st, ac, self.variant, want_reward=False)
        nxt = nxt[0].cpu().numpy()
        return nxt[:2].copy(), nxt[2::2].copy(), nxt[3::2].copy()

    def compute_accelerations(self, torque, G_dot, theta, theta_dot):
        """(G_dotdot[2], theta_dotdot[n]) -- remy_swimmer_env.py:95-114."""
        st = self._up(self._pack(G_dot, theta, theta_dot), (1, -1))
        ac = self._up(torque, (1, self.n - 1))
        acc = ops.accelerations_batched(self.params(), st, ac, self.variant)[0].cpu().numpy()
        return acc[:2].copy(), acc[2:].copy()

    def get_state(self):
        ob = self.G_dot.tolist()
        for i in range(self.n):
            ob += [self.theta[i], self.theta_dot[i]]
        return ob

    def set_state(self, s):
        assert len(s) == 2 + 2 * self.n, f"State {s} has not the right dimension"
        self.reset()
        self.G_dot = np.array(s[:2], dtype=np.float64)
        for i in range(self.n):
            self.theta[i] = s[2 + 2 * i]
            self.theta_dot[i] = s[3 + 2 * i]

    def get_reward(self):
        return self.G_dot.dot(self.direction)

    def check_terminal(self):
        return False

    def render(self, mode="human"):
        return

    def close(self):
        return

    # ---- batched surface (new) ----
    def reset_batched(self, B):
        """-> states[B, 2n+2] on the device; also becomes the batch the env is stepping."""
        self._batch = ops.reset_state(self.n, self.variant, self._dev()).repeat(B, 1).contiguous()
        return self._batch

    def set_state_batched(self, states):
        states = torch.as_tensor(states, dtype=torch.float64).to(self._dev()).contiguous()
        assert states.dim() == 2 and states.shape[1] == 2 * self.n + 2
        self._batch = states.clone()
        return self._batch

    def get_state_batched(self):
        return self._batch

    def step_batched(self, actions):
        """actions[B, n-1] -> (states[B, 2n+2], rewards[B], dones[B], {}) advancing the batch in
        place.  No clipping (the gym path never clips)."""
        if self._batch is None:
            raise RuntimeError("call reset_batched(B) or set_state_batched(states) first")
        actions = torch.as_tensor(actions, dtype=torch.float64).to(self._dev()).contiguous()
        _, rew = ops.step_batched(self.params(), self._batch, actions, self.variant, out=self._batch)
        dones = torch.zeros(self._batch.shape[0], dtype=torch.bool, device=self._batch.device)
        return self._batch, rew, dones, {}

    def rollout_batched(self, H, actions=None, policies=None, **kw):
        """Fused H-step rollout of a batch from reset (or `init_state=`): ops.rollout on this
        environment's parameters.  Returns an ops.RolloutResult of device tensors."""
        if actions is not None:
            actions = torch.as_tensor(actions, dtype=torch.float64).to(self._dev())
        if policies is not None:
            policies = torch.as_tensor(policies, dtype=torch.float64).to(self._dev())
        return ops.rollout(self.params(), H, variant=self.variant, actions=actions,
                           policies=policies, **kw)

    def rollout_plan(self, H, *, n_sub=16, chunk=64, actions=None, policies=None, **kw):
        """A reusable schedule of one fused rollout as `n_sub` sub-batches x `chunk`-step launches on several
        streams (ops.ChunkedRollout): `plan.run()` enqueues it (capturable in a CUDA graph) and returns the same
        results as `rollout_batched` -- final states bit-identical -- while filling the SMs better when the
        batch is a fractional number of warps per SM sub-partition (65,536 envs on a B200: +18 %).  The
        tensors are captured by reference: update them in place between runs."""
        src = actions if actions is not None else policies
        src = torch.as_tensor(src, dtype=torch.float64).to(self._dev())
        B = src.shape[0] * (kw.get("rollouts_per_policy", 1) if policies is not None else 1)
        return ops.ChunkedRollout(self.params(), H, B=B, n_sub=n_sub, chunk=chunk, variant=self.variant,
                                  actions=src if actions is not None else None,
                                  policies=src if policies is not None else None, **kw)

    # ---- host-buffer entry point (pinned memory in, pinned memory out), double-buffered ----
    def rollout_batched_host(self, H, actions_host, returns_host, final_host=None, **kw):
        """Fused rollout of one batch whose actions live in (pinned) HOST memory and whose results
        are delivered into (pinned) host tensors: H2D copy -> swm_rollout -> D2H copies, all enqueued
        on one of two internal streams so that consecutive calls overlap (the upload and download of
        one batch hide behind the kernel of its neighbour).  Returns a CUDA event; the host tensors
        are valid after `event.synchronize()` (or `synchronize_host()`).  The caller must not reuse
        a host output tensor before the call that fills it has completed."""
        dev = self._dev()
        B = actions_host.shape[0]
        no = 2 * self.n + 2
        if getattr(self, "_slots", None) is None:
            self._slots, self._slot_i = [], 0
        if not self._slots or self._slots[0]["act"].shape[0] != B:
            self.synchronize_host()
            self._slots = [{"stream": torch.cuda.Stream(device=dev),
                            "act": torch.empty(B, self.n - 1, dtype=torch.float64, device=dev),
                            "out": {"returns": torch.empty(B, dtype=torch.float64, device=dev),
                                    "final_state": torch.empty(B, no, dtype=torch.float64, device=dev)},
                            "event": None} for _ in range(2)]
        slot = self._slots[self._slot_i]
        self._slot_i ^= 1
        st = slot["stream"]
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            slot["act"].copy_(actions_host, non_blocking=True)
            res = ops.rollout(self.params(), H, variant=self.variant, actions=slot["act"],
                              want_final=final_host is not None,
                              out=slot["out"] if final_host is not None else {"returns": slot["out"]["returns"]},
                              **kw)
            returns_host.copy_(res.returns, non_blocking=True)
            if final_host is not None:
                final_host.copy_(res.final_state, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        slot["event"] = ev
        return ev

    def synchronize_host(self):
        """Waits for every outstanding `rollout_batched_host` call; the current stream also waits."""
        for slot in getattr(self, "_slots", None) or []:
            if slot["event"] is not None:
                torch.cuda.current_stream(self._dev()).wait_event(slot["event"])
                slot["event"].synchronize()

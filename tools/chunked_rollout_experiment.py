"""Experiment: does splitting one config[1] batch into S sub-batches x K time-chunks (separate launches on S
streams, state handed through final_state -> init_state, chunk lengths multiples of 64 so that the tracked
trig re-evaluation lines up and results stay bit-identical) remove the sub-partition quantisation of a
65,536-env batch?  Replayed from one CUDA graph so that launch overhead does not matter."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def build(p, actions, H, n_sub, chunk):
    B = actions.shape[0]
    no = 2 * p.n + 2
    bounds = np.linspace(0, B, n_sub + 1).astype(int)
    lens = []
    t = 0
    while t < H:
        lens.append(min(chunk, H - t))
        t += lens[-1]
    K = len(lens)
    state = torch.empty(B, no, dtype=torch.float64, device="cuda")
    rets = torch.zeros(K, B, dtype=torch.float64, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(n_sub)]
    subs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_sub)]

    def enqueue():
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for c, L in enumerate(lens):
            for (lo, hi), st in zip(subs, streams):
                with torch.cuda.stream(st):
                    S.ops.rollout(p, L, actions=actions[lo:hi], init_state=None if c == 0 else state[lo:hi],
                                  want_final=True, out={"returns": rets[c, lo:hi], "final_state": state[lo:hi]})
        for st in streams:
            cur.wait_stream(st)
    return enqueue, rets, state


def main():
    p = S.make_params(n=3)
    rng = np.random.default_rng(0)
    B, H = 65536, 1000
    actions = torch.as_tensor(rng.uniform(-5, 5, (B, 2))).cuda()
    ref = S.ops.rollout(p, H, actions=actions, want_final=True)
    for n_sub, chunk in ((1, 1000), (8, 64), (16, 64), (32, 64), (16, 128), (32, 128)):
        enqueue, rets, state = build(p, actions, H, n_sub, chunk)
        enqueue()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            enqueue()
        g.replay()
        torch.cuda.synchronize()
        ok_state = torch.equal(state, ref.final_state)
        err = float((rets.sum(0) - ref.returns).abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("sub-batches %2d, chunk %4d: %.4f ms  %.3e env-steps/s  final state bit-identical: %s, return diff %.1e"
              % (n_sub, chunk, ms, B * H / ms * 1e3, ok_state, err), flush=True)


if __name__ == "__main__":
    main()

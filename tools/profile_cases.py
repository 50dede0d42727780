"""Launches one named workload of the fused rollout kernel a few times (for `ncu -k regex:rollout_kernel`).

  python tools/profile_cases.py <case> [reps]

cases:  n3_fixed   BASELINE config[1]: n=3, 65,536 envs, fixed actions, H=1000
        n5_v2      BASELINE config[2]: n=5, ARS V2 + moments, 1,024 directions (2,048 envs), H=1000
                   (AUTO kernel choice = lane-split; n5_v2_thread forces one thread per environment,
                   n5_v2_256 is the per-GPU share at 8 GPUs)
        n10_grp    BASELINE config[4] per-GPU share at 8 GPUs: n=10, 512 directions x 2 x 128 rollouts
                   (131,072 envs), V2 + moments, H=1000
        step_n3    batched single step (swm_step_batched), n=3, 4,194,304 envs: the HBM-bound entry point
        n3_safe    BASELINE config[3]: n=3, 256 directions, per-step screened rollouts, H=1000
Prints the CUDA-event time per launch.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def main():
    case = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rng = np.random.default_rng(0)
    H = 1000
    if case in ("n3_fixed", "n3_fixed_plain"):   # default call (library schedule) / one forced plain launch
        p = S.make_params(n=3)
        ac = torch.as_tensor(rng.uniform(-5, 5, (65536, 2))).cuda()
        fn = lambda: S.ops.rollout(p, H, actions=ac, want_final=True, schedule="plain" if case.endswith("plain") else None)
        B = 65536
    elif case in ("n5_v2", "n5_v2_thread", "n5_v2_256", "n5_v2_1024", "n10_grp"):
        n, D, R = (10, 512, 128) if case == "n10_grp" else (5, {"n5_v2_256": 128, "n5_v2_1024": 512}.get(case, 1024), 1)
        kern = S.KERNEL_THREAD if case == "n5_v2_thread" else int(os.environ.get("SWM_PROFILE_KERNEL", S.KERNEL_AUTO))
        p = S.make_params(n=n)
        no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
        W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.05).cuda()
        mean = torch.zeros(no, dtype=torch.float64, device="cuda")
        inv = torch.ones_like(mean)
        piv = S.ops.reset_state(n)
        B = 2 * D * R
        fn = lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.01, seed=1, mean=mean, inv_sigma=inv,
                                   stats_pivot=piv, rollouts_per_policy=R, kernel=kern,
                                   init_perturb=1e-2 if R > 1 else 0.0)
    elif case == "step_n3":
        p = S.make_params(n=3)
        B = 1 << 22
        st = torch.as_tensor(rng.normal(size=(B, 8))).cuda()
        ac = torch.as_tensor(rng.uniform(-5, 5, (B, 2))).cuda()
        o = torch.empty_like(st)
        H = 1
        fn = lambda: S.ops.step_batched(p, st, ac, out=o)
    elif case == "n3_safe":
        p = S.make_params(n=3, l_i=0.8, m_i=1.2, k=10.2)
        sim = S.make_params(n=3, l_i=0.8006, m_i=1.2006, k=10.2006)
        W = torch.as_tensor(rng.uniform(-1, 1, 16) * 0.05).cuda()
        B = 512
        fn = lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.01, seed=1,
                                   screen=dict(sim_params=sim, sim_thresh=50.0, real_thresh=51.0))
    else:
        raise SystemExit("unknown case " + case)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %.4f ms per launch, %d envs x %d steps -> %.4e env-steps/s" % (case, ms, B, H, B * H / ms * 1e3))


if __name__ == "__main__":
    main()

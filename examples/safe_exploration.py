"""ars/safe_exploration.py of the reference with only the imports changed (no ray, no matplotlib):
1. train a "hand" controller without safe exploration, recording its trajectories and policy;
2. safety threshold l = 0.99 x its final mean return;
3. for several Lipschitz constants A and approximation errors epsilon, train 8 seeds from the hand policy
   with reward-constraint safe exploration (every +-delta is first rolled out in a simulator whose
   parameters are off by epsilon; only directions whose simulated returns exceed l + alpha(H) epsilon are
   tried in the real world) and record the minimum real return and the best mean learning curve.

    python examples/safe_exploration.py [--quick]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swimmer_ars_b200 import ARSParam, EnvParam, Experiment, Threshold  # reference: from ars.parameters / ars.experiment

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true", help="few iterations / seeds / sweep points (smoke run)")
ap.add_argument("--data_dir", default="ars_data")
args = ap.parse_args()
os.makedirs(args.data_dir, exist_ok=True)
hand_iters, real_iters, n_seed = (200, 400, 8) if not args.quick else (10, 5, 2)
A_values = [0.1, 0.3, 0.5, 0.7] if not args.quick else [0.1]
epsilon_range = np.linspace(0.0001, 0.01, 10) if not args.quick else np.array([0.0001, 0.01])
H = 1000 if not args.quick else 200

real_env_param = EnvParam('LeonSwimmer-RealWorld', n=3, H=H, l_i=.8, m_i=1.2, h=1e-3, k=10.2, epsilon=0.001)

# the hand controller (safe_exploration.py:29-40)
hand_agent = ARSParam('HandControl', V1=True, n_iter=hand_iters, H=H, N=1, b=1, alpha=0.0075, nu=0.01, safe=False,
                      threshold=0, initial_w='Zero')
data_file = os.path.join(args.data_dir, "real_world_2.npz")
policy_file = os.path.join(args.data_dir, "saved_hand_policy")
hand_exp = Experiment(real_env_param, data_path=None, save_data_path=data_file, save_policy_path=policy_file,
                      guess_param=None)
returns = hand_exp.plot(n_seed=1, agent_param=hand_agent)

# safety threshold from the known controller (safe_exploration.py:42-46)
l = float(np.mean(returns, axis=0)[-1] * 0.99)
print(f"\nSafety threshold: {l}")
np.savetxt(os.path.join(args.data_dir, "threshold.txt"), np.array([l]))

K, B = 1, 0.001
for A in A_values:
    sim_thresh = Threshold(K=K, A=A, B=B)
    alpha = sim_thresh.compute_alpha(H)
    print(f"B = {B}; alpha = {alpha}")
    min_return, max_mean_returns = [], []
    for epsilon in epsilon_range:
        real_env_param = EnvParam('LeonSwimmer-RealWorld', n=3, H=H, l_i=.8, m_i=1.2, h=1e-3, k=10.2,
                                  epsilon=float(epsilon))
        real_agent = ARSParam('RLControl', V1=True, n_iter=real_iters, H=H, N=1, b=1, alpha=0.0075, nu=0.01,
                              safe=True, threshold=l, initial_w=policy_file + '.npy')
        real_exp = Experiment(real_env_param, data_path=data_file, save_data_path=None, save_policy_path=None,
                              guess_param=None, approx_error=float(epsilon), sim_thresh=sim_thresh)
        r_graphs = real_exp.plot(n_seed=n_seed, agent_param=real_agent)
        min_return.append(np.nanmin(r_graphs))
        max_mean_returns.append(np.nanmax(np.mean(r_graphs, axis=0)))
    out = os.path.join(args.data_dir, f"epsilon_sim_threshold_H={H}_K={K}_A={A}_B={B}.npz")
    np.savez(out, epsilon=epsilon_range, min_return=min_return, max_mean_returns=max_mean_returns,
             sim_threshold=[l + alpha * e for e in epsilon_range], safety_threshold=l)
    print(f"A={A}: min real return {min(min_return):.4g} (threshold {l:.4g}); saved {out}")

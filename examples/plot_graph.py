"""ars/plot_graph.py of the reference with only the imports changed (no ray, no matplotlib): learning
curves of ARS V1 on the 3-segment swimmer for several seeds, saved as .npy under results/gym/array/.

    python examples/plot_graph.py [--n_seed 4] [--n_iter 100]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swimmer_ars_b200 import ARSParam, EnvParam, Experiment  # reference: from ars.experiment / ars.parameters

ap = argparse.ArgumentParser()
ap.add_argument("--n_seed", type=int, default=4)
ap.add_argument("--n_iter", type=int, default=100)
ap.add_argument("--results_path", default="results/gym/")
args = ap.parse_args()

# parameters of ars/plot_graph.py:14-21
real_env_param = EnvParam('LeonSwimmer-RealWorld', n=3, H=1000, l_i=.8, m_i=1.2, h=1e-3, k=10.2, epsilon=0)
ars_agent_param = ARSParam('RLControl', V1=True, n_iter=args.n_iter, H=1000, N=1, b=1, alpha=0.0075, nu=0.01,
                           safe=False, threshold=0, initial_w='Zero')
exp = Experiment(real_env_param, results_path=args.results_path)
r_graphs = exp.plot(args.n_seed, ars_agent_param)  # all seeds advance concurrently on one GPU
print("learning curves", r_graphs.shape, "final mean returns per seed:", r_graphs[:, -1])

"""Configuration surface of the ARS scripts, field-for-field the reference's dataclasses
(ars/parameters.py:11-45) so that existing experiment scripts construct them unchanged."""
from dataclasses import dataclass


@dataclass
class EnvParam:
    """Swimmer environment parameters (ars/parameters.py:11-20)."""
    name: str
    n: int          # number of segments
    H: int          # rollout length
    l_i: float      # segment length
    m_i: float      # segment mass
    h: float        # integration step
    k: float        # viscosity coefficient
    epsilon: float  # approximation error of the simulator parameters


@dataclass
class ARSParam:
    """ARS agent parameters (ars/parameters.py:24-36)."""
    name: str
    V1: bool          # True: ARS V1 (no observation normalisation), False: V2
    n_iter: int       # training iterations
    H: int            # rollout length
    N: int            # sampled perturbation directions
    b: int            # directions used for the update
    alpha: float      # step size
    nu: float         # perturbation scale
    safe: bool        # reward-constraint safe exploration through the simulator
    threshold: float  # safety threshold on the real return
    initial_w: str    # 'Zero' or a path to a .npy policy


@dataclass
class Threshold:
    """Lipschitz constants of the simulator-threshold bound (ars/parameters.py:38-45)."""
    K: float  # reward function
    A: float  # transition function w.r.t. the parameters
    B: float  # transition function w.r.t. the state

    def compute_alpha(self, H):
        # alpha(H) = K A / (1 - B) * (H - B (1 - B^H) / (1 - B))
        g = 1.0 - self.B
        return self.K * self.A / g * (H - self.B * (1.0 - self.B ** H) / g)

"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference Python files from /root/reference so that they can
be run as the parity oracle and used to generate the golden fixtures under
tests/golden/ (see oracle/make_golden.py).  /root/reference only exists in the build
container, not on the GPU box: everything that runs on the GPU box uses the committed
fixtures and the C restatement (oracle/swimmer_oracle.c) instead.

The reference needs three packages that are not installed here (gym, ray, cma).
They are only touched for a base class, a `spaces.Box(...).shape`, a class decorator
and an import statement, so tiny stand-ins are injected into sys.modules:

  gym.Env                       base class of SwimmerEnv  (remy_swimmer_env.py:13)
  gym.spaces.Box                .shape only               (remy_swimmer_env.py:36-39)
  gym.envs.swimmer.remy_swimmer_env   import path used by ars/environment.py:6
  ray.remote / init / get       decorator on ARSAgent     (ars/ars_agent.py:15)
  cma                           imported by ars/estimator.py:13 (never called here)
"""
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("SWIMMER_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "envs/gym_swimmer/swimmer/remy_swimmer_env.py"))


def _install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_swimmer_stub", False):
        return
    gym = types.ModuleType("gym")
    gym._swimmer_stub = True

    class Env:
        metadata = {}

        def close(self):
            return None

    class Box:
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    gym.Env = Env
    gym.spaces = spaces
    envs = types.ModuleType("gym.envs")
    envs.__path__ = []
    swimmer = types.ModuleType("gym.envs.swimmer")
    swimmer.__path__ = []
    gym.envs = envs
    envs.swimmer = swimmer
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                        "gym.envs.swimmer": swimmer})

    ray = types.ModuleType("ray")
    ray.remote = lambda c: c
    ray.init = lambda *a, **k: None
    ray.get = lambda x: x
    sys.modules["ray"] = ray
    sys.modules.setdefault("cma", types.ModuleType("cma"))


def load():
    """Returns a namespace with the reference classes, loaded unmodified."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    # remy_swimmer_env.py:130-133 assigns 1-element arrays to scalars (numpy>=1.25 warns)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    name = "gym.envs.swimmer.remy_swimmer_env"
    if name not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REFERENCE_ROOT,
                               "envs/gym_swimmer/swimmer/remy_swimmer_env.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        sys.modules["gym.envs.swimmer"].remy_swimmer_env = mod
        sys.modules["gym.envs.swimmer"].SwimmerEnv = mod.SwimmerEnv
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.SwimmerEnv = sys.modules[name].SwimmerEnv
    from ars.parameters import EnvParam, ARSParam, Threshold
    from ars.environment import Environment
    from ars.ars_agent import ARSAgent
    from ars.database import Database
    from safe_ars.ars import Basic_ARS, Safe_ARS
    ns.EnvParam, ns.ARSParam, ns.Threshold = EnvParam, ARSParam, Threshold
    ns.Environment, ns.ARSAgent, ns.Database = Environment, ARSAgent, Database
    ns.Basic_ARS, ns.Safe_ARS = Basic_ARS, Safe_ARS
    return ns

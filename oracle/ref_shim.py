"""TEST INFRASTRUCTURE ONLY -- import the real aravindr93/mjrl reference on CPU.

Usable where a reference checkout exists: $MJRL_REF, /root/reference (the build container) or
baseline/_ref (the offline `pip install --target` of the unmodified reference, git-ignored, which
travels to the GPU box with the snapshot; written by __graft_entry__.build()).  Used by oracle/make_golden.py to generate the
committed fixtures in tests/golden/ and by tests/test_oracle_vs_reference.py to pin the
restatement in oracle/npg_oracle.py against the reference itself.

Import recipe:
  1. pre-register an empty `mjrl` package whose __path__ points at <ref>/mjrl, which
     bypasses mjrl/__init__.py (`import mjrl.envs` -> gym + mujoco_py);
  2. stub `gym` (attr Env), `matplotlib`, `matplotlib.pyplot`;
  3. stub `mjrl.samplers.batch_sampler` (trpo.py:15 imports a module that does not exist).
Nothing from the reference is copied; it is imported where it lies.
"""
import os
import sys
import types


def reference_root():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("MJRL_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "mjrl", "algos")):
            return cand
    return None


def available():
    return reference_root() is not None


_loaded = None


def load():
    """Return a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("mjrl reference not found ($MJRL_REF, /root/reference or baseline/_ref)")
    if "mjrl" in sys.modules and not getattr(sys.modules["mjrl"], "_b200_shim", False):
        raise RuntimeError("a real `mjrl` package is already imported; cannot shim")
    pkg = types.ModuleType("mjrl")
    pkg.__path__ = [os.path.join(root, "mjrl")]
    pkg._b200_shim = True
    sys.modules["mjrl"] = pkg
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = object
        sys.modules["gym"] = gym
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    import importlib
    importlib.import_module("mjrl.samplers")
    sys.modules.setdefault("mjrl.samplers.batch_sampler", types.ModuleType("mjrl.samplers.batch_sampler"))

    ns = types.SimpleNamespace()
    from mjrl.utils.gym_env import EnvSpec
    from mjrl.policies.gaussian_mlp import MLP
    from mjrl.policies.gaussian_linear import LinearPolicy
    from mjrl.baselines.mlp_baseline import MLPBaseline
    from mjrl.algos.npg_cg import NPG
    from mjrl.algos.trpo import TRPO
    from mjrl.algos.dapg import DAPG
    from mjrl.algos.batch_reinforce import BatchREINFORCE
    from mjrl.utils.cg_solve import cg_solve
    import mjrl.utils.process_samples as process_samples
    ns.EnvSpec, ns.MLP, ns.LinearPolicy, ns.MLPBaseline = EnvSpec, MLP, LinearPolicy, MLPBaseline
    ns.NPG, ns.TRPO, ns.DAPG, ns.BatchREINFORCE = NPG, TRPO, DAPG, BatchREINFORCE
    ns.cg_solve, ns.process_samples = cg_solve, process_samples
    from mjrl.baselines.linear_baseline import LinearBaseline
    from mjrl.baselines.quadratic_baseline import QuadraticBaseline
    ns.LinearBaseline, ns.QuadraticBaseline = LinearBaseline, QuadraticBaseline
    _loaded = ns
    return ns

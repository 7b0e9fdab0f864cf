"""Developer tool: host-side wall-clock breakdown of one cfg3 device step (with a synchronize after every piece)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200 import runtime  # noqa: E402
from mjrl_b200.engine import Engine  # noqa: E402
from oracle import npg_oracle as O  # noqa: E402

n_paths, T, obs_dim, act_dim, hidden = 1000, 1000, 17, 6, (128, 128)
n = n_paths * T
rng = np.random.RandomState(0)
eng = Engine(obs_dim, act_dim, hidden, max_samples=n + 8, max_paths=n_paths + 1)
spec = O.PolicySpec(obs_dim, act_dim, hidden)
eng.set_params(O.init_policy_params(spec, 1))
eng.vf_set_state(O.VFState(obs_dim, (128, 128), seed=2).w)
eng.upload_flat(rng.randn(n, obs_dim).astype(np.float32), rng.randn(n, act_dim), rng.randn(n), np.full(n_paths, T, np.int32),
                np.zeros(n_paths, np.uint8))


def timed(name, fn, acc):
    eng.synchronize()
    t = time.perf_counter()
    r = fn()
    eng.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
    return r


acc = {}
for it in range(4):
    timed("compute_returns", lambda: eng.compute_returns(0.995), acc)
    timed("vf_predict", lambda: eng.vf_predict(), acc)
    timed("compute_advantages", lambda: eng.compute_advantages(0.995, 0.97), acc)
    perm = timed("host permutation (runtime.global_permutation)", lambda: runtime.global_permutation(n), acc)
    timed("vf_fit_begin (upload, features, launch; returns before the fit ends)", lambda: eng.vf_fit_begin(perm, 64, 1e-3, 1e-3), acc)
    timed("process_paths", lambda: eng.process_paths(), acc)
    timed("policy step (concurrent with the fit)", lambda: eng.step("trpo", step_size=0.1, cg_iters=10, damping=1e-4), acc)
    timed("vf_fit_end (join + prepare predict weights)", lambda: eng.vf_fit_end(), acc)
tot = 0.0
for k, v in acc.items():
    m = float(np.mean(v[1:]))
    tot += m
    print("%-72s %8.2f ms" % (k, m))
print("%-72s %8.2f ms" % ("sum", tot))

"""Developer tool (GPU box): drift of the wide-input (K-split) tensor-core fit against the fp32 oracle, next to the fp32-FMA
kernel and the oracle's own fp64 chain, on the shapes of tests/test_engine_gpu.py::test_baseline_fit_wide_inputs."""
import sys

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402
from oracle import npg_oracle as O  # noqa: E402
import torch  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


for obs_dim in (39, 100, 376):
    paths = O.synthetic_paths(obs_dim, 3, 16, 200, seed=obs_dim)
    O.compute_returns(paths, 0.995)
    n = sum(len(p["rewards"]) for p in paths)
    perms = [np.random.RandomState(5 + i).permutation(n).astype(np.int32) for i in range(3)]
    out = {}
    for name, dt in (("o32", torch.float32), ("o64", torch.float64)):
        vf = O.VFState(obs_dim, (128, 128), seed=4)
        w0 = vf.w.copy()
        O.vf_fit(vf, paths, perms[:2], 2, 64, 1e-3, 1e-3, dtype=dt)
        a = vf.w.copy()
        O.vf_fit(vf, paths, perms[2:], 1, 64, 1e-3, 1e-3, dtype=dt)
        out[name] = (a, vf.w.copy(), vf.m.copy(), vf.v.copy())
    for name, tc in (("tc", True), ("fma", False)):
        eng = Engine(obs_dim, 3, (64, 64), max_samples=n + 8, max_paths=32)
        eng.vf_set_state(w0)
        eng.vf_set_tensor_cores(tc)
        eng.upload_paths(paths)
        eng.compute_returns(0.995)
        eng.vf_fit(perms[:2], 64, 1e-3, 1e-3)
        a = eng.vf_get_state()[0]
        eng.vf_fit(perms[2:], 64, 1e-3, 1e-3)
        w, m, v, _ = eng.vf_get_state()
        out[name] = (a, w, m, v)
        eng.close()
    for name in ("o64", "tc", "fma"):
        print("obs %3d %-4s vs oracle-fp32: call1 w %.2e | call2 w %.2e m %.2e v %.2e" % (
            obs_dim, name, rel(out[name][0], out["o32"][0]), rel(out[name][1], out["o32"][1]),
            rel(out[name][2], out["o32"][2]), rel(out[name][3], out["o32"][3])), flush=True)

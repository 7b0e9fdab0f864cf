"""Developer tool (GPU box): how far does the baseline-fit chain of each kernel drift from the fp32 oracle, as a function
of the number of sequential Adam steps, up to cfg3's full 15 624 steps (1e6 timesteps / 64 - 1)?

The chain is chaotic (ReLU units switch on and off), so ANY two fp32 implementations with different summation orders
separate exponentially before they re-converge in function space; the yardstick printed next to each kernel is the
oracle against itself in fp64 (same inputs rounded to fp32, exact arithmetic from there): that distance is what fp32
rounding alone does to the reference's own chain.

    python tools/fit_fullsize_report.py [--max-steps 15624]
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402
from oracle import npg_oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-steps", type=int, default=15624)
    args = ap.parse_args()
    obs_dim, act_dim, horizon = 17, 6, 1000
    torch.set_num_threads(8)
    lines = []
    for steps in [s for s in (250, 1000, 4000, 15624) if s <= args.max_steps]:
        n = 64 * (steps + 1)
        n_paths = (n + horizon - 1) // horizon
        rng = np.random.RandomState(0)
        paths = []
        left = n
        for i in range(n_paths):
            T = min(horizon, left)
            left -= T
            paths.append(dict(observations=rng.randn(T, obs_dim), actions=rng.randn(T, act_dim), rewards=rng.randn(T),
                              terminated=False))
        O.compute_returns(paths, 0.995)
        perm = np.random.RandomState(3).permutation(n).astype(np.int32)
        vf0 = O.VFState(obs_dim, (128, 128), seed=1)
        res = {}
        for name, dt in (("oracle fp32", torch.float32), ("oracle fp64", torch.float64)):
            vf = O.VFState(obs_dim, (128, 128), seed=1)
            t0 = time.time()
            e = O.vf_fit(vf, paths, [perm], 1, 64, 1e-3, 1e-3, return_errors=True, dtype=dt)
            res[name] = (vf.w.astype(np.float64), e, time.time() - t0)
        for name, tc in (("vf_fit_tc_kernel (tcgen05)", True), ("vf_fit_kernel (fp32 FMA)", False)):
            eng = Engine(obs_dim, act_dim, (128, 128), max_samples=n + 8, max_paths=n_paths + 1)
            eng.vf_set_state(vf0.w, np.zeros_like(vf0.w), np.zeros_like(vf0.w), 0)
            eng.vf_set_tensor_cores(tc)
            eng.upload_paths(paths)
            eng.compute_returns(0.995)
            e = eng.vf_fit(perm, 64, 1e-3, 1e-3, return_errors=True)
            res[name] = (eng.vf_get_state()[0].astype(np.float64), e, eng.last_fit_ms() * 1e-3)
            eng.close()
        w32 = res["oracle fp32"][0]
        line = "steps %6d (N=%d)" % (steps, n)
        for name in ("oracle fp64", "vf_fit_tc_kernel (tcgen05)", "vf_fit_kernel (fp32 FMA)"):
            w, e, t = res[name]
            line += " | %s: weights rel vs oracle-fp32 %.2e, err_after %.6f (oracle-fp32 %.6f), %.2f s" % (
                name, rel(w, w32), e[1], res["oracle fp32"][1][1], t)
        print(line, flush=True)
        lines.append(line)


if __name__ == "__main__":
    main()

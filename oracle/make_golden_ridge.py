"""TEST INFRASTRUCTURE: golden vectors of the reference's LinearBaseline / QuadraticBaseline, produced by
importing the UNMODIFIED reference (oracle/ref_shim.py) in the build container:

    python oracle/make_golden_ridge.py        ->  tests/golden/ridge_{linear,quadratic}_{pm,swim}.npz

Each fixture: the synthetic trajectories' seeds / shapes, the returns, the coefficients of two consecutive fits (the
second on a fresh batch, so error_before is non-trivial), the (error_before, error_after) pairs, the per-sample
predictions and the GAE advantages computed by the reference's process_samples with that baseline.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import npg_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
GAMMA, LAM = 0.995, 0.97
from oracle import ridge_oracle as RO  # noqa: E402

CASES = RO.FIXTURE_CASES


def run(R, cls_name, cfg):
    spec = R.EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"])
    bl = getattr(R, cls_name)(spec)
    out = {}
    for rnd, seed in enumerate((0, 1)):
        paths = RO.fixture_paths(cfg, seed)
        R.process_samples.compute_returns(paths, GAMMA)
        R.process_samples.compute_advantages(paths, bl, GAMMA, LAM)       # PRE-fit baseline, as batch_reinforce.py:98
        out["adv%d" % rnd] = np.concatenate([p["advantages"] for p in paths])
        out["base%d" % rnd] = np.concatenate([p["baseline"] for p in paths])
        out["errs%d" % rnd] = np.array(bl.fit(paths, return_errors=True))
        out["coeffs%d" % rnd] = bl._coeffs.copy()
        out["pred%d" % rnd] = np.concatenate([bl.predict(p) for p in paths])
        out["returns%d" % rnd] = np.concatenate([p["returns"] for p in paths])
    out["meta"] = np.array(repr(dict(cfg, gamma=GAMMA, lam=LAM, cls=cls_name, reg_coeff=bl._reg_coeff, path_seeds=(0, 1))))
    return out


def main():
    R = ref_shim.load()
    for cls_name, tag in (("LinearBaseline", "linear"), ("QuadraticBaseline", "quadratic")):
        for name, cfg in CASES.items():
            out = run(R, cls_name, cfg)
            path = os.path.join(GOLDEN_DIR, "ridge_%s_%s.npz" % (tag, name))
            np.savez_compressed(path, **out)
            print("%-28s K=%-4d errs %s %s  %.0f KB" % (os.path.basename(path), out["coeffs0"].shape[0], out["errs0"], out["errs1"],
                                                        os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's ridge-regression baselines.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(mjrl_b200/baselines/linear_baseline.py, quadratic_baseline.py -> csrc/ridge.cu) never does.

Follows baselines/linear_baseline.py:11-60 and baselines/quadratic_baseline.py:11-68 of the reference; pinned against
the imported reference by oracle/make_golden.py (fixtures tests/golden/ridge_*.npz) and tests/test_oracle.py.
"""
import copy

import numpy as np


def features(paths, kind):
    """linear_baseline.py:11-36 (kind 0) / quadratic_baseline.py:11-43 (kind 1): float64 feature matrix."""
    o = np.concatenate([p["observations"] for p in paths])
    o = np.clip(o, -10, 10) / 10.0
    if o.ndim > 2:
        o = o.reshape(o.shape[0], -1)
    N, n = o.shape
    cols = [o]
    if kind == 1:
        quad = [o[:, i] * o[:, j] for i in range(n) for j in range(i, n)]
        cols.append(np.stack(quad, axis=1))
    cols.append(np.ones((N, 1)))
    al = np.concatenate([np.arange(len(p["rewards"])) / 1000.0 for p in paths])
    cols.append(np.stack([al ** (j + 1) for j in range(4)], axis=1))
    return np.concatenate(cols, axis=1)


def fit(paths, kind, reg_coeff, coeffs_before=None):
    """linear_baseline.py:38-58: returns (coeffs, error_before, error_after)."""
    F = features(paths, kind)
    y = np.concatenate([p["returns"] for p in paths])
    pred = F.dot(coeffs_before) if coeffs_before is not None else np.zeros(y.shape)
    error_before = np.sum((y - pred) ** 2) / np.sum(y ** 2)
    reg = copy.deepcopy(reg_coeff)
    for _ in range(10):
        c = np.linalg.lstsq(F.T.dot(F) + reg * np.identity(F.shape[1]), F.T.dot(y), rcond=-1)[0]
        if not np.any(np.isnan(c)):
            break
        reg *= 10
    error_after = np.sum((y - F.dot(c)) ** 2) / np.sum(y ** 2)
    return c, error_before, error_after


def predict(path, kind, coeffs):
    """linear_baseline.py:60-63."""
    if coeffs is None:
        return np.zeros(len(path["rewards"]))
    return features([path], kind).dot(coeffs)


# ---- the synthetic regression batches of the ridge fixtures (shared by oracle/make_golden_ridge.py and the tests)
FIXTURE_CASES = {
    "pm": dict(obs_dim=6, act_dim=2, n_paths=12, horizon=50, ragged=True),
    "swim": dict(obs_dim=8, act_dim=2, n_paths=30, horizon=250, ragged=False),
}


def fixture_paths(cfg, seed):
    from oracle import npg_oracle as O
    paths = O.synthetic_paths(cfg["obs_dim"], cfg["act_dim"], cfg["n_paths"], cfg["horizon"], seed=seed, ragged=cfg["ragged"])
    for p in paths:                         # make the regression non-trivial: rewards depend on the observations
        o = p["observations"]
        p["rewards"] = p["rewards"] * 0.1 + o[:, 0] - 0.5 * o[:, 1] ** 2 + 0.3 * o[:, 0] * o[:, 2]
    return paths

"""CPU: the numpy restatement of the reference's Linear / Quadratic baselines (oracle/ridge_oracle.py) against fixtures
produced by the unmodified reference (oracle/make_golden_ridge.py), and against the reference itself where importable."""
import ast
import os

import numpy as np
import pytest

from oracle import npg_oracle as O
from oracle import ridge_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = {"linear": 0, "quadratic": 1}


def load(tag, name):
    g = dict(np.load(os.path.join(GOLDEN, "ridge_%s_%s.npz" % (tag, name)), allow_pickle=False))
    g["meta"] = ast.literal_eval(str(g["meta"]))
    return g


@pytest.mark.parametrize("name", ["pm", "swim"])
@pytest.mark.parametrize("tag", ["linear", "quadratic"])
def test_ridge_oracle_matches_reference_fixture(tag, name):
    g = load(tag, name)
    m = g["meta"]
    kind = KINDS[tag]
    coeffs = None
    for rnd, seed in enumerate(m["path_seeds"]):
        paths = RO.fixture_paths(RO.FIXTURE_CASES[name], seed)
        O.compute_returns(paths, m["gamma"])
        np.testing.assert_array_equal(np.concatenate([p["returns"] for p in paths]), g["returns%d" % rnd])
        base = np.concatenate([RO.predict(p, kind, coeffs) for p in paths])
        np.testing.assert_allclose(base, g["base%d" % rnd], rtol=1e-9, atol=1e-9)
        coeffs, eb, ea = RO.fit(paths, kind, m["reg_coeff"], coeffs)
        np.testing.assert_allclose([eb, ea], g["errs%d" % rnd], rtol=1e-9)
        pred = np.concatenate([RO.predict(p, kind, coeffs) for p in paths])
        np.testing.assert_allclose(pred, g["pred%d" % rnd], rtol=1e-8, atol=1e-8)
        # coefficients: the ridge system is ill-conditioned (reg 1e-5 / 1e-3); same LAPACK path, so they agree closely
        np.testing.assert_allclose(coeffs, g["coeffs%d" % rnd], rtol=1e-6, atol=1e-8)


def test_feature_layout():
    """Column order of the reference: [o | (o_i o_j, i <= j) | 1 | al al^2 al^3 al^4] (linear_baseline.py:19-36)."""
    p = dict(observations=np.array([[0.0, 20.0, -30.0], [1.0, 2.0, 3.0]]), rewards=np.zeros(2))
    F0 = RO.features([p], 0)
    assert F0.shape == (2, 8)
    np.testing.assert_allclose(F0[0], [0.0, 1.0, -1.0, 1.0, 0.0, 0.0, 0.0, 0.0])
    np.testing.assert_allclose(F0[1], [0.1, 0.2, 0.3, 1.0, 1e-3, 1e-6, 1e-9, 1e-12])
    F1 = RO.features([p], 1)
    assert F1.shape == (2, 3 + 6 + 5)
    np.testing.assert_allclose(F1[1, 3:9], [0.01, 0.02, 0.03, 0.04, 0.06, 0.09])

"""CPU: the host half of the device ridge baselines (mjrl_b200/baselines/linear_baseline.py) -- the reference's
np.linalg.lstsq retry loop and the error_before / error_after bookkeeping from the Gram matrix -- driven by a stand-in engine
that computes what csrc/ridge.cu computes (Gram of [F | y], predictions) with numpy in float64."""
import ast
import os

import numpy as np
import pytest

from mjrl_b200.baselines.linear_baseline import LinearBaseline
from mjrl_b200.baselines.quadratic_baseline import QuadraticBaseline
from mjrl_b200.utils.gym_env import EnvSpec
from oracle import npg_oracle as O
from oracle import ridge_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class NumpyEngine:
    """The slice of mjrl_b200.engine.Engine the ridge baselines use."""

    def __init__(self, paths):
        self.paths = paths
        self.n = sum(len(p["rewards"]) for p in paths)
        self._base = np.zeros(self.n, np.float32)

    def ridge_gram(self, kind):
        F = RO.features(self.paths, kind)
        y = np.concatenate([p["returns"] for p in self.paths])
        return F.T.dot(F), F.T.dot(y), float(y.dot(y))

    def ridge_predict(self, kind, coeffs, want_sq_err=False):
        pred = RO.features(self.paths, kind).dot(coeffs)
        self._base = pred.astype(np.float32)
        if want_sq_err:
            y = np.concatenate([p["returns"] for p in self.paths])
            return float(np.sum((y - pred) ** 2))

    def set_baseline(self, b):
        self._base = np.asarray(b, np.float32)

    def baseline(self):
        return self._base


@pytest.mark.parametrize("name", ["pm", "swim"])
@pytest.mark.parametrize("tag,cls", [("linear", LinearBaseline), ("quadratic", QuadraticBaseline)])
def test_fit_and_predict_resident_match_reference_fixture(tag, cls, name):
    g = dict(np.load(os.path.join(GOLDEN, "ridge_%s_%s.npz" % (tag, name)), allow_pickle=False))
    m = ast.literal_eval(str(g["meta"]))
    cfg = RO.FIXTURE_CASES[name]
    bl = cls(EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"]))
    assert bl._reg_coeff == m["reg_coeff"] and bl._coeffs is None
    for rnd, seed in enumerate(m["path_seeds"]):
        paths = RO.fixture_paths(cfg, seed)
        O.compute_returns(paths, m["gamma"])
        eng = NumpyEngine(paths)
        bl.predict_resident(eng)                               # pre-fit baseline of the round (zeros before the first fit)
        np.testing.assert_allclose(eng.baseline(), g["base%d" % rnd].astype(np.float32), rtol=1e-6, atol=1e-6)
        eb, ea = bl.fit_resident(eng, return_errors=True)      # errors from the Gram matrix, not from predictions
        np.testing.assert_allclose([eb, ea], g["errs%d" % rnd], rtol=1e-8)
        np.testing.assert_allclose(bl._coeffs, g["coeffs%d" % rnd], rtol=1e-6, atol=1e-8)
        assert bl.fit_resident(eng) is None                    # return_errors=False returns nothing, like the reference


def test_regulariser_is_raised_while_the_solution_has_nans(monkeypatch):
    """linear_baseline.py:48-56: up to ten attempts, reg_coeff x10 after each NaN solution; self._reg_coeff is not changed."""
    bl = LinearBaseline(EnvSpec(3, 1, 10), reg_coeff=1e-5)
    regs = []
    real = np.linalg.lstsq

    def fake(a, b, rcond=None):
        regs.append(a[0, 0] - 1.0)                             # gram = identity below: the diagonal shows the regulariser
        if len(regs) < 3:
            return (np.full(b.shape, np.nan),)
        return real(a, b, rcond=rcond)

    monkeypatch.setattr(np.linalg, "lstsq", fake)
    bl._solve(np.identity(8), np.arange(8.0))
    np.testing.assert_allclose(regs, [1e-5, 1e-4, 1e-3], rtol=1e-6)
    assert bl._reg_coeff == 1e-5 and not np.isnan(bl._coeffs).any()


def test_constructor_and_unsupported_input():
    spec = EnvSpec(4, 2, 10)
    assert LinearBaseline(spec)._reg_coeff == 1e-5 and QuadraticBaseline(spec)._reg_coeff == 1e-3
    assert LinearBaseline(spec, inp_dim=7).n == 7 and QuadraticBaseline(spec).n == 4
    with pytest.raises(NotImplementedError):
        LinearBaseline(spec, inp='env_features')
    assert np.array_equal(LinearBaseline(spec).predict(dict(rewards=np.zeros(5))), np.zeros(5))

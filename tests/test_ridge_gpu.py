"""GPU: Linear / Quadratic baselines on the device-resident batch (csrc/ridge.cu) against fixtures of the
unmodified reference and against the numpy oracle, through the reference-facing classes and the C ABI."""
import ast
import os

import numpy as np
import pytest

from oracle import npg_oracle as O
from oracle import ridge_oracle as RO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = {"linear": 0, "quadratic": 1}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def load(tag, name):
    g = dict(np.load(os.path.join(GOLDEN, "ridge_%s_%s.npz" % (tag, name)), allow_pickle=False))
    g["meta"] = ast.literal_eval(str(g["meta"]))
    return g


@pytest.mark.parametrize("name", ["pm", "swim"])
@pytest.mark.parametrize("tag", ["linear", "quadratic"])
def test_gram_and_predictions_vs_oracle(tag, name, cuda_device):
    """The Gram pass is exact float64 arithmetic on the fp32-resident observations: compare with the oracle's feature
    matrix built from the same rounded observations (rel <= 1e-12), then predictions / squared error for given c."""
    from mjrl_b200.engine import Engine
    kind = KINDS[tag]
    cfg = RO.FIXTURE_CASES[name]
    paths = RO.fixture_paths(cfg, 0)
    O.compute_returns(paths, 0.995)
    n = sum(len(p["rewards"]) for p in paths)
    eng = Engine(cfg["obs_dim"], cfg["act_dim"], (32, 32), max_samples=n + 8, max_paths=len(paths) + 1)
    eng.upload_paths(paths)
    eng.compute_returns(0.995)
    rounded = [dict(p, observations=p["observations"].astype(np.float32).astype(np.float64)) for p in paths]
    F = RO.features(rounded, kind)
    y = np.concatenate([p["returns"] for p in paths])
    G, b, yy = eng.ridge_gram(kind)
    assert G.shape == (F.shape[1], F.shape[1]) == (eng.ridge_features(kind),) * 2
    assert rel(G, F.T.dot(F)) < 1e-12 and rel(b, F.T.dot(y)) < 1e-12 and abs(yy / y.dot(y) - 1) < 1e-12
    assert np.array_equal(G, G.T)
    G2, b2, yy2 = eng.ridge_gram(kind)                       # fixed summation order: bit-identical repeat
    assert np.array_equal(G, G2) and np.array_equal(b, b2) and yy == yy2
    c = np.random.RandomState(3).randn(F.shape[1])
    sq = eng.ridge_predict(kind, c, want_sq_err=True)
    pred = F.dot(c)
    np.testing.assert_allclose(eng.baseline(), pred.astype(np.float32), rtol=2e-7, atol=1e-6)
    assert abs(sq / np.sum((y - pred) ** 2) - 1) < 1e-10
    eng.close()


@pytest.mark.parametrize("name", ["pm", "swim"])
@pytest.mark.parametrize("tag", ["linear", "quadratic"])
def test_baseline_classes_vs_reference_fixture(tag, name, cuda_device):
    """Two consecutive rounds of compute_returns -> compute_advantages (pre-fit baseline) -> fit, exactly as the generator
    ran the reference.  The device holds fp32 observations (the reference works on the float64 originals), so the gates
    are on what the baseline is used for: predictions / advantages rel-L2 <= 2e-6, errors rel <= 1e-5."""
    from mjrl_b200 import runtime
    from mjrl_b200.baselines.linear_baseline import LinearBaseline
    from mjrl_b200.baselines.quadratic_baseline import QuadraticBaseline
    from mjrl_b200.utils import process_samples
    from mjrl_b200.utils.gym_env import EnvSpec
    g = load(tag, name)
    m = g["meta"]
    cfg = RO.FIXTURE_CASES[name]
    bl = (LinearBaseline if tag == "linear" else QuadraticBaseline)(EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"]))
    assert bl._reg_coeff == m["reg_coeff"]
    for rnd, seed in enumerate(m["path_seeds"]):
        paths = RO.fixture_paths(cfg, seed)
        process_samples.compute_returns(paths, m["gamma"])
        np.testing.assert_array_equal(np.concatenate([p["returns"] for p in paths]), g["returns%d" % rnd])
        process_samples.compute_advantages(paths, bl, m["gamma"], m["lam"])
        base = np.concatenate([p["baseline"] for p in paths])
        adv = np.concatenate([p["advantages"] for p in paths])
        if rnd == 0:
            assert not base.any()
        assert rel(base, g["base%d" % rnd]) < 2e-6 and rel(adv, g["adv%d" % rnd]) < 2e-6
        errs = bl.fit(paths, return_errors=True)
        np.testing.assert_allclose(errs, g["errs%d" % rnd], rtol=1e-5)
        pred = np.concatenate([bl.predict(p) for p in paths])
        assert rel(pred, g["pred%d" % rnd]) < 2e-6
        assert bl._coeffs.shape == g["coeffs%d" % rnd].shape
    runtime.shutdown()


def test_agent_with_linear_baseline(cuda_device):
    """NPG.train_step's post-rollout half with a LinearBaseline: the fit goes through fit_resident (no second upload) and
    the logged errors equal a stand-alone fit of the same batch."""
    from mjrl_b200 import runtime
    from mjrl_b200.algos.npg_cg import NPG
    from mjrl_b200.baselines.linear_baseline import LinearBaseline
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    cfg = RO.FIXTURE_CASES["swim"]
    spec = EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"])
    pol = MLP(spec, hidden_sizes=(32, 32), seed=500)
    bl = LinearBaseline(spec)
    agent = NPG(None, pol, bl, normalized_step_size=0.05, seed=1, save_logs=True)
    paths = RO.fixture_paths(cfg, 0)
    n = sum(len(p["rewards"]) for p in paths)
    eng = agent._eng(n, len(paths))
    up0 = eng.transfer_stats()[2]
    agent.update_from_paths([dict(p) for p in paths], 0.995, 0.97)
    assert eng.transfer_stats()[2] == up0 + 1                       # one trajectory upload for the whole step
    ref = LinearBaseline(spec)
    p2 = RO.fixture_paths(cfg, 0)
    O.compute_returns(p2, 0.995)
    eb, ea = ref.fit(p2, return_errors=True)
    assert abs(agent.logger.get_current_log()["VF_error_after"] / ea - 1) < 1e-9
    assert agent.logger.get_current_log()["VF_error_before"] == 1.0
    np.testing.assert_allclose(bl._coeffs, ref._coeffs, rtol=1e-9, atol=1e-12)
    runtime.shutdown()

"""GPU tests of the drop-in layer: classes with the reference's names/signatures (NPG / TRPO / DAPG, MLP /
LinearPolicy, MLPBaseline, process_samples) driving the CUDA engine, checked against the golden fixtures the
real reference produced and against the CPU oracle on the same trajectories."""
import copy
import pickle

import numpy as np
import pytest

from conftest import golden_paths, load_golden, one_minus_cos, rel
from oracle import npg_oracle as O

pytestmark = pytest.mark.gpu


def build(g, algo="npg", **kw):
    from mjrl_b200.algos.dapg import DAPG
    from mjrl_b200.algos.npg_cg import NPG
    from mjrl_b200.algos.trpo import TRPO
    from mjrl_b200.baselines.mlp_baseline import MLPBaseline
    from mjrl_b200.policies.gaussian_linear import LinearPolicy
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    m = g["meta"]
    es = EnvSpec(m["obs_dim"], m["act_dim"], m["horizon"])
    pol = LinearPolicy(es, seed=m["policy_seed"]) if len(m["hidden"]) == 0 else MLP(es, hidden_sizes=m["hidden"], seed=m["policy_seed"])
    assert np.array_equal(pol.get_param_values(), g["theta0"])      # same init draws as the reference
    bl = MLPBaseline(es, reg_coef=1e-3, batch_size=64, epochs=2, learn_rate=1e-3)
    bl.set_flat_weights(g["vf_w0"])
    cls = dict(npg=NPG, trpo=TRPO, dapg=DAPG)[algo]
    return cls(None, pol, bl, FIM_invert_args={"iters": m["cg_iters"], "damping": m["damping"]}, save_logs=True, **kw), pol, bl


@pytest.mark.parametrize("case", ["swim_40x250", "cheetah_24x500", "linear_30x200", "pm_5x50"])
def test_process_samples_and_npg_step(case, cuda_device):
    from mjrl_b200.utils import process_samples
    g = load_golden(case)
    m = g["meta"]
    paths = golden_paths(g)
    agent, pol, bl = build(g, "npg", normalized_step_size=m["npg_step"])
    process_samples.compute_returns(paths, m["gamma"])
    process_samples.compute_advantages(paths, bl, m["gamma"], m["lam"])
    cat = lambda k: np.concatenate([p[k] for p in paths])
    assert np.array_equal(cat("returns"), g["returns"])
    np.testing.assert_allclose(cat("baseline"), g["baseline"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(cat("advantages"), g["advantages"], rtol=0, atol=2e-4)
    assert paths[3]["advantages"].shape == paths[3]["rewards"].shape      # indexing: sample k of path i
    th0 = g["theta0"]
    stats = agent.train_from_paths(paths)
    np.testing.assert_allclose(stats, g["base_stats"], rtol=1e-12)
    new = pol.get_param_values()
    wc = eng_well_conditioned(g)
    assert one_minus_cos(new - th0, g["npg_theta"] - th0) < (1e-4 if wc else 5e-3)
    if wc:
        assert rel(new, g["npg_theta"]) < 2e-4
    old = np.concatenate([p.data.numpy().ravel() for p in pol.old_params])
    assert np.array_equal(old, new)                                    # set_new & set_old (npg_cg.py:142)
    log = agent.logger.get_current_log()
    for k in ("alpha", "delta", "time_vpg", "time_npg", "kl_dist", "surr_improvement", "running_score",
              "stoc_pol_mean", "stoc_pol_std", "stoc_pol_max", "stoc_pol_min"):
        assert k in log
    assert abs(log["alpha"] / g["npg_alpha"] - 1) < (5e-3 if wc else 5e-2)
    # the policy object stays a picklable CPU object with current weights
    pol2 = pickle.loads(pickle.dumps(pol))
    assert np.array_equal(pol2.get_param_values(), new)
    a, info = pol2.get_action(np.zeros(m["obs_dim"]))
    assert a.shape == (m["act_dim"],) and set(info) == {"mean", "log_std", "evaluation"}


def eng_well_conditioned(g):
    n = int(g["path_len"].sum())
    return n * g["meta"]["act_dim"] > 4 * g["theta0"].shape[0] or len(g["meta"]["hidden"]) == 0


@pytest.mark.parametrize("case", ["swim_40x250", "linear_30x200"])
def test_update_from_paths_matches_oracle_pipeline(case, cuda_device):
    """Whole post-rollout step (returns -> GAE -> NPG -> baseline fit) twice in a row, against the oracle run
    on the CPU with the same host RNG stream for the fit permutations."""
    g = load_golden(case)
    m = g["meta"]
    agent, pol, bl = build(g, "npg", normalized_step_size=m["npg_step"])
    spec = O.PolicySpec(m["obs_dim"], m["act_dim"], m["hidden"])
    vf = O.VFState(m["obs_dim"], (128, 128))
    vf.w = g["vf_w0"].copy()
    theta = g["theta0"].copy()
    for it in range(2):
        paths = O.synthetic_paths(m["obs_dim"], m["act_dim"], m["n_paths"], m["horizon"], seed=20 + it, ragged=True)
        ref_paths = copy.deepcopy(paths)
        np.random.seed(100 + it)
        agent.update_from_paths(paths, m["gamma"], m["lam"])
        # oracle
        np.random.seed(100 + it)
        O.compute_returns(ref_paths, m["gamma"])
        O.compute_advantages(ref_paths, lambda p: O.vf_predict(vf, p), m["gamma"], m["lam"])
        obs = np.concatenate([p["observations"] for p in ref_paths])
        act = np.concatenate([p["actions"] for p in ref_paths])
        adv = O.whiten(np.concatenate([p["advantages"] for p in ref_paths]))
        out = O.policy_update(spec, theta, obs, act, adv, "npg", step_size=m["npg_step"], cg_iters=m["cg_iters"],
                              damping=m["damping"])
        n = obs.shape[0]
        perms = [np.random.permutation(n) for _ in range(2)]
        errs = O.vf_fit(vf, ref_paths, perms, 2, 64, 1e-3, 1e-3, return_errors=True)
        new = pol.get_param_values()
        assert one_minus_cos(new - theta, out["new_params"] - theta) < 1e-4
        assert rel(new, out["new_params"]) < 5e-4
        log = agent.logger.get_current_log()
        np.testing.assert_allclose([log["VF_error_before"], log["VF_error_after"]], errs, rtol=2e-3)
        np.testing.assert_allclose(np.concatenate([p["advantages"] for p in paths]),
                                   np.concatenate([p["advantages"] for p in ref_paths]), rtol=0, atol=5e-4)
        assert bl.adam_step == vf.t
        theta = out["new_params"]
        # keep the two chains from drifting apart through fp32 noise amplified by the next CG solve
        pol.set_param_values(theta, True, True)
        agent._pushed = None
    pred = bl.predict(paths[0])
    np.testing.assert_allclose(pred, O.vf_predict(vf, ref_paths[0]), rtol=0, atol=5e-3)


def test_trpo_and_dapg_classes(cuda_device):
    g = load_golden("swim_40x250")
    m = g["meta"]
    th0 = g["theta0"]
    for kl, tag in ((0.01, "trpo"), (0.5, "trpo_big")):
        paths = golden_paths(g)
        for p, a in zip(paths, np.split(g["advantages"], np.cumsum(g["path_len"])[:-1])):
            p["advantages"] = a
        agent, pol, bl = build(g, "trpo", kl_dist=kl)
        agent.train_from_paths(paths)
        assert agent.last_step.backtracks == int(g[tag + "_backtracks"])
        assert one_minus_cos(pol.get_param_values() - th0, g[tag + "_theta"] - th0) < 1e-4
    paths = golden_paths(g)
    for p, a in zip(paths, np.split(g["advantages"], np.cumsum(g["path_len"])[:-1])):
        p["advantages"] = a
    demo = golden_paths(g, demo=True)
    agent, pol, bl = build(g, "dapg", demo_paths=demo, kl_dist=0.01, lam_0=1.0, lam_1=0.95)
    agent.iter_count = 3.0
    agent.train_from_paths(paths)
    assert agent.iter_count == 4.0
    assert one_minus_cos(pol.get_param_values() - th0, g["dapg_theta"] - th0) < 1e-4
    assert abs(agent.last_step.alpha / g["dapg_alpha"] - 1) < 2e-3


def test_reference_signature_helpers(cuda_device):
    """CPI_surrogate / kl_old_new / flat_vpg / HVP / build_Hvp_eval take concatenated arrays like the reference."""
    from mjrl_b200.utils.cg_solve import cg_solve
    g = load_golden("swim_40x250")
    paths = golden_paths(g)
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    agent, pol, bl = build(g, "npg")
    assert abs(agent.CPI_surrogate(obs, act, g["adv_white"]) - g["surr0"]) < 1e-6
    vpg = agent.flat_vpg(obs, act, g["adv_white"])
    assert rel(vpg, g["vpg"]) < 1e-5
    assert rel(agent.HVP(obs, act, g["fvp_vec"], g["meta"]["damping"]), g["fvp_out"]) < 1e-5
    hvp = agent.build_Hvp_eval([obs, act], regu_coef=g["meta"]["damping"])
    x = cg_solve(hvp, g["vpg"], x_0=g["vpg"].copy(), cg_iters=g["meta"]["cg_iters"])
    assert one_minus_cos(x, g["cg_x"]) < 1e-6
    pol.set_param_values(g["theta_pert"], set_new=True, set_old=False)
    assert abs(agent.kl_old_new(obs, act) / g["kl_pert"] - 1) < 1e-3
    assert abs(agent.CPI_surrogate(obs, act, g["adv_white"]) - g["surr_pert"]) < 1e-3 * max(1, abs(g["surr_pert"]))


def test_hvp_subsample_uses_global_rng(cuda_device):
    g = load_golden("swim_40x250")
    paths = golden_paths(g)
    for p, a in zip(paths, np.split(g["advantages"], np.cumsum(g["path_len"])[:-1])):
        p["advantages"] = a
    agent, pol, bl = build(g, "npg", normalized_step_size=g["meta"]["npg_step"], hvp_sample_frac=0.5)
    np.random.seed(77)                                  # the fixture drew its indices from this state
    agent.train_from_paths(paths)
    assert one_minus_cos(pol.get_param_values() - g["theta0"], g["sub_theta"] - g["theta0"]) < 1e-4


def test_foreign_baseline_object(cuda_device):
    """Any object with predict(path)/fit(paths) keeps working (LinearBaseline & co. stay on the host)."""
    from mjrl_b200.utils import process_samples

    class Zero:
        def predict(self, path):
            return np.zeros(len(path["rewards"]))

    paths = O.synthetic_paths(4, 2, 5, 30, seed=1, ragged=True)
    process_samples.compute_returns(paths, 0.9)
    process_samples.compute_advantages(paths, Zero(), 0.9, 0.8)
    for p in paths:
        assert np.array_equal(p["returns"], O.discount_sum(p["rewards"], 0.9))
        assert np.array_equal(p["advantages"], O.gae_path(p["rewards"], np.zeros(len(p["rewards"]), np.float32),
                                                          p["terminated"], 0.9, 0.8))

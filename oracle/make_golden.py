"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL mjrl reference.

Run in the build container (needs /root/reference or $MJRL_REF):
    python -m oracle.make_golden
Every array written here is an output of unmodified reference code (imported through
oracle/ref_shim.py) on deterministic synthetic trajectories (oracle.npg_oracle.synthetic_paths).
The fixtures are what pins both the oracle restatement and the CUDA engine on the GPU box, where
the reference itself is not available.
"""
import contextlib
import copy
import io
import os

import numpy as np
import torch

from oracle import npg_oracle as O
from oracle import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: obs, act, hidden, n_paths, horizon, ragged
    "pm_5x50": dict(obs_dim=6, act_dim=2, hidden=(32, 32), n_paths=5, horizon=50, ragged=False),
    "pm_40x25_ragged": dict(obs_dim=6, act_dim=2, hidden=(32, 32), n_paths=40, horizon=25, ragged=True),
    "swim_40x250": dict(obs_dim=8, act_dim=2, hidden=(64, 64), n_paths=40, horizon=250, ragged=True),
    "cheetah_24x500": dict(obs_dim=17, act_dim=6, hidden=(128, 128), n_paths=24, horizon=500, ragged=True),
    "linear_30x200": dict(obs_dim=40, act_dim=5, hidden=(), n_paths=30, horizon=200, ragged=True),
}
GAMMA, LAM, DAMPING, CG_ITERS, POLICY_SEED = 0.995, 0.97, 1e-4, 10, 500


def flat_params(module):
    return np.concatenate([p.data.numpy().ravel() for p in module.parameters()]).astype(np.float32)


def make_policy(R, cfg, spec_env):
    if len(cfg["hidden"]) == 0:
        return R.LinearPolicy(spec_env, seed=POLICY_SEED)
    return R.MLP(spec_env, hidden_sizes=cfg["hidden"], seed=POLICY_SEED)


def run_case(R, name, cfg):
    out = {}
    spec_env = R.EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"])
    paths = O.synthetic_paths(cfg["obs_dim"], cfg["act_dim"], cfg["n_paths"], cfg["horizon"], seed=0,
                              ragged=cfg["ragged"])
    out["path_len"] = np.array([len(p["rewards"]) for p in paths], np.int32)
    out["terminated"] = np.array([p["terminated"] for p in paths], np.uint8)
    # inputs are NOT stored: tests regenerate them with synthetic_paths(seed=0) and check these sums
    N = int(out["path_len"].sum())
    out["input_checksum"] = np.array([np.concatenate([p[k].ravel() for p in paths]).sum()
                                      for k in ("observations", "actions", "rewards")])

    policy = make_policy(R, cfg, spec_env)
    torch.manual_seed(11)
    baseline = R.MLPBaseline(spec_env, reg_coef=1e-3, batch_size=64, epochs=2, learn_rate=1e-3)
    out["theta0"] = policy.get_param_values()
    out["vf_w0"] = flat_params(baseline.model)

    # ---- returns / GAE (process_samples.py) --------------------------------------------------
    R.process_samples.compute_returns(paths, GAMMA)
    R.process_samples.compute_advantages(paths, baseline, GAMMA, LAM)
    out["returns"] = np.concatenate([p["returns"] for p in paths])
    out["baseline"] = np.concatenate([p["baseline"] for p in paths])
    out["advantages"] = np.concatenate([p["advantages"] for p in paths])
    p2 = copy.deepcopy(paths)
    R.process_samples.compute_advantages(p2, baseline, GAMMA, None)
    out["advantages_nogae"] = np.concatenate([p["advantages"] for p in p2])

    agent = R.NPG(None, policy, baseline, normalized_step_size=0.05, seed=123,
                  FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING})
    obs, act, adv, base_stats, _ = agent.process_paths(paths)
    out["adv_white"] = adv
    out["base_stats"] = np.array(base_stats)

    # ---- surrogate / KL / VPG / FVP ----------------------------------------------------------
    out["surr0"] = agent.CPI_surrogate(obs, act, adv).data.numpy().ravel()[0]
    out["vpg"] = agent.flat_vpg(obs, act, adv)
    v = np.random.RandomState(1).randn(policy.d).astype(np.float32)
    out["fvp_vec"] = v
    out["fvp_out"] = agent.HVP(obs, act, v, DAMPING)
    hvp = agent.build_Hvp_eval([obs, act], regu_coef=DAMPING)
    out["cg_x"] = R.cg_solve(hvp, out["vpg"], x_0=out["vpg"].copy(), cg_iters=CG_ITERS)
    # evaluation at a perturbed parameter vector (new != old)
    theta_p = out["theta0"] + 0.01 * np.random.RandomState(2).randn(policy.d).astype(np.float32)
    policy.set_param_values(theta_p, set_new=True, set_old=False)
    out["theta_pert"] = policy.get_param_values()
    out["surr_pert"] = agent.CPI_surrogate(obs, act, adv).data.numpy().ravel()[0]
    out["kl_pert"] = agent.kl_old_new(obs, act).data.numpy().ravel()[0]
    out["vpg_pert"] = agent.flat_vpg(obs, act, adv)
    policy.set_param_values(out["theta0"], set_new=True, set_old=True)

    # ---- NPG step ----------------------------------------------------------------------------
    agent.save_logs = True
    agent.logger = _Log()
    agent.train_from_paths(paths)
    out["npg_theta"] = policy.get_param_values()
    for k in ("alpha", "delta", "kl_dist", "surr_improvement"):
        out["npg_" + k] = np.float64(agent.logger.kv[k])

    # ---- TRPO step (kl_dist small enough to force backtracking on some cases) -----------------
    for tag, kl in (("trpo", 0.01), ("trpo_big", 0.5)):
        pol = make_policy(R, cfg, spec_env)
        tr = R.TRPO(None, pol, baseline, kl_dist=kl, FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING},
                    save_logs=True)
        tr.logger = _Log()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            tr.train_from_paths(paths)
        out[tag + "_theta"] = pol.get_param_values()
        out[tag + "_backtracks"] = np.int32(buf.getvalue().count("Backtracking"))
        for k in ("alpha", "kl_dist", "surr_improvement"):
            out[tag + "_" + k] = np.float64(tr.logger.kv[k])

    # ---- DAPG step (25% extra demo samples) ---------------------------------------------------
    demo = O.synthetic_paths(cfg["obs_dim"], cfg["act_dim"], max(2, cfg["n_paths"] // 4), cfg["horizon"], seed=5)
    out["demo_len"] = np.array([len(p["rewards"]) for p in demo], np.int32)
    pol = make_policy(R, cfg, spec_env)
    dg = R.DAPG(None, pol, baseline, demo_paths=demo, kl_dist=0.01, lam_0=1.0, lam_1=0.95,
                FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING}, save_logs=True)
    dg.logger = _Log()
    dg.iter_count = 3.0
    dg.train_from_paths(paths)
    out["dapg_theta"] = pol.get_param_values()
    out["dapg_alpha"] = np.float64(dg.logger.kv["alpha"])
    out["dapg_kl_dist"] = np.float64(dg.logger.kv["kl_dist"])

    # ---- hvp_sample_frac < 1: the reference draws np.random.choice per FVP (npg_cg.py:65-69) ----
    pol = make_policy(R, cfg, spec_env)
    sub = R.NPG(None, pol, baseline, normalized_step_size=0.05, hvp_sample_frac=0.5,
                FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING})
    np.random.seed(77)
    st = np.random.get_state()
    sub.train_from_paths(paths)
    np.random.set_state(st)
    out["sub_idx"] = np.stack([np.random.choice(N, size=int(0.5 * N)) for _ in range(CG_ITERS)]).astype(np.int32)
    out["sub_theta"] = pol.get_param_values()

    # ---- baseline fit: two consecutive calls (Adam state persists, mlp_baseline.py:33) ---------
    if N >= 128:
        np.random.seed(7)
        st = np.random.get_state()
        e = baseline.fit(paths, return_errors=True)
        out["fit1_err"] = np.array(e, np.float64)
        out["fit1_w"] = flat_params(baseline.model)
        baseline.fit(paths)
        out["fit2_w"] = flat_params(baseline.model)
        np.random.set_state(st)
        out["fit_perms"] = np.stack([np.random.permutation(N) for _ in range(4)]).astype(np.int32)
        ostate = baseline.optimizer.state
        params = list(baseline.model.parameters())
        out["fit2_m"] = np.concatenate([ostate[p]["exp_avg"].numpy().ravel() for p in params])
        out["fit2_v"] = np.concatenate([ostate[p]["exp_avg_sq"].numpy().ravel() for p in params])
        out["fit2_step"] = np.int64(int(ostate[params[0]]["step"]))
        out["fit2_predict"] = np.concatenate([baseline.predict(p) for p in paths])

    out["meta"] = np.array(repr(dict(cfg, gamma=GAMMA, lam=LAM, damping=DAMPING, cg_iters=CG_ITERS,
                                     policy_seed=POLICY_SEED, baseline_seed=11, path_seed=0, demo_seed=5,
                                     npg_step=0.05, vf=dict(reg_coef=1e-3, batch_size=64, epochs=2, lr=1e-3))))
    return out


def run_inorm_case(R, cfg, alpha=0.7):
    """input_normalization (npg_cg.py:101-107): the reference blends the observation moments into
    policy.model ONLY -- old_model keeps its stale transforms, so LR != 1 and mu_new != mu_old at theta_new = theta_old.
    Every array is an output of the unmodified reference on the swim_40x250 trajectories."""
    out = {}
    spec_env = R.EnvSpec(cfg["obs_dim"], cfg["act_dim"], cfg["horizon"])
    paths = O.synthetic_paths(cfg["obs_dim"], cfg["act_dim"], cfg["n_paths"], cfg["horizon"], seed=0, ragged=cfg["ragged"])
    policy = make_policy(R, cfg, spec_env)
    torch.manual_seed(11)
    baseline = R.MLPBaseline(spec_env, reg_coef=1e-3, batch_size=64, epochs=2, learn_rate=1e-3)
    R.process_samples.compute_returns(paths, GAMMA)
    R.process_samples.compute_advantages(paths, baseline, GAMMA, LAM)
    agent = R.NPG(None, policy, baseline, normalized_step_size=0.05, seed=123, input_normalization=alpha,
                  FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING}, save_logs=True)
    agent.logger = _Log()
    out["theta0"] = policy.get_param_values()
    agent.train_from_paths(paths)
    m, om = policy.model, policy.old_model
    for k in ("in_shift", "in_scale", "out_shift", "out_scale"):
        out["new_" + k] = getattr(m, k).data.numpy().copy()
        out["old_" + k] = getattr(om, k).data.numpy().copy()
    out["theta1"] = policy.get_param_values()
    for k in ("alpha", "delta", "kl_dist", "surr_improvement"):
        out["npg_" + k] = np.float64(agent.logger.kv[k])
    # the pieces, re-evaluated at theta0 with the blended (new) / stale (old) transforms
    policy.set_param_values(out["theta0"], set_new=True, set_old=True)
    obs, act, adv, _, _ = agent.process_paths(paths)
    out["surr"] = agent.CPI_surrogate(obs, act, adv).data.numpy().ravel()[0]
    out["kl"] = agent.kl_old_new(obs, act).data.numpy().ravel()[0]
    out["vpg"] = agent.flat_vpg(obs, act, adv)
    v = np.random.RandomState(1).randn(policy.d).astype(np.float32)
    out["fvp_vec"] = v
    out["fvp_out"] = agent.HVP(obs, act, v, DAMPING)
    out["meta"] = np.array(repr(dict(cfg, gamma=GAMMA, lam=LAM, damping=DAMPING, cg_iters=CG_ITERS, policy_seed=POLICY_SEED,
                                     baseline_seed=11, path_seed=0, npg_step=0.05, input_normalization=alpha)))
    return out


class _Log:
    def __init__(self):
        self.kv = {}

    def log_kv(self, k, v):
        self.kv[k] = v


def kat():
    """Known-answer vectors from SURVEY.md section 8c, re-derived here by executing reference code."""
    R = ref_shim.load()
    ps = R.process_samples
    out = {}
    out["ds_in"] = np.array([1.0, 2.0, 3.0, 4.0])
    out["ds_out"] = ps.discount_sum(out["ds_in"], 0.5)
    out["ds_f32_out"] = ps.discount_sum(out["ds_in"].astype(np.float32), 0.5)
    r = np.array([1.0, 0.0, 2.0, -1.0])
    b = np.array([0.5, 1.0, -1.0, 2.0], np.float32)

    class B:
        def predict(self, path):
            return b

    for term in (False, True):
        p = [dict(rewards=r, terminated=term)]
        ps.compute_returns(p, 0.9)
        ps.compute_advantages(p, B(), 0.9, 0.5)
        out["gae_ret"] = p[0]["returns"]
        out["gae_adv_term%d" % term] = p[0]["advantages"]
    p = [dict(rewards=r, terminated=False)]
    ps.compute_returns(p, 0.9)
    ps.compute_advantages(p, B(), 0.9, None)
    out["nogae_adv"] = p[0]["advantages"]
    out["gae_r"], out["gae_b"] = r, b
    A = np.array([[4.0, 1.0], [1.0, 3.0]])
    bb = np.array([1.0, 2.0])
    out["cg_1"] = R.cg_solve(lambda v: A.dot(v), bb, x_0=np.array([9.0, 9.0]), cg_iters=1)
    out["cg_2"] = R.cg_solve(lambda v: A.dot(v), bb, cg_iters=2)
    bl = R.MLPBaseline(R.EnvSpec(3, 1, 10))
    out["feat_obs"] = np.array([[0.0, 20.0, -30.0], [1.0, 2.0, 3.0]])
    out["feat_out"] = bl._features([dict(observations=out["feat_obs"], rewards=np.zeros(2))])
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    R = ref_shim.load()
    torch.set_num_threads(1)          # bitwise-repeatable reference outputs
    np.savez_compressed(os.path.join(GOLDEN_DIR, "kat.npz"), **kat())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inorm_swim_40x250.npz"), **run_inorm_case(R, CASES["swim_40x250"]))
    for name, cfg in CASES.items():
        out = run_case(R, name, cfg)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-18s N=%-6d d=%-6d trpo_bt=%d/%d  %.0f KB" % (
            name, int(out["path_len"].sum()), out["theta0"].shape[0], out["trpo_backtracks"],
            out["trpo_big_backtracks"], os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

"""GPU tests of (1) batch residency -- every public call works on the arrays it is handed, like the reference
(algos/batch_reinforce.py:94-112 reads `paths` afresh on every train_step): a fresh list over the same arrays, an
in-place-mutated batch and recycled object addresses all re-upload; (2) the observation / action transforms of the
MLP policy on the device kernels (utils/fc_network.py:27-51) and `input_normalization` (algos/npg_cg.py:101-107,
SURVEY A10) against a fixture produced by the unmodified reference."""
import copy

import numpy as np
import pytest

from conftest import golden_paths, load_golden, one_minus_cos, rel
from oracle import npg_oracle as O
from test_agents_gpu import build

pytestmark = pytest.mark.gpu


def _uploads(eng):
    return eng.transfer_stats()[2]


def test_update_from_paths_always_uploads(cuda_device):
    g = load_golden("swim_40x250")
    m = g["meta"]
    paths = golden_paths(g)
    agent, pol, bl = build(g, "npg", normalized_step_size=m["npg_step"])
    np.random.seed(0)
    agent.update_from_paths(paths, m["gamma"], m["lam"])
    eng = agent._engine
    u0, h0 = _uploads(eng), eng.transfer_stats()[0]
    n = sum(len(p["rewards"]) for p in paths)
    traj_bytes = n * ((m["obs_dim"] + m["act_dim"]) * 4 + 8)      # obs / act cross PCIe as fp32, rewards as fp64
    # (a) the SAME list object again: uploaded again
    agent.update_from_paths(paths, m["gamma"], m["lam"])
    assert _uploads(eng) == u0 + 1
    assert eng.transfer_stats()[0] - h0 >= traj_bytes
    # (b) a fresh list over the same arrays -- the loop CPython recycles list addresses for -- 13 times in a row
    for _ in range(13):
        before = _uploads(eng)
        fresh = [dict(p) for p in paths]
        agent.update_from_paths(fresh, m["gamma"], m["lam"])
        del fresh
        assert _uploads(eng) == before + 1
    # (c) the arrays mutated IN PLACE (same list, same array objects): the new contents are what the step sees
    ret_before = [p["returns"].copy() for p in paths]
    for p in paths:
        p["rewards"] *= 2.0                                   # exact in floating point
    agent.update_from_paths(paths, m["gamma"], m["lam"])
    for p, r0 in zip(paths, ret_before):
        assert np.array_equal(p["returns"], 2.0 * r0)
    assert eng.session_paths is None                          # the pin is dropped when the call returns


def test_flat_helpers_read_their_arguments(cuda_device):
    """CPI_surrogate / flat_vpg / HVP on concatenated arrays: edited or recycled arrays are never served stale."""
    g = load_golden("swim_40x250")
    paths = golden_paths(g)
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    adv = g["adv_white"]
    agent, pol, bl = build(g, "npg")
    spec = O.PolicySpec(g["meta"]["obs_dim"], g["meta"]["act_dim"], g["meta"]["hidden"])
    th = g["theta0"]
    v1 = agent.flat_vpg(obs, act, adv)
    assert rel(v1, g["vpg"]) < 1e-5
    obs *= 0.5                                                # in place: same object, same id
    v2 = agent.flat_vpg(obs, act, adv)
    assert rel(v2, O.flat_vpg(spec, th, obs, act, adv)) < 1e-5 and rel(v2, v1) > 1e-3
    # 100 temporaries in a row (freed arrays hand their address to the next one)
    for i in range(100):
        o = obs + 0.01 * i
        got = agent.CPI_surrogate(o, act, adv)
        del o
        if i % 25 == 0:
            want = float(O.surrogate(spec, th, th, obs + 0.01 * i, act, adv))
            assert abs(got - want) < 1e-5 * max(1.0, abs(want))
    # build_Hvp_eval: one upload for the ten products of a CG solve, a new upload when the batch changes in between
    eng = agent._engine
    hvp = agent.build_Hvp_eval([obs, act], regu_coef=1e-4)
    u0 = _uploads(eng)
    a = hvp(g["fvp_vec"]); b = hvp(g["fvp_vec"])
    assert _uploads(eng) == u0 + 1 and np.array_equal(a, b)
    agent.flat_vpg(obs + 1.0, act, adv)                       # somebody else's batch replaces it
    c = hvp(g["fvp_vec"])
    assert _uploads(eng) == u0 + 3 and np.array_equal(a, c)


def test_process_paths_reads_host_advantages(cuda_device):
    """train_from_paths(paths) as a stand-alone call: path['advantages'] edited by the caller are the ones used."""
    g = load_golden("swim_40x250")
    m = g["meta"]
    paths = golden_paths(g)
    for p, a in zip(paths, np.split(g["advantages"], np.cumsum(g["path_len"])[:-1])):
        p["advantages"] = a.copy()
    agent, pol, bl = build(g, "npg", normalized_step_size=m["npg_step"])
    agent.train_from_paths(paths)
    th1 = pol.get_param_values()
    assert one_minus_cos(th1 - g["theta0"], g["npg_theta"] - g["theta0"]) < 1e-4
    # flipped advantages, in place, same objects: the step must flip too
    pol.set_param_values(g["theta0"], True, True)
    for p in paths:
        p["advantages"] *= -1.0
    agent.train_from_paths(paths)
    th2 = pol.get_param_values()
    assert one_minus_cos(th2 - g["theta0"], -(g["npg_theta"] - g["theta0"])) < 1e-4


def test_closed_engine_fails_loudly(cuda_device):
    from mjrl_b200.engine import Engine, MjbError
    eng = Engine(4, 2, (32, 32), max_samples=256, max_paths=8)
    eng.close()
    with pytest.raises(MjbError):
        eng.get_params()


@pytest.mark.parametrize("hidden", [(64, 64), (128, 128)], ids=["fma64", "tc128"])
def test_mlp_device_transforms(hidden, cuda_device):
    """Non-identity in/out transforms on the MLP tile kernels (FMA kernels at 64x64, tcgen05 FVP at 128x128): EVAL with
    different NEW and OLD transform sets, VPG and FVP with equal ones, against the oracle (fc_network.py:46-51)."""
    from mjrl_b200.engine import Engine
    rng = np.random.RandomState(5)
    obs_dim, act_dim = 11, 4
    paths = O.synthetic_paths(obs_dim, act_dim, 12, 300, seed=2, ragged=True)
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    adv = rng.randn(obs.shape[0]).astype(np.float32)
    tr_new = dict(in_shift=0.3 * rng.randn(obs_dim), in_scale=0.5 + rng.rand(obs_dim),
                  out_shift=0.1 * rng.randn(act_dim), out_scale=0.5 + rng.rand(act_dim))
    tr_old = dict(in_shift=0.2 * rng.randn(obs_dim), in_scale=0.7 + rng.rand(obs_dim),
                  out_shift=0.05 * rng.randn(act_dim), out_scale=0.6 + rng.rand(act_dim))
    f32 = lambda d: {k: v.astype(np.float32) for k, v in d.items()}
    spec_n = O.PolicySpec(obs_dim, act_dim, hidden, **f32(tr_new))
    spec_o = O.PolicySpec(obs_dim, act_dim, hidden, **f32(tr_old))
    th = O.init_policy_params(spec_n, 3)
    th[-act_dim:] = 0.1 * rng.randn(act_dim)
    th_old = (th + 0.02 * rng.randn(th.shape[0])).astype(np.float32)
    eng = Engine(obs_dim, act_dim, hidden, max_samples=obs.shape[0] + 8, max_paths=16)
    eng.upload_paths(paths)
    eng.set_white(adv)
    # ---- EVAL: new (theta, tr_new) against old (theta_old, tr_old)
    eng.set_transforms(**f32(tr_new), old=False)
    eng.set_transforms(**f32(tr_old), old=True)
    eng.set_params(th, True, False)
    eng.set_params(th_old, False, True)
    surr, kl = eng.eval()
    want_s = float(O.surrogate(spec_n, th, th_old, obs, act, adv, spec_old=spec_o))
    want_k = float(O.mean_kl(spec_n, th, th_old, obs, spec_old=spec_o))
    assert abs(surr - want_s) < 1e-4 * max(1.0, abs(want_s)), (surr, want_s)
    assert abs(kl / want_k - 1) < 1e-3, (kl, want_k)
    # ---- VPG / FVP with the same transforms on both sides (LR = 1)
    eng.set_transforms(**f32(tr_new), old=True)
    eng.set_params(th, True, True)
    assert rel(eng.vpg(), O.flat_vpg(spec_n, th, obs, act, adv)) < 1e-5
    v = rng.randn(th.shape[0]).astype(np.float32)
    want_f = O.fvp(spec_n, th, obs, v, 1e-4)
    tc = eng.set_tensor_cores(True)
    assert tc == (hidden == (128, 128))
    assert rel(eng.fvp(v, 1e-4), want_f) < 1e-5
    eng.set_tensor_cores(False)
    assert rel(eng.fvp(v, 1e-4), want_f) < 1e-5
    eng.close()


def test_input_normalization_against_reference_fixture(cuda_device):
    """NPG(input_normalization=0.7): transforms blended into policy.model only (old_model stays stale, A10), pieces and
    the resulting step against tests/golden/inorm_swim_40x250.npz (unmodified reference, oracle/make_golden.py)."""
    from mjrl_b200.algos.npg_cg import NPG
    from mjrl_b200.baselines.mlp_baseline import MLPBaseline
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    g = load_golden("inorm_swim_40x250")
    base = load_golden("swim_40x250")
    m = g["meta"]
    paths = golden_paths(base)
    for p, a in zip(paths, np.split(base["advantages"], np.cumsum(base["path_len"])[:-1])):
        p["advantages"] = a
    es = EnvSpec(m["obs_dim"], m["act_dim"], m["horizon"])
    pol = MLP(es, hidden_sizes=m["hidden"], seed=m["policy_seed"])
    bl = MLPBaseline(es, reg_coef=1e-3, batch_size=64, epochs=2, learn_rate=1e-3)
    bl.set_flat_weights(base["vf_w0"])
    agent = NPG(None, pol, bl, normalized_step_size=m["npg_step"], input_normalization=m["input_normalization"],
                FIM_invert_args={"iters": m["cg_iters"], "damping": m["damping"]}, save_logs=True)
    assert np.array_equal(pol.get_param_values(), g["theta0"])
    agent.train_from_paths(paths)
    for k in ("in_shift", "in_scale", "out_shift", "out_scale"):
        np.testing.assert_allclose(getattr(pol.model, k).numpy(), g["new_" + k], rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(getattr(pol.old_model, k).numpy(), g["old_" + k])      # stale, as in the reference
    th1 = pol.get_param_values()
    assert one_minus_cos(th1 - g["theta0"], g["theta1"] - g["theta0"]) < 1e-4
    log = agent.logger.get_current_log()
    assert abs(log["alpha"] / g["npg_alpha"] - 1) < 5e-3
    assert abs(log["kl_dist"] / g["npg_kl_dist"] - 1) < 2e-2
    # ---- the pieces at theta0 with the blended NEW and the stale OLD transforms
    pol.set_param_values(g["theta0"], True, True)
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    adv = base["adv_white"]
    assert abs(agent.CPI_surrogate(obs, act, adv) - g["surr"]) < 1e-6
    assert abs(agent.kl_old_new(obs, act) - g["kl"]) < 1e-7
    assert rel(agent.flat_vpg(obs, act, adv), g["vpg"]) < 1e-5
    # The engine's Fisher product is the Gauss-Newton form J^T W J with the NEW transforms; the reference differentiates
    # mean_kl(new, old) twice, which adds a term proportional to (mu_new - mu_old) that vanishes whenever new == old --
    # i.e. always, except under input_normalization.  Measured here: 6e-5 relative, 2e-9 in direction.
    f = agent.HVP(obs, act, g["fvp_vec"], m["damping"])
    assert rel(f, g["fvp_out"]) < 2e-4 and one_minus_cos(f, g["fvp_out"]) < 1e-7


def test_deferred_fit_is_joined_by_readers(cuda_device):
    """update_from_paths returns as soon as theta is back; the baseline fit keeps running on its stream (so the next
    batch's upload overlaps it) and every reader of the baseline joins it: pickling right after the call holds the
    POST-fit weights and Adam state -- identical to a run that waits for the fit explicitly."""
    import pickle
    g = load_golden("cheetah_24x500")
    m = g["meta"]

    def run(save_logs):
        paths = golden_paths(g)
        agent, pol, bl = build(g, "npg", normalized_step_size=m["npg_step"])
        agent.save_logs = save_logs
        np.random.seed(5)
        agent.update_from_paths(paths, m["gamma"], m["lam"])
        return agent, pol, bl

    agent, pol, bl = run(False)
    assert bl._fit_pending                                   # not joined yet
    blob = pickle.dumps(bl)                                  # __getstate__ joins
    assert not bl._fit_pending
    mid = pickle.loads(blob)
    w_mid, step_mid = mid.get_flat_weights(), mid.adam_step
    assert step_mid == (sum(len(p) for p in g["fit_perms"][:1]) // 64 - 1) * 2        # 2 epochs ran to completion
    assert not np.array_equal(w_mid, g["vf_w0"])
    agent2, pol2, bl2 = run(True)                            # save_logs: joins inside update_from_paths (VF_error_after)
    assert not bl2._fit_pending
    assert np.array_equal(bl2.get_flat_weights(), w_mid) and bl2.adam_step == step_mid
    assert np.array_equal(pol2.get_param_values(), pol.get_param_values())
    # a deferred fit followed by another update (upload overlaps the fit in flight) equals two joined updates
    for a in (agent, agent2):
        np.random.seed(6)
        a.update_from_paths(golden_paths(g), m["gamma"], m["lam"])
    assert np.array_equal(bl.get_flat_weights(), bl2.get_flat_weights())
    assert np.array_equal(pol.get_param_values(), pol2.get_param_values())


def test_update_from_rollouts_equals_update_from_paths(cuda_device):
    """Device-resident batched rollouts [n_traj, H, dim] (the learned-model hand-off of model_accel_npg.py:107-181) with
    per-trajectory prefix lengths and termination flags: same parameters, baseline weights and statistics -- bit for
    bit -- as the same trajectories arriving as host path dicts."""
    import torch
    from mjrl_b200.algos.trpo import TRPO
    from mjrl_b200.baselines.mlp_baseline import MLPBaseline
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    rng = np.random.RandomState(3)
    n_traj, H, od, ad = 37, 160, 17, 6
    obs, act, rew = rng.randn(n_traj, H, od), rng.randn(n_traj, H, ad), rng.randn(n_traj, H)
    lens = rng.randint(5, H + 1, size=n_traj).astype(np.int32)
    lens[:5] = H
    term = (lens < H).astype(np.uint8)
    es = EnvSpec(od, ad, H)

    def agent():
        pol = MLP(es, hidden_sizes=(128, 128), seed=11)
        torch.manual_seed(5)
        bl = MLPBaseline(es, reg_coef=1e-3, batch_size=64, epochs=1, learn_rate=1e-3)
        a = TRPO(None, pol, bl, kl_dist=0.01)
        a.verbose = False
        return a, pol, bl

    a1, p1, b1 = agent()
    paths = [dict(observations=obs[i, :lens[i]].copy(), actions=act[i, :lens[i]].copy(), rewards=rew[i, :lens[i]].copy(),
                  terminated=bool(term[i])) for i in range(n_traj)]
    np.random.seed(9)
    s1 = a1.update_from_paths(paths, 0.995, 0.97)
    w1 = b1.get_flat_weights()
    a2, p2, b2 = agent()
    dev = torch.device("cuda:0")
    roll = dict(observations=torch.from_numpy(obs).to(dev), actions=torch.from_numpy(act).to(dev), rewards=torch.from_numpy(rew).to(dev))
    h2d0 = None
    np.random.seed(9)
    eng_before = a2._eng(int(lens.sum()), n_traj)
    h2d0 = eng_before.transfer_stats()[0]
    s2 = a2.update_from_rollouts(roll, 0.995, 0.97, lengths=lens, terminated=term)
    moved = a2._engine.transfer_stats()[0] - h2d0
    assert moved < 2 * 1024 * 1024, moved                    # parameters / permutation only: the samples never cross PCIe
    assert np.array_equal(p1.get_param_values(), p2.get_param_values())
    assert np.array_equal(w1, b2.get_flat_weights())
    np.testing.assert_allclose(s1, s2, rtol=1e-12)
    assert a1.last_step.backtracks == a2.last_step.backtracks
    # float32 device tensors take the same path (rounded where the host staging would round)
    a3, p3, b3 = agent()
    np.random.seed(9)
    a3.update_from_rollouts({k: v.float() for k, v in roll.items()}, 0.995, 0.97, lengths=lens, terminated=term)
    th0 = MLP(es, hidden_sizes=(128, 128), seed=11).get_param_values()
    assert one_minus_cos(p3.get_param_values() - th0, p1.get_param_values() - th0) < 1e-4

"""GPU parity tests: the CUDA engine (through the C ABI) against the golden fixtures produced by the real
reference and against the CPU oracle on the same seeded inputs.

Tolerances (fp32 path; reference self-noise between 1 and 8 CPU threads is 2.5e-7 on the FVP and 3e-6 on
the CG direction):
  returns ........ bit-exact;  GAE: bit-exact given identical baseline predictions
  VPG / FVP ...... rel-L2 <= 1e-5
  CG direction ... 1 - cos <= 1e-4 (north_star), tighter where N >> d
  surrogate / KL / alpha ... rel <= 1e-3
"""
import numpy as np
import pytest

from conftest import ALL_CASES, golden_paths, load_golden, one_minus_cos, rel
from oracle import npg_oracle as O

pytestmark = pytest.mark.gpu


def make_engine(g, cuda_device, demo=0):
    from mjrl_b200.engine import Engine
    m = g["meta"]
    n = int(g["path_len"].sum())
    eng = Engine(m["obs_dim"], m["act_dim"], m["hidden"], max_samples=n + demo + 8, max_paths=len(g["path_len"]) + 1)
    eng.set_params(g["theta0"])
    eng.vf_set_state(g["vf_w0"], np.zeros_like(g["vf_w0"]), np.zeros_like(g["vf_w0"]), 0)
    return eng


def well_conditioned(g):
    n = int(g["path_len"].sum())
    d = g["theta0"].shape[0]
    return n * g["meta"]["act_dim"] > 4 * d or len(g["meta"]["hidden"]) == 0


@pytest.mark.parametrize("case", ALL_CASES)
def test_returns_gae(case, cuda_device):
    g = load_golden(case)
    paths = golden_paths(g)
    m = g["meta"]
    eng = make_engine(g, cuda_device)
    eng.upload_paths(paths)
    eng.compute_returns(m["gamma"])
    assert np.array_equal(eng.returns(), g["returns"])            # bit-exact fp64 scan
    eng.vf_predict()
    base = eng.baseline()
    np.testing.assert_allclose(base, g["baseline"], rtol=0, atol=5e-6)
    eng.compute_advantages(m["gamma"], m["lam"])
    adv = eng.advantages()
    # bit-exact against the reference formula evaluated on the engine's own baseline predictions
    k = 0
    for p in paths:
        T = len(p["rewards"])
        want = O.gae_path(p["rewards"], base[k:k + T], p["terminated"], m["gamma"], m["lam"])
        assert np.array_equal(adv[k:k + T], want)
        k += T
    np.testing.assert_allclose(adv, g["advantages"], rtol=0, atol=2e-4)
    eng.compute_advantages(m["gamma"], None)
    assert np.array_equal(eng.advantages(), g["returns"] - base.astype(np.float64))
    # whitening + return statistics, from the reference's own advantages
    eng.set_advantages(g["advantages"])
    st = eng.process_paths()
    np.testing.assert_allclose(eng.adv_white(), g["adv_white"].astype(np.float32), rtol=0, atol=1e-6)
    np.testing.assert_allclose([st.mean_return, st.std_return, st.min_return, st.max_return], g["base_stats"],
                               rtol=1e-12)
    eng.close()


@pytest.mark.parametrize("case", ALL_CASES)
def test_vpg_fvp_eval(case, cuda_device):
    g = load_golden(case)
    paths = golden_paths(g)
    eng = make_engine(g, cuda_device)
    eng.upload_paths(paths)
    eng.set_advantages(g["advantages"])
    eng.process_paths()
    vpg = eng.vpg()
    assert rel(vpg, g["vpg"]) < 1e-5
    fv = eng.fvp(g["fvp_vec"], g["meta"]["damping"])
    assert rel(fv, g["fvp_out"]) < 1e-5
    surr, kl = eng.eval()
    assert abs(surr - g["surr0"]) < 1e-6 and kl == 0.0
    # new != old: surrogate / KL / gradient at the perturbed parameters
    eng.set_params(g["theta_pert"], set_new=True, set_old=False)
    surr, kl = eng.eval()
    assert abs(surr - g["surr_pert"]) < 1e-3 * max(1.0, abs(g["surr_pert"]))
    assert abs(kl - g["kl_pert"]) < 1e-3 * max(1e-3, g["kl_pert"])
    assert rel(eng.vpg(), g["vpg_pert"]) < 2e-5
    # CG direction from the reference's own gradient
    eng.set_params(g["theta0"])
    x = eng.cg(g["vpg"], iters=g["meta"]["cg_iters"], damping=g["meta"]["damping"])
    assert one_minus_cos(x, g["cg_x"]) < (1e-6 if well_conditioned(g) else 5e-3)
    eng.close()


@pytest.mark.parametrize("case", ALL_CASES)
def test_policy_steps(case, cuda_device):
    g = load_golden(case)
    paths = golden_paths(g)
    m = g["meta"]
    th = g["theta0"]
    wc = well_conditioned(g)
    ctol, stol = (1e-4, 2e-3) if wc else (2e-2, 5e-2)
    demo = golden_paths(g, demo=True)
    eng = make_engine(g, cuda_device, demo=sum(len(p["rewards"]) for p in demo))

    def fresh():
        eng.upload_paths(paths)
        eng.set_params(th)
        eng.set_advantages(g["advantages"])
        eng.process_paths()

    fresh()
    st = eng.step("npg", step_size=m["npg_step"], cg_iters=m["cg_iters"], damping=m["damping"])
    new = eng.get_params()
    assert one_minus_cos(new - th, g["npg_theta"] - th) < ctol
    if wc:
        assert rel(new, g["npg_theta"]) < 1e-4
    assert abs(st.alpha / g["npg_alpha"] - 1) < stol
    assert abs(st.kl_dist / g["npg_kl_dist"] - 1) < 3 * stol
    assert abs((st.surr_after - st.surr_before) / g["npg_surr_improvement"] - 1) < 3 * stol
    assert np.array_equal(eng.get_params(old=True), new)
    for tag, kl in (("trpo", 0.01), ("trpo_big", 0.5)):
        fresh()
        st = eng.step("trpo", step_size=kl, cg_iters=m["cg_iters"], damping=m["damping"])
        # the accept test is KL < kl_dist; on the rank-deficient (N < d) fixtures the reference's own KL sits within
        # fp32-CG noise of the threshold (0.009957 vs 0.01 on pm_40x25_ragged), so there the count may differ by one
        borderline = (not wc) and abs(float(g[tag + "_kl_dist"]) / kl - 1) < 0.25
        if borderline:
            assert abs(st.backtracks - int(g[tag + "_backtracks"])) <= 1
        else:
            assert st.backtracks == int(g[tag + "_backtracks"])
        assert one_minus_cos(eng.get_params() - th, g[tag + "_theta"] - th) < max(ctol, 2e-2 if borderline else 0)
        if st.backtracks == int(g[tag + "_backtracks"]):
            assert abs(st.alpha / g[tag + "_alpha"] - 1) < stol
            assert abs(st.kl_dist / g[tag + "_kl_dist"] - 1) < 3 * stol
    fresh()
    eng.upload_paths(demo, which=1)
    st = eng.step("dapg", step_size=0.01, cg_iters=m["cg_iters"], damping=m["damping"], demo_lam=1.0 * 0.95 ** 3.0)
    assert one_minus_cos(eng.get_params() - th, g["dapg_theta"] - th) < ctol
    assert abs(st.alpha / g["dapg_alpha"] - 1) < stol
    fresh()
    st = eng.step("npg", step_size=m["npg_step"], cg_iters=m["cg_iters"], damping=m["damping"], hvp_idx=g["sub_idx"])
    assert one_minus_cos(eng.get_params() - th, g["sub_theta"] - th) < ctol
    eng.close()


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("case", ["pm_5x50", "pm_40x25_ragged", "swim_40x250", "cheetah_24x500", "linear_30x200"])
def test_baseline_fit(case, tensor_cores, cuda_device):
    """tensor_cores: the single-SM tcgen05 kernel (the default) / the single-CTA fp32-FMA kernel.  linear_30x200
    has 44 input features and cheetah_24x500 21: both widths of the tensor-core kernel's layer-1 K range."""
    g = load_golden(case)
    paths = golden_paths(g)
    eng = make_engine(g, cuda_device)
    eng.vf_set_tensor_cores(tensor_cores)
    eng.upload_paths(paths)
    eng.compute_returns(g["meta"]["gamma"])
    err = eng.vf_fit(g["fit_perms"][:2], 64, 1e-3, 1e-3, return_errors=True)
    np.testing.assert_allclose(err, g["fit1_err"], rtol=2e-4)
    w, mm, vv, step = eng.vf_get_state()
    # SURVEY 8(d) gate: weights rel <= 1e-4 after the first call with the same permutation.  Measured on B200
    # (tools/fit_parity_report.py): <= 1.5e-5 for every kernel on every fixture of this test (summation order of
    # 155-220 chaotic Adam steps decides the 5th digit)
    assert rel(w, g["fit1_w"]) < 1e-4
    eng.vf_fit(g["fit_perms"][2:4], 64, 1e-3, 1e-3)
    w, mm, vv, step = eng.vf_get_state()
    assert step == int(g["fit2_step"])
    # Second call (Adam state carried over, 4 epochs in total): the chain is chaotic, so this gate is tied to what the
    # CPU oracle itself reaches against the reference's fixture (tests/test_oracle.py; dead-unit drift under Adam).  On
    # cheetah_24x500 (560 steps) the fp32 ORACLE is already 2.0e-2 (weights) / 2.2e-3 (second moments) away from the
    # reference -- the gate is 3x that; on the shorter fixtures the oracle sits at <= 2e-3 / 1e-6.
    long_chain = case == "cheetah_24x500"
    assert rel(w, g["fit2_w"]) < (6e-2 if long_chain else 2e-2)
    assert rel(vv, g["fit2_v"]) < (7e-3 if long_chain else 1e-3)
    eng.vf_predict()
    # (cheetah_24x500: the fp32 oracle's own predictions are 6.5e-3 away from the reference's after the second call)
    np.testing.assert_allclose(eng.baseline(), g["fit2_predict"], rtol=0, atol=2e-2 if long_chain else 2e-4)
    eng.close()


@pytest.mark.parametrize("obs_dim", [39, 92, 100, 376])
def test_baseline_fit_wide_inputs(obs_dim, cuda_device):
    """More than 32 input features: the tensor-core kernel splits layer 1 over helper CTAs of one cluster (K = obs_dim + 4
    = 43 -> 1 helper with a partial slice; 96 -> 1 full; 104 -> 2; 380 = cfg5's humanoid -> 6).  Checked against the
    oracle's chain and against the fp32-FMA kernel on the same permutations, Adam state carried over a second call.

    The yardstick is the oracle in fp32 AND in fp64 (inputs rounded to fp32 first): on the 380-wide case the fp32 torch
    chain itself takes one ReLU flip the exact chain does not (fp64 oracle, tcgen05 kernel and FMA kernel all sit at the
    same 7.59e-3 from it and within 1e-6 of each other -- tools/fit_wide_report.py), so each gate takes the nearer of
    the two oracle chains.  Measured: <= 8e-7 (weights) on every shape."""
    import torch
    from mjrl_b200.engine import Engine
    paths = O.synthetic_paths(obs_dim, 3, 16, 200, seed=obs_dim)
    gamma = 0.995
    O.compute_returns(paths, gamma)
    n = sum(len(p["rewards"]) for p in paths)
    perms = [np.random.RandomState(5 + i).permutation(n).astype(np.int32) for i in range(3)]
    ora = []
    for dt in (torch.float32, torch.float64):
        vf = O.VFState(obs_dim, (128, 128), seed=4)
        w0 = vf.w.copy()
        e1 = O.vf_fit(vf, paths, perms[:2], 2, 64, 1e-3, 1e-3, return_errors=True, dtype=dt)
        w1 = vf.w.copy()
        O.vf_fit(vf, paths, perms[2:], 1, 64, 1e-3, 1e-3, dtype=dt)
        ora.append((e1, w1, vf.w.copy(), vf.v.copy(), vf.t))
    near = lambda x, k: min(rel(x, o[k]) for o in ora)
    res = {}
    for tc in (True, False):
        eng = Engine(obs_dim, 3, (64, 64), max_samples=n + 8, max_paths=32)
        eng.vf_set_state(w0)
        eng.vf_set_tensor_cores(tc)
        eng.upload_paths(paths)
        eng.compute_returns(gamma)
        err = eng.vf_fit(perms[:2], 64, 1e-3, 1e-3, return_errors=True)
        np.testing.assert_allclose(err, ora[0][0], rtol=2e-4)
        assert near(eng.vf_get_state()[0], 1) < 1e-5
        eng.vf_fit(perms[2:], 64, 1e-3, 1e-3)
        w, mm, vv, step = eng.vf_get_state()
        assert step == ora[0][4]
        assert near(w, 2) < 1e-5 and near(vv, 3) < 1e-4, (near(w, 2), near(vv, 3))
        eng.vf_predict()
        res[tc] = (w, eng.baseline())
        eng.close()
    assert rel(res[True][0], res[False][0]) < 1e-5
    np.testing.assert_allclose(res[True][1], res[False][1], rtol=0, atol=1e-4)


def test_fit_too_small_raises(cuda_device):
    from mjrl_b200.engine import Engine, MjbError
    paths = O.synthetic_paths(3, 1, 2, 50, seed=0)
    eng = Engine(3, 1, (32, 32), max_samples=256, max_paths=8)
    eng.upload_paths(paths)
    eng.compute_returns(0.9)
    with pytest.raises(MjbError):
        eng.vf_fit(np.arange(100, dtype=np.int32))
    eng.close()


def test_empty_and_ragged_edges(cuda_device):
    """Length-1 paths, a single path, unequal lengths crossing tile boundaries."""
    from mjrl_b200.engine import Engine
    rng = np.random.RandomState(3)
    lens = [1, 257, 1, 130, 2, 511]
    paths = [dict(observations=rng.randn(T, 5), actions=rng.randn(T, 3), rewards=rng.randn(T),
                  terminated=bool(i % 2)) for i, T in enumerate(lens)]
    spec = O.PolicySpec(5, 3, (64, 64))
    th = O.init_policy_params(spec, 1)
    eng = Engine(5, 3, (64, 64), max_samples=1024, max_paths=8)
    eng.set_params(th)
    vf = O.VFState(5, (128, 128), seed=2)
    eng.vf_set_state(vf.w)
    eng.upload_paths(paths)
    eng.compute_returns(0.99)
    assert np.array_equal(eng.returns(), np.concatenate([O.discount_sum(p["rewards"], 0.99) for p in paths]))
    eng.vf_predict()
    base = eng.baseline()
    np.testing.assert_allclose(base, np.concatenate([O.vf_predict(vf, p) for p in paths]), atol=5e-6, rtol=0)
    eng.compute_advantages(0.99, 0.95)
    adv = eng.advantages()
    k = 0
    for p in paths:
        T = len(p["rewards"])
        assert np.array_equal(adv[k:k + T], O.gae_path(p["rewards"], base[k:k + T], p["terminated"], 0.99, 0.95))
        k += T
    eng.process_paths()
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    white = eng.adv_white()
    assert rel(eng.vpg(), O.flat_vpg(spec, th, obs, act, white)) < 1e-5
    v = rng.randn(spec.d).astype(np.float32)
    assert rel(eng.fvp(v, 1e-4), O.fvp(spec, th, obs, v, 1e-4)) < 1e-5
    eng.close()


@pytest.mark.parametrize("shape", [(39, 28, (256, 256), 3000), (70, 4, (64, 32), 2000), (376, 17, (), 3000),
                                   (3, 1, (32, 32), 1500), (45, 9, (100, 60), 1800)])
def test_shapes_vs_oracle(shape, cuda_device):
    """Shapes outside the goldens (cfg4 door 39/28 256x256, cfg5 humanoid-linear 376/17, unequal / unpadded
    hidden widths, obs wider than one 32-feature chunk) against the fp64 closed-form oracle."""
    from mjrl_b200.engine import Engine
    obs_dim, act_dim, hidden, n = shape
    rng = np.random.RandomState(7)
    paths = O.synthetic_paths(obs_dim, act_dim, 6, n // 6, seed=11, ragged=True)
    spec = O.PolicySpec(obs_dim, act_dim, hidden)
    th = O.init_policy_params(spec, 3)
    th[-act_dim:] = 0.1 * rng.randn(act_dim)                      # non-trivial log_std
    eng = Engine(obs_dim, act_dim, hidden, max_samples=n + 8, max_paths=8)
    eng.set_params(th)
    eng.upload_paths(paths)
    N = eng.n
    adv = rng.randn(N)
    eng.set_advantages(adv)
    eng.process_paths()
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    white = O.whiten(adv)
    assert rel(eng.vpg(), O.flat_vpg(spec, th, obs, act, white)) < 1e-5
    v = rng.randn(spec.d).astype(np.float32)
    assert rel(eng.fvp(v, 1e-4), O.fvp(spec, th, obs, v, 1e-4)) < 1e-5
    idx = rng.randint(0, N, size=N // 2).astype(np.int32)
    assert rel(eng.fvp(v, 1e-4, idx=idx), O.fvp(spec, th, obs[idx], v, 1e-4)) < 1e-5
    th2 = spec.clamp(th + 0.01 * rng.randn(spec.d).astype(np.float32))
    eng.set_params(th2, set_new=True, set_old=False)
    surr, kl = eng.eval()
    assert abs(surr - float(O.surrogate(spec, th2, th, obs, act, white))) < 1e-4
    assert abs(kl / float(O.mean_kl(spec, th2, th, obs)) - 1) < 1e-3
    assert rel(eng.vpg(), O.flat_vpg(spec, th2, obs, act, white, theta_old=th)) < 2e-5
    eng.close()


def test_tensor_core_fvp_matches_fma_and_oracle(cuda_device):
    """The tcgen05 FVP (two-term fp16 split, fp32 accumulation in TMEM) against the fp32-FMA kernel, the fp64 oracle
    and the reference fixture, including a ragged tail tile, a subsample gather and a badly scaled tangent."""
    g = load_golden("cheetah_24x500")
    paths = golden_paths(g)
    m = g["meta"]
    eng = make_engine(g, cuda_device)
    eng.upload_paths(paths)
    obs = np.concatenate([p["observations"] for p in paths])
    spec = O.PolicySpec(m["obs_dim"], m["act_dim"], m["hidden"])
    assert eng.set_tensor_cores(True) is True
    rng = np.random.RandomState(5)
    for scale in (1.0, 1e-6, 3e4):
        v = (scale * rng.randn(spec.d)).astype(np.float32)
        want = O.fvp(spec, g["theta0"], obs, v, m["damping"])
        eng.set_tensor_cores(True)
        tc = eng.fvp(v, m["damping"])
        eng.set_tensor_cores(False)
        fma = eng.fvp(v, m["damping"])
        assert rel(fma, want) < 1e-5
        assert rel(tc, want) < 1e-5, rel(tc, want)
    eng.set_tensor_cores(True)
    assert rel(eng.fvp(g["fvp_vec"], m["damping"]), g["fvp_out"]) < 1e-5
    idx = rng.randint(0, eng.n, size=eng.n // 3).astype(np.int32)
    v = rng.randn(spec.d).astype(np.float32)
    assert rel(eng.fvp(v, m["damping"], idx=idx), O.fvp(spec, g["theta0"], obs[idx], v, m["damping"])) < 1e-5
    # non-default log_std / transforms flow through the tensor-core path too
    th = g["theta0"].copy()
    th[-m["act_dim"]:] = 0.3 * rng.randn(m["act_dim"])
    eng.set_params(th)
    assert rel(eng.fvp(v, 1e-4), O.fvp(spec, th, obs, v, 1e-4)) < 1e-5
    # the linear policy has its own tensor-core FVP kernel (HBM-bound path): golden + wide-observation shape
    for case_obs in (None, 376, 93):
        if case_obs is None:
            gl = load_golden("linear_30x200")
            pl = golden_paths(gl)
            el = make_engine(gl, cuda_device)
            th_l = gl["theta0"]
            specl = O.PolicySpec(gl["meta"]["obs_dim"], gl["meta"]["act_dim"], ())
        else:
            pl = O.synthetic_paths(case_obs, 17, 9, 300, seed=3, ragged=True)
            specl = O.PolicySpec(case_obs, 17, ())
            th_l = O.init_policy_params(specl, 2)
            th_l[-17:] = 0.2 * rng.randn(17)
            from mjrl_b200.engine import Engine
            el = Engine(case_obs, 17, (), max_samples=4000, max_paths=16)
            el.set_params(th_l)
        el.upload_paths(pl)
        obs_l = np.concatenate([p["observations"] for p in pl])
        vl = rng.randn(specl.d).astype(np.float32)
        want = O.fvp(specl, th_l, obs_l, vl, 1e-4)
        assert el.set_tensor_cores(True) is True
        assert rel(el.fvp(vl, 1e-4), want) < 1e-5, (case_obs, rel(el.fvp(vl, 1e-4), want))
        el.set_tensor_cores(False)
        assert rel(el.fvp(vl, 1e-4), want) < 1e-5
        el.set_tensor_cores(True)
        idl = rng.randint(0, el.n, size=el.n // 2).astype(np.int32)
        assert rel(el.fvp(vl, 1e-4, idx=idl), O.fvp(specl, th_l, obs_l[idl], vl, 1e-4)) < 1e-5
        if case_obs is not None:     # non-identity observation / action transforms (the kernel's general path)
            tr = dict(in_shift=0.3 * rng.randn(case_obs), in_scale=0.5 + rng.rand(case_obs),
                      out_shift=0.1 * rng.randn(17), out_scale=0.5 + rng.rand(17))
            spect = O.PolicySpec(case_obs, 17, (), **tr)
            el.set_transforms(**{k: v.astype(np.float32) for k, v in tr.items()})
            want_t = O.fvp(spect, th_l, obs_l, vl, 1e-4)
            assert rel(el.fvp(vl, 1e-4), want_t) < 1e-5
            el.set_tensor_cores(False)
            assert rel(el.fvp(vl, 1e-4), want_t) < 1e-5
        el.close()
    # shapes without a tensor-core kernel report it and keep working on the FMA kernels
    g2 = load_golden("swim_40x250")
    e2 = make_engine(g2, cuda_device)
    assert e2.set_tensor_cores(True) is False
    e2.close()
    eng.close()

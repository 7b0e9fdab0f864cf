"""GPU tests at BASELINE.json's FULL sizes (cfg3: 1e6 timesteps, 128x128; cfg5: 5e5 timesteps, 376-dim linear) through
size-independent properties -- the oracle is too slow for a dense comparison at these sizes, so the checks are
identities the reference's algorithm satisfies at any size:

  returns ........... R_t = r_t + gamma R_{t+1} per path, bit-exact (fp64 recursion, process_samples.py:6-14)
  whitening ......... mean 0 / std 1 of the whitened advantages (batch_reinforce.py:186-188)
  FVP ............... linear, symmetric (v.Fw = w.Fv), positive (v.Fv > damping |v|^2) (npg_cg.py:62-81)
  FVP kernels ....... tensor-core and fp32-FMA kernels agree to 1e-5; subsampled == dense on the gathered rows
  CG ................ the step satisfies alpha = sqrt(|delta / g.x|) and KL ~ delta/2 after the NPG step (npg_cg.py:108-127)
  fit ............... N/64 - 1 optimizer steps, error decreases, bit-identical when repeated (fixed-order sums)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GAMMA, LAM = 0.995, 0.97


def synthetic_batch(obs_dim, act_dim, n_paths, horizon, seed):
    rng = np.random.RandomState(seed)
    n = n_paths * horizon
    obs = rng.randn(n, obs_dim).astype(np.float32)
    act = rng.randn(n, act_dim).astype(np.float32)
    rew = rng.randn(n)
    lens = np.full(n_paths, horizon, np.int32)
    return obs, act, rew, lens


def make(obs_dim, act_dim, hidden, n_paths, horizon, seed=0):
    from mjrl_b200.engine import Engine
    from oracle import npg_oracle as O
    obs, act, rew, lens = synthetic_batch(obs_dim, act_dim, n_paths, horizon, seed)
    eng = Engine(obs_dim, act_dim, hidden, max_samples=n_paths * horizon + 8, max_paths=n_paths + 1)
    spec = O.PolicySpec(obs_dim, act_dim, hidden)
    th = O.init_policy_params(spec, 1)
    th[-act_dim:] = -0.5
    eng.set_params(th)
    vf = O.VFState(obs_dim, (128, 128), seed=2)
    eng.vf_set_state(vf.w)
    eng.upload_flat(obs, act, rew, lens, np.zeros(n_paths, np.uint8))
    return eng, rew, spec


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / (np.linalg.norm(np.asarray(b, np.float64)) + 1e-30))


@pytest.mark.parametrize("shape", [(17, 6, (128, 128), 1000, 1000), (376, 17, (), 500, 1000)], ids=["cfg3", "cfg5"])
def test_fullsize_properties(shape, cuda_device):
    obs_dim, act_dim, hidden, n_paths, horizon = shape
    eng, rew, spec = make(*shape)
    n = n_paths * horizon
    # ---- returns: exact fp64 recursion on every path ----
    eng.compute_returns(GAMMA)
    ret = eng.returns().reshape(n_paths, horizon)
    r = rew.reshape(n_paths, horizon)
    assert np.array_equal(ret[:, -1], r[:, -1])
    assert np.array_equal(ret[:, :-1], r[:, :-1] + GAMMA * ret[:, 1:])
    # ---- advantages / whitening ----
    eng.vf_predict()
    eng.compute_advantages(GAMMA, LAM)
    eng.process_paths()
    w = eng.adv_white()
    assert abs(float(w.mean())) < 1e-5 and abs(float(w.std()) - 1.0) < 1e-4
    # ---- FVP: linearity, symmetry, positivity ----
    rng = np.random.RandomState(7)
    d = eng.d
    v, u = rng.randn(d).astype(np.float32), rng.randn(d).astype(np.float32)
    damp = 1e-4
    Fv, Fu = eng.fvp(v, damp), eng.fvp(u, damp)
    Fc = eng.fvp((0.5 * v - 2.0 * u).astype(np.float32), damp)
    assert rel(Fc, 0.5 * Fv.astype(np.float64) - 2.0 * Fu.astype(np.float64)) < 2e-5
    vu, uv = float(np.dot(v.astype(np.float64), Fu)), float(np.dot(u.astype(np.float64), Fv))
    assert abs(vu - uv) <= 1e-4 * max(abs(vu), abs(uv), float(np.linalg.norm(Fu)) * 1e-2)
    assert float(np.dot(v.astype(np.float64), Fv)) > damp * float(np.dot(v, v)) * 0.999
    # ---- the two FVP kernels agree; the subsampled FVP equals the dense one on the gathered rows ----
    if eng.set_tensor_cores(True):
        eng.set_tensor_cores(False)
        assert rel(eng.fvp(v, damp), Fv) < 1e-5
        eng.set_tensor_cores(True)
    idx = rng.randint(0, n, size=n // 10).astype(np.int32)
    Fs = eng.fvp(v, damp, idx=idx)
    assert np.all(np.isfinite(Fs)) and rel(Fs, Fv) < 0.2            # same operator up to sampling noise
    # ---- NPG step: step-size rule and resulting KL ----
    delta = 0.05
    st = eng.step("npg", step_size=delta, cg_iters=10, damping=damp)
    g, x = eng.last_vectors()
    gx = float(np.dot(g.astype(np.float64), x.astype(np.float64)))
    assert gx > 0
    assert abs(st.alpha - np.sqrt(abs(delta / gx))) <= 1e-4 * st.alpha
    assert abs(st.vpg_dot_npg - gx) <= 1e-4 * abs(gx)
    assert 0.1 * delta < st.kl_dist < 1.0 * delta                    # KL ~ alpha^2 x.Fx / 2 ~ delta / 2 (CG not fully converged)
    assert st.surr_after > st.surr_before
    eng.close()


def test_fullsize_fit_properties(cuda_device):
    """cfg3's baseline fit: 15 624 Adam steps per epoch on the default (tensor-core) kernel."""
    obs_dim, act_dim, hidden, n_paths, horizon = 17, 6, (128, 128), 1000, 1000
    eng, rew, spec = make(obs_dim, act_dim, hidden, n_paths, horizon)
    n = n_paths * horizon
    eng.compute_returns(GAMMA)
    w0, m0, v0, s0 = eng.vf_get_state()
    perm = np.random.RandomState(3).permutation(n).astype(np.int32)
    e0, e1 = eng.vf_fit(perm, 64, 1e-3, 1e-3, return_errors=True)
    w1, m1, v1, s1 = eng.vf_get_state()
    assert s1 - s0 == n // 64 - 1
    assert np.all(np.isfinite(w1)) and e1 < e0
    # bit-identical when repeated from the same state (fixed-order sums, no atomics)
    eng.vf_set_state(w0, m0, v0, s0)
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    w2 = eng.vf_get_state()[0]
    assert np.array_equal(w1, w2)
    eng.close()

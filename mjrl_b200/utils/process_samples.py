"""compute_returns / compute_advantages with the reference's signatures (mjrl/utils/process_samples.py:3-35),
evaluated by the CUDA engine: one upload of the trajectories, fp64 reverse scans in the reference's exact
operation order (bit-identical returns; bit-identical GAE given the same baseline predictions), then the
per-path arrays are written back into the path dicts as the reference does."""
import numpy as np

from mjrl_b200 import runtime


def _engine_for(paths, baseline=None):
    p0 = paths[0]
    obs_dim = int(np.prod(np.asarray(p0["observations"]).shape[1:]))
    act_dim = int(np.prod(np.asarray(p0["actions"]).shape[1:]))
    n = int(sum(len(p["rewards"]) for p in paths))
    if baseline is not None and hasattr(baseline, "_eng"):
        return baseline._eng(n, len(paths))
    return runtime.get_engine(obs_dim, act_dim, None, need_samples=n, need_paths=len(paths))


def _scatter(paths, key, flat):
    k = 0
    for p in paths:
        T = len(p["rewards"])
        p[key] = flat[k:k + T]
        k += T


def compute_returns(paths, gamma):
    if len(paths) == 0:
        return
    returns_on(_engine_for(paths), paths, gamma)


def returns_on(eng, paths, gamma, write_back=True):
    runtime.ensure_resident(eng, paths)
    eng.compute_returns(gamma)
    if write_back:
        _scatter(paths, "returns", eng.returns())


def returns_write_back(eng, paths):
    _scatter(paths, "returns", eng.returns())


def compute_advantages(paths, baseline, gamma, gae_lambda=None, normalize=False):
    if len(paths) == 0:
        return
    advantages_on(_engine_for(paths, baseline), paths, baseline, gamma, gae_lambda, normalize)


def advantages_on(eng, paths, baseline, gamma, gae_lambda=None, normalize=False, fit_in_flight=False):
    """fit_in_flight: this step's baseline fit was already launched (it only needs the returns); the advantages use the
    pre-fit baseline exactly as the reference's program order does (batch_reinforce.py:98 before :108)."""
    runtime.ensure_resident(eng, paths)
    if not eng.have_returns:
        if "returns" in paths[0]:
            eng.set_returns(np.concatenate([p["returns"] for p in paths]))
        else:
            eng.compute_returns(gamma)
    if hasattr(baseline, "_eng"):
        if not fit_in_flight:
            baseline._bind(eng)
            baseline._eng()                            # push the host weights if they changed
        eng.vf_predict(prefit=fit_in_flight)           # all paths in one launch (mlp_baseline.py:97-105)
        base = eng.baseline()
    elif hasattr(baseline, "predict_resident"):        # ridge baselines: one launch over the resident batch
        baseline.predict_resident(eng)
        base = eng.baseline()
    else:                                              # any other baseline object keeps working on the host
        base = np.concatenate([np.asarray(baseline.predict(p), dtype=np.float32).ravel() for p in paths])
        eng.set_baseline(base)
    eng.compute_advantages(gamma, gae_lambda)
    adv = eng.advantages()
    if normalize:                                      # process_samples.py:14-19,30-35 (unused by the agents)
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        eng.set_advantages(adv)
        eng.adv_on_device = True       # what the device holds IS what the path dicts receive below
    _scatter(paths, "baseline", base)
    _scatter(paths, "advantages", adv)


def discount_sum(x, gamma, terminal=0.0):
    """Host helper kept for API compatibility (process_samples.py:37-44)."""
    x = np.asarray(x)
    y = np.empty_like(x, dtype=np.result_type(x.dtype, np.float32))
    run = terminal
    for t in range(len(x) - 1, -1, -1):
        run = x[t] + gamma * run
        y[t] = run
    return y

"""LinearBaseline with the N-dependent work on the device-resident batch (reference: baselines/linear_baseline.py).

Same constructor, `fit(paths, return_errors)`, `predict(path)` and `_coeffs` as the reference.  When the agent has the
batch resident on the engine (`fit_resident` / `predict_resident`, used by `BatchREINFORCE` and `process_samples`), the
feature matrix is never built: one CUDA pass accumulates F^T F, F^T y and y^T y in float64 (`mjb_ridge_gram`), the K x K
system is solved on the host exactly as the reference does (`np.linalg.lstsq`, regulariser x10 while the solution has
NaNs), and the predictions of all paths are written into the device baseline buffer by one launch (`mjb_ridge_predict`).
Called with path dicts that are not resident (e.g. evaluation code), `fit` / `predict` upload them first.
"""
import copy

import numpy as np

from .. import runtime

KIND = 0


class LinearBaseline:
    _kind = KIND

    def __init__(self, env_spec, inp_dim=None, inp='obs', reg_coeff=1e-5):
        if inp != 'obs':
            raise NotImplementedError("mjrl_b200 ridge baselines take inp='obs' (the observations resident on the device)")
        self.n = inp_dim if inp_dim is not None else env_spec.observation_dim
        self.inp = inp
        self._reg_coeff = reg_coeff
        self._coeffs = None
        self.act_dim = getattr(env_spec, "action_dim", 1)

    # ---- host side of the solve: linear_baseline.py:48-56 / quadratic_baseline.py:57-65
    def _solve(self, gram, rhs):
        reg_coeff = copy.deepcopy(self._reg_coeff)
        for _ in range(10):
            self._coeffs = np.linalg.lstsq(gram + reg_coeff * np.identity(gram.shape[0]), rhs, rcond=-1)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg_coeff *= 10

    # ---- resident entry points (the agent's engine already holds the batch and its returns)
    def fit_resident(self, eng, return_errors=False):
        gram, rhs, yy = eng.ridge_gram(self._kind)
        if return_errors:
            if self._coeffs is not None:
                c = self._coeffs
                error_before = (yy - 2.0 * rhs.dot(c) + c.dot(gram).dot(c)) / yy
            else:
                error_before = 1.0
        self._solve(gram, rhs)
        if return_errors:
            c = self._coeffs
            error_after = (yy - 2.0 * rhs.dot(c) + c.dot(gram).dot(c)) / yy
            return error_before, error_after

    def predict_resident(self, eng):
        """Predictions of every resident path into the device baseline buffer (zeros before the first fit)."""
        if self._coeffs is None:
            eng.set_baseline(np.zeros(eng.n, dtype=np.float32))
        else:
            eng.ridge_predict(self._kind, self._coeffs)

    # ---- reference API on path dicts
    def _engine_for(self, paths):
        n = int(sum(len(p["rewards"]) for p in paths))
        eng = runtime.get_engine(self.n, self.act_dim, None, need_samples=n, need_paths=len(paths))
        runtime.ensure_resident(eng, paths)
        return eng

    def fit(self, paths, return_errors=False):
        eng = self._engine_for(paths)
        eng.set_returns(np.concatenate([p["returns"] for p in paths]))     # the dicts' returns are authoritative here
        return self.fit_resident(eng, return_errors=return_errors)

    def predict(self, path):
        if self._coeffs is None:
            return np.zeros(len(path["rewards"]))
        eng = self._engine_for([path])
        self.predict_resident(eng)
        return eng.baseline().astype(np.float64)

"""MLP value-function baseline with the reference's interface (mjrl/baselines/mlp_baseline.py:11-105).

`fit` / `predict` run on the GPU engine; the object itself stays a picklable CPU container (nn.Sequential
weights + Adam moments + step count) that is refreshed from the device after every fit, so checkpoints and
the reference's `train_agent` keep working."""
import numpy as np
import torch
import torch.nn as nn

from mjrl_b200 import runtime


class MLPBaseline:
    def __init__(self, env_spec, inp_dim=None, inp='obs', learn_rate=1e-3, reg_coef=0.0,
                 batch_size=64, epochs=1, use_gpu=False, hidden_sizes=(128, 128)):
        self.n = inp_dim if inp_dim is not None else env_spec.observation_dim
        self.act_dim = getattr(env_spec, "action_dim", 1)
        self.batch_size, self.epochs, self.reg_coef, self.learn_rate = batch_size, epochs, reg_coef, learn_rate
        self.use_gpu = use_gpu                    # accepted for signature compatibility; the engine is always CUDA
        self.inp = inp
        if inp != 'obs':
            raise NotImplementedError("inp='env_features' is not supported by the CUDA baseline")
        self.hidden_sizes = tuple(hidden_sizes)
        self.model = nn.Sequential()
        sizes = (self.n + 4,) + self.hidden_sizes + (1,)
        for i in range(len(sizes) - 1):
            self.model.add_module('fc_' + str(i), nn.Linear(sizes[i], sizes[i + 1]))
            if i != len(sizes) - 2:
                self.model.add_module('relu_' + str(i), nn.ReLU())
        d = sum(p.numel() for p in self.model.parameters())
        # Adam state of torch.optim.Adam(lr, weight_decay=reg_coef), kept flat (mlp_baseline.py:33)
        self.adam_m = np.zeros(d, np.float32)
        self.adam_v = np.zeros(d, np.float32)
        self.adam_step = 0
        self._engine = None
        self._device_current = False

    # ---- state <-> engine ----
    def get_flat_weights(self):
        return np.concatenate([p.data.numpy().ravel() for p in self.model.parameters()]).astype(np.float32)

    def set_flat_weights(self, w):
        k = 0
        for p in self.model.parameters():
            n = p.numel()
            p.data = torch.from_numpy(np.ascontiguousarray(w[k:k + n]).reshape(tuple(p.shape))).float()
            k += n
        self._device_current = False

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        st["_device_current"] = False
        return st

    def _bind(self, engine):
        if engine is not self._engine:
            self._engine = engine
            self._device_current = False

    def _eng(self, need_samples=0, need_paths=0):
        if self._engine is None or need_samples > self._engine.max_samples or need_paths > self._engine.max_paths:
            self._bind(runtime.get_engine(self.n, self.act_dim, None, self.hidden_sizes,
                                          need_samples=need_samples, need_paths=need_paths))
        if not self._device_current:
            self._engine.vf_set_state(self.get_flat_weights(), self.adam_m, self.adam_v, self.adam_step)
            self._device_current = True
        return self._engine

    def _pull(self):
        w, m, v, step = self._engine.vf_get_state()
        self.set_flat_weights(w)
        self.adam_m, self.adam_v, self.adam_step = m, v, step
        self._device_current = True

    # ---- reference API ----
    def _features(self, paths):
        """Host restatement of the feature map for API compatibility (mlp_baseline.py:36-58); the engine builds
        the same features on the fly inside its kernels."""
        o = np.concatenate([path["observations"] for path in paths])
        o = np.clip(o, -10, 10) / 10.0
        if o.ndim > 2:
            o = o.reshape(o.shape[0], -1)
        feat = np.ones((o.shape[0], o.shape[1] + 4))
        feat[:, :o.shape[1]] = o
        k = 0
        for p in paths:
            l = len(p["rewards"])
            tau = np.arange(l) / 1000.0
            for j in range(4):
                feat[k:k + l, -4 + j] = tau ** (j + 1)
            k += l
        return feat

    def fit(self, paths, return_errors=False):
        """mlp_baseline.py:61-95."""
        e0 = self.fit_begin(paths, return_errors)
        e1 = self.fit_end(return_errors)
        if return_errors:
            return e0, e1

    def fit_begin(self, paths, return_errors=False):
        """Launch the fit on the engine's side stream and return immediately (error_before if requested).  The chain
        only needs the returns, so the agents start it before the policy update and join with fit_end() after."""
        n = int(sum(len(p["rewards"]) for p in paths))
        eng = self._eng(n, len(paths))
        runtime.ensure_resident(eng, paths)
        if not eng.have_returns:
            # returns were computed by someone else: the reference reads path["returns"] (mlp_baseline.py:64)
            eng.set_returns(np.concatenate([p["returns"] for p in paths]))
        n_glob = eng.n_global()
        # host RNG draw at the reference's program point (optimize_model.py:22): one permutation per epoch, from
        # numpy's global RandomState (bit-identical order and RNG state, see runtime.global_permutation)
        perms = np.stack([runtime.global_permutation(n_glob) for _ in range(self.epochs)])
        return eng.vf_fit_begin(perms, self.batch_size, self.learn_rate, self.reg_coef, return_errors=return_errors)

    def fit_end(self, return_errors=False):
        err = self._engine.vf_fit_end(return_errors)
        self._pull()
        return err

    def predict(self, path):
        eng = self._eng(len(path["rewards"]), 1)
        runtime.ensure_resident(eng, [path], force=True)     # replaces (and un-pins) whatever batch was resident
        eng.vf_predict()
        return eng.baseline()

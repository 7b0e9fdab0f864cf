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
        self._engine = None
        self._device_current = False
        self._fit_pending = False          # a fit is (or may still be) in flight on the device: host copies are stale
        model = nn.Sequential()
        sizes = (self.n + 4,) + self.hidden_sizes + (1,)
        for i in range(len(sizes) - 1):
            model.add_module('fc_' + str(i), nn.Linear(sizes[i], sizes[i + 1]))
            if i != len(sizes) - 2:
                model.add_module('relu_' + str(i), nn.ReLU())
        self._model = model
        d = sum(p.numel() for p in model.parameters())
        # Adam state of torch.optim.Adam(lr, weight_decay=reg_coef), kept flat (mlp_baseline.py:33)
        self._adam = [np.zeros(d, np.float32), np.zeros(d, np.float32), 0]

    # The host copies (nn.Sequential weights, Adam moments, step count) are what checkpoints and rollout workers see.
    # After update_from_paths the fit may still be running on the device (the agents do not wait for it: the next
    # batch's upload overlaps its tail); every read of the host state joins it first, so nobody can observe pre-fit
    # weights -- in particular pickling (__getstate__) mid-flight yields the post-fit state.
    def _join(self):
        if self._fit_pending:
            self._fit_pending = False
            self._engine.vf_fit_end(False)
            self._pull()

    @property
    def model(self):
        self._join()
        return self._model

    @model.setter
    def model(self, m):
        self._join()
        self._model = m
        self._device_current = False

    @property
    def adam_m(self):
        self._join()
        return self._adam[0]

    @adam_m.setter
    def adam_m(self, v):
        self._adam[0] = v

    @property
    def adam_v(self):
        self._join()
        return self._adam[1]

    @adam_v.setter
    def adam_v(self, v):
        self._adam[1] = v

    @property
    def adam_step(self):
        self._join()
        return self._adam[2]

    @adam_step.setter
    def adam_step(self, v):
        self._adam[2] = v

    # ---- state <-> engine ----
    def get_flat_weights(self):
        return np.concatenate([p.data.numpy().ravel() for p in self.model.parameters()]).astype(np.float32)

    def set_flat_weights(self, w):
        self._join()
        k = 0
        for p in self._model.parameters():
            n = p.numel()
            p.data = torch.from_numpy(np.ascontiguousarray(w[k:k + n]).reshape(tuple(p.shape))).float()
            k += n
        self._device_current = False

    def __getstate__(self):
        self._join()                       # a fit still in flight finishes first: the pickle holds post-fit weights
        st = dict(self.__dict__)
        st["_engine"] = None
        st["_device_current"] = False
        st["_fit_pending"] = False
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        if "_model" not in st and "model" in st:            # pickles written before the lazy-join refactor
            self._model = st["model"]
            self._adam = [st.get("adam_m"), st.get("adam_v"), st.get("adam_step", 0)]
            self._fit_pending = False

    def _bind(self, engine):
        if engine is self._engine:
            return
        if self._fit_pending:
            if self._engine is not None and not self._engine.closed:
                self._join()                       # finish on the engine that runs the fit, then move
            else:
                # runtime.get_engine replaced the engine by a larger one and carried the (joined) device state over:
                # the new engine is the authority, the host copies are refreshed from it
                self._engine, self._fit_pending = engine, False
                self._pull()
                return
        self._engine = engine
        self._device_current = False

    def _eng(self, need_samples=0, need_paths=0):
        if self._engine is None or need_samples > self._engine.max_samples or need_paths > self._engine.max_paths:
            self._bind(runtime.get_engine(self.n, self.act_dim, None, self.hidden_sizes,
                                          need_samples=need_samples, need_paths=need_paths))
        eng = self._engine
        owner = getattr(eng, "vf_owner", None)
        if owner is not self:
            # another baseline object used this engine's value-net slot last: it takes its state home first (joining
            # a fit of its own that may still be running), then this object's weights go up
            if owner is not None:
                owner._release()
            eng.vf_owner = self
            self._device_current = False
        if not self._device_current:
            eng.vf_set_state(self.get_flat_weights(), self.adam_m, self.adam_v, self.adam_step)
            self._device_current = True
        return eng

    def _release(self):
        self._join()
        self._device_current = False

    def _pull(self):
        w, m, v, step = self._engine.vf_get_state()
        k = 0
        for p in self._model.parameters():
            n = p.numel()
            p.data = torch.from_numpy(np.ascontiguousarray(w[k:k + n]).reshape(tuple(p.shape))).float()
            k += n
        self._adam = [m, v, step]
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
        perms = runtime.global_permutations(n_glob, self.epochs)
        return eng.vf_fit_begin(perms, self.batch_size, self.learn_rate, self.reg_coef, return_errors=return_errors)

    def fit_begin_resident(self, eng, return_errors=False):
        """fit_begin for a batch that is already the engine's resident rollout batch with its returns computed on the
        device (BatchREINFORCE.update_from_rollouts): no path dicts, no upload."""
        self._bind(eng)
        self._eng()
        assert eng.have_returns, "compute the returns on the engine first"
        n_glob = eng.n_global()
        perms = runtime.global_permutations(n_glob, self.epochs)
        return eng.vf_fit_begin(perms, self.batch_size, self.learn_rate, self.reg_coef, return_errors=return_errors)

    def fit_end(self, return_errors=False):
        self._fit_pending = False
        err = self._engine.vf_fit_end(return_errors)
        self._pull()
        return err

    def fit_defer(self):
        """Leave the fit launched by fit_begin() running; whoever reads the baseline next joins it (see _join)."""
        self._fit_pending = True

    def predict(self, path):
        eng = self._eng(len(path["rewards"]), 1)
        runtime.ensure_resident(eng, [path], force=True)     # replaces (and un-pins) whatever batch was resident
        eng.vf_predict()
        return eng.baseline()

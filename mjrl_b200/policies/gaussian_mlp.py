"""Host-side Gaussian MLP policy object behind the reference's duck-typed interface
(mjrl/policies/gaussian_mlp.py:8-145).

The object stays a picklable CPU container (samplers pickle it into rollout workers, train_agent into checkpoints)
whose tensors are refreshed from the device after every update.  All batched math of the update path -- likelihoods,
KL, gradients, Fisher-vector products -- runs in the CUDA engine on the flat parameter vector
[W1 (h1 x obs), b1, W2, b2, W3, b3, log_std]; what lives here is the flat <-> tensor bookkeeping, single-observation
action sampling, and small-input helpers with the reference's signatures."""
import math

import numpy as np
import torch

from mjrl_b200.utils.fc_network import FCNetwork


class _FlatView:
    """A list of tensors seen as one flat float32 vector (network parameters followed by log_std)."""

    def __init__(self, tensors):
        self.tensors = tensors
        self.shapes = [tuple(t.shape) for t in tensors]
        self.sizes = [int(t.numel()) for t in tensors]
        self.bounds = np.concatenate([[0], np.cumsum(self.sizes)])

    def read(self):
        return np.concatenate([t.detach().reshape(-1).numpy() for t in self.tensors]).astype(np.float32, copy=True)

    def write(self, flat, log_std_floor):
        flat = np.asarray(flat)
        for t, shape, lo, hi in zip(self.tensors, self.shapes, self.bounds[:-1], self.bounds[1:]):
            t.data = torch.tensor(np.reshape(flat[lo:hi], shape), dtype=torch.float32)
        last = self.tensors[-1]                                     # log_std >= min_log_std (gaussian_mlp.py:73-75,85-87)
        last.data = last.data.clamp(min=log_std_floor)


class MLP:
    hidden_sizes_default = (64, 64)

    def __init__(self, env_spec, hidden_sizes=(64, 64), min_log_std=-3, init_log_std=0, seed=None):
        self.n, self.m = env_spec.observation_dim, env_spec.action_dim
        self.min_log_std = min_log_std
        self.hidden_sizes = tuple(hidden_sizes)
        if seed is not None:                                         # same seeding points as the reference (:26-28)
            torch.manual_seed(seed)
            np.random.seed(seed)

        def fresh_log_std():
            return torch.ones(self.m) * init_log_std

        # current policy: nn.Linear default init, output layer scaled down (:31-35)
        self.model = FCNetwork(self.n, self.m, self.hidden_sizes)
        out_w, out_b = list(self.model.parameters())[-2:]
        out_w.data.mul_(1e-2)
        out_b.data.mul_(1e-2)
        self.log_std = fresh_log_std().requires_grad_(True)
        self.trainable_params = [*self.model.parameters(), self.log_std]
        # "old" policy: an independent copy used by the surrogate / KL (:39-44)
        self.old_model = FCNetwork(self.n, self.m, self.hidden_sizes)
        self.old_log_std = fresh_log_std()
        self.old_params = [*self.old_model.parameters(), self.old_log_std]
        for dst, src in zip(self.old_params, self.trainable_params):
            dst.data = src.data.clone()
        self._new_view, self._old_view = _FlatView(self.trainable_params), _FlatView(self.old_params)
        self.param_shapes, self.param_sizes = list(self._new_view.shapes), list(self._new_view.sizes)
        self.d = int(np.sum(self.param_sizes))
        self.log_std_val = np.float64(self.log_std.detach().numpy().ravel())
        self.obs_var = torch.randn(self.n)                           # placeholder the reference keeps (:54)

    # pickles written before _FlatView existed (or by the reference layout) still load
    def __setstate__(self, state):
        self.__dict__.update(state)
        if "_new_view" not in state:
            self._new_view, self._old_view = _FlatView(self.trainable_params), _FlatView(self.old_params)

    # ---- flat parameter access (gaussian_mlp.py:60-87) ----
    def get_param_values(self):
        return self._new_view.read()

    def set_param_values(self, new_params, set_new=True, set_old=True):
        if set_new:
            self._new_view.write(new_params, self.min_log_std)
            self.log_std_val = np.float64(self.log_std.detach().numpy().ravel())
        if set_old:
            self._old_view.write(new_params, self.min_log_std)

    # ---- single-observation sampling for the rollout workers (gaussian_mlp.py:91-97) ----
    def get_action(self, observation):
        self.obs_var.data = torch.from_numpy(np.float32(np.reshape(observation, (1, -1))))
        with torch.no_grad():
            mu = self.model(self.obs_var).numpy().ravel()
        sampled = mu + np.exp(self.log_std_val) * np.random.randn(self.m)      # one global-RNG draw, as the reference
        return [sampled, dict(mean=mu, log_std=self.log_std_val, evaluation=mu)]

    # ---- small-input helpers with the reference's signatures (:99-145); the engine does the batched versions ----
    @staticmethod
    def _as_float_tensor(x):
        return x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x)).float()

    def mean_LL(self, observations, actions, model=None, log_std=None):
        net = model if model is not None else self.model
        ls = log_std if log_std is not None else self.log_std
        mu = net(self._as_float_tensor(observations))
        z = (self._as_float_tensor(actions) - mu) / ls.exp()
        ll = -0.5 * (z * z).sum(dim=1) - ls.sum() - 0.5 * self.m * math.log(2.0 * math.pi)
        return mu, ll

    def log_likelihood(self, observations, actions, model=None, log_std=None):
        return self.mean_LL(observations, actions, model, log_std)[1].detach().numpy()

    def _dist_info(self, observations, actions, net, ls):
        mu, ll = self.mean_LL(observations, actions, net, ls)
        return [ll, mu, ls]

    def old_dist_info(self, observations, actions):
        return self._dist_info(observations, actions, self.old_model, self.old_log_std)

    def new_dist_info(self, observations, actions):
        return self._dist_info(observations, actions, self.model, self.log_std)

    def likelihood_ratio(self, new_dist_info, old_dist_info):
        return (new_dist_info[0] - old_dist_info[0]).exp()

    def mean_kl(self, new_dist_info, old_dist_info):
        _, mu_new, ls_new = new_dist_info
        _, mu_old, ls_old = old_dist_info
        var_new, var_old = (2.0 * ls_new).exp(), (2.0 * ls_old).exp()
        per_dim = ((mu_old - mu_new) ** 2 + var_old - var_new) / (2.0 * var_new + 1e-8) + ls_new - ls_old
        return per_dim.sum(dim=1).mean()

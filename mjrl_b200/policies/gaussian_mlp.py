"""Gaussian MLP policy object with the reference's interface (mjrl/policies/gaussian_mlp.py:8-145).

It stays a picklable CPU object (the samplers pickle it into rollout workers, train_agent pickles it
into checkpoints) whose weights are current after every update; the batched math of the update path
(likelihoods, KL, gradients, Fisher products) runs in the CUDA engine on its flat parameter vector, whose
layout is the reference's: [W1 (h1 x obs), b1, W2, b2, W3, b3, log_std]."""
import numpy as np
import torch

from mjrl_b200.utils.fc_network import FCNetwork


class MLP:
    hidden_sizes_default = (64, 64)

    def __init__(self, env_spec, hidden_sizes=(64, 64), min_log_std=-3, init_log_std=0, seed=None):
        self.n = env_spec.observation_dim
        self.m = env_spec.action_dim
        self.min_log_std = min_log_std
        self.hidden_sizes = tuple(hidden_sizes)
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
        self.model = FCNetwork(self.n, self.m, self.hidden_sizes)
        for param in list(self.model.parameters())[-2:]:      # small last layer (gaussian_mlp.py:34-35)
            param.data = 1e-2 * param.data
        self.log_std = torch.ones(self.m) * init_log_std
        self.log_std.requires_grad_(True)
        self.trainable_params = list(self.model.parameters()) + [self.log_std]
        self.old_model = FCNetwork(self.n, self.m, self.hidden_sizes)
        self.old_log_std = torch.ones(self.m) * init_log_std
        self.old_params = list(self.old_model.parameters()) + [self.old_log_std]
        for idx, param in enumerate(self.old_params):
            param.data = self.trainable_params[idx].data.clone()
        self.log_std_val = np.float64(self.log_std.data.numpy().ravel())
        self.param_shapes = [p.data.numpy().shape for p in self.trainable_params]
        self.param_sizes = [p.data.numpy().size for p in self.trainable_params]
        self.d = np.sum(self.param_sizes)
        self.obs_var = torch.randn(self.n)

    # ---- flat parameter access (gaussian_mlp.py:60-87) ----
    def get_param_values(self):
        return np.concatenate([p.contiguous().view(-1).data.numpy() for p in self.trainable_params]).copy()

    def _assign(self, params, new_params, log_std_index=-1):
        k = 0
        for idx, param in enumerate(params):
            size = self.param_sizes[idx]
            vals = np.asarray(new_params[k:k + size]).reshape(self.param_shapes[idx])
            param.data = torch.from_numpy(np.ascontiguousarray(vals)).float()
            k += size
        params[log_std_index].data = torch.clamp(params[log_std_index], self.min_log_std).data

    def set_param_values(self, new_params, set_new=True, set_old=True):
        if set_new:
            self._assign(self.trainable_params, new_params)
            self.log_std_val = np.float64(self.log_std.data.numpy().ravel())
        if set_old:
            self._assign(self.old_params, new_params)

    # ---- sampling (gaussian_mlp.py:91-97) ----
    def get_action(self, observation):
        o = np.float32(observation.reshape(1, -1))
        self.obs_var.data = torch.from_numpy(o)
        mean = self.model(self.obs_var).data.numpy().ravel()
        noise = np.exp(self.log_std_val) * np.random.randn(self.m)
        return [mean + noise, {'mean': mean, 'log_std': self.log_std_val, 'evaluation': mean}]

    # ---- host-side batched helpers, same signatures as the reference (:99-145); small inputs only ----
    def mean_LL(self, observations, actions, model=None, log_std=None):
        model = self.model if model is None else model
        log_std = self.log_std if log_std is None else log_std
        obs = observations if torch.is_tensor(observations) else torch.from_numpy(observations).float()
        act = actions if torch.is_tensor(actions) else torch.from_numpy(actions).float()
        mean = model(obs)
        zs = (act - mean) / torch.exp(log_std)
        LL = -0.5 * torch.sum(zs ** 2, dim=1) - torch.sum(log_std) - 0.5 * self.m * np.log(2 * np.pi)
        return mean, LL

    def log_likelihood(self, observations, actions, model=None, log_std=None):
        return self.mean_LL(observations, actions, model, log_std)[1].data.numpy()

    def old_dist_info(self, observations, actions):
        mean, LL = self.mean_LL(observations, actions, self.old_model, self.old_log_std)
        return [LL, mean, self.old_log_std]

    def new_dist_info(self, observations, actions):
        mean, LL = self.mean_LL(observations, actions, self.model, self.log_std)
        return [LL, mean, self.log_std]

    def likelihood_ratio(self, new_dist_info, old_dist_info):
        return torch.exp(new_dist_info[0] - old_dist_info[0])

    def mean_kl(self, new_dist_info, old_dist_info):
        old_log_std, new_log_std = old_dist_info[2], new_dist_info[2]
        old_std, new_std = torch.exp(old_log_std), torch.exp(new_log_std)
        Nr = (old_dist_info[1] - new_dist_info[1]) ** 2 + old_std ** 2 - new_std ** 2
        Dr = 2 * new_std ** 2 + 1e-8
        return torch.mean(torch.sum(Nr / Dr + new_log_std - old_log_std, dim=1))

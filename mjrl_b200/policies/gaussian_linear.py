"""Gaussian linear policy (mjrl/policies/gaussian_linear.py:10-139): the MLP container with no hidden layer."""
from mjrl_b200.policies.gaussian_mlp import MLP


class LinearPolicy(MLP):
    def __init__(self, env_spec, min_log_std=-3, init_log_std=0, seed=None):
        super().__init__(env_spec, hidden_sizes=(), min_log_std=min_log_std, init_log_std=init_log_std, seed=seed)

"""Host-side (CPU) network container with the reference's FCNetwork interface (mjrl/utils/fc_network.py:6-52):
used for single-observation action sampling inside rollout workers and for pickling.  The batched update
path never runs this module -- it runs the CUDA kernels on the flat parameter vector."""
import numpy as np
import torch
import torch.nn as nn


class FCNetwork(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden_sizes=(64, 64), nonlinearity='tanh',
                 in_shift=None, in_scale=None, out_shift=None, out_scale=None):
        super().__init__()
        assert type(hidden_sizes) == tuple
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.layer_sizes = (obs_dim,) + hidden_sizes + (act_dim,)
        self.set_transformations(in_shift, in_scale, out_shift, out_scale)
        self.fc_layers = nn.ModuleList(
            [nn.Linear(a, b) for a, b in zip(self.layer_sizes[:-1], self.layer_sizes[1:])])
        self.nonlinearity = torch.relu if nonlinearity == 'relu' else torch.tanh

    def set_transformations(self, in_shift=None, in_scale=None, out_shift=None, out_scale=None):
        self.transformations = dict(in_shift=in_shift, in_scale=in_scale, out_shift=out_shift, out_scale=out_scale)
        as_t = lambda v, n, fill: torch.full((n,), fill) if v is None else torch.from_numpy(np.float32(v))
        self.in_shift, self.in_scale = as_t(in_shift, self.obs_dim, 0.0), as_t(in_scale, self.obs_dim, 1.0)
        self.out_shift, self.out_scale = as_t(out_shift, self.act_dim, 0.0), as_t(out_scale, self.act_dim, 1.0)

    def forward(self, x):
        h = (x.to('cpu') - self.in_shift) / (self.in_scale + 1e-8)
        for layer in self.fc_layers[:-1]:
            h = self.nonlinearity(layer(h))
        return self.fc_layers[-1](h) * self.out_scale + self.out_shift

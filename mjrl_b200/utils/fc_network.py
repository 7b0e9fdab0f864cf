"""CPU network container behind the reference's FCNetwork interface (mjrl/utils/fc_network.py:6-52).

Only single-observation action sampling inside rollout workers and pickling use this module; the batched update path
runs the CUDA kernels on the flat parameter vector instead.  Attribute names (`fc_layers`, `layer_sizes`,
`transformations`, `in_shift` ...) are the reference's because checkpoints and samplers reach for them."""
import numpy as np
import torch
from torch import nn

_ACTIVATIONS = {"tanh": torch.tanh, "relu": torch.relu}


def _affine_term(value, width, neutral):
    """None -> the neutral element (0 for shifts, 1 for scales), else the given vector as float32."""
    if value is None:
        return torch.full((width,), float(neutral))
    return torch.from_numpy(np.float32(value))


class FCNetwork(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden_sizes=(64, 64), nonlinearity='tanh',
                 in_shift=None, in_scale=None, out_shift=None, out_scale=None):
        super().__init__()
        if not isinstance(hidden_sizes, tuple):
            raise AssertionError("hidden_sizes must be a tuple (fc_network.py:17)")
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.layer_sizes = (obs_dim, *hidden_sizes, act_dim)
        self.set_transformations(in_shift, in_scale, out_shift, out_scale)
        widths = self.layer_sizes
        self.fc_layers = nn.ModuleList(nn.Linear(widths[i], widths[i + 1]) for i in range(len(widths) - 1))
        self.nonlinearity = _ACTIVATIONS.get(nonlinearity, torch.tanh)

    def set_transformations(self, in_shift=None, in_scale=None, out_shift=None, out_scale=None):
        self.transformations = {"in_shift": in_shift, "in_scale": in_scale, "out_shift": out_shift, "out_scale": out_scale}
        self.in_shift = _affine_term(in_shift, self.obs_dim, 0)
        self.in_scale = _affine_term(in_scale, self.obs_dim, 1)
        self.out_shift = _affine_term(out_shift, self.act_dim, 0)
        self.out_scale = _affine_term(out_scale, self.act_dim, 1)

    def forward(self, x):
        act = (x.to('cpu') - self.in_shift) / (self.in_scale + 1e-8)
        *hidden, head = self.fc_layers
        for lin in hidden:
            act = self.nonlinearity(lin(act))
        return head(act) * self.out_scale + self.out_shift

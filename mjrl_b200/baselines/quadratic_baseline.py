"""QuadraticBaseline on the device-resident batch (reference: baselines/quadratic_baseline.py): linear features plus all
products o_i o_j (i <= j), n + n(n+1)/2 + 5 columns -- see linear_baseline.py for how fit / predict are split between the
CUDA Gram / prediction kernels and the host solve."""
from .linear_baseline import LinearBaseline


class QuadraticBaseline(LinearBaseline):
    _kind = 1

    def __init__(self, env_spec, inp_dim=None, inp='obs', reg_coeff=1e-3):
        super().__init__(env_spec, inp_dim=inp_dim, inp=inp, reg_coeff=reg_coeff)

"""Developer tool: accuracy of the FVP kernels against the fp64 closed-form oracle as the batch grows (accumulation error)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402
from oracle import npg_oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


for (obs_dim, act_dim, hidden) in ((17, 6, (128, 128)), (376, 17, ())):
    for n in (20000, 200000, 1000000):
        if not hidden and n > 500000:
            continue
        rng = np.random.RandomState(0)
        obs = rng.randn(n, obs_dim).astype(np.float32)
        eng = Engine(obs_dim, act_dim, hidden, max_samples=n + 8, max_paths=8)
        spec = O.PolicySpec(obs_dim, act_dim, hidden)
        th = O.init_policy_params(spec, 1)
        th[-act_dim:] = -0.5
        eng.set_params(th)
        eng.upload_flat(obs, rng.randn(n, act_dim), rng.randn(n), np.array([n], np.int32), np.zeros(1, np.uint8))
        v = rng.randn(spec.d).astype(np.float32)
        t0 = time.time()
        want = O.fvp(spec, th, obs, v, 1e-4)
        t1 = time.time() - t0
        out = {}
        for tc in (True, False):
            eng.set_tensor_cores(tc)
            out[tc] = eng.fvp(v, 1e-4)
        d = spec.d - act_dim
        print("%s n=%7d  tensor-core rel %.2e (weights block %.2e)   fp32-FMA rel %.2e (weights block %.2e)   [oracle %.1fs]" % (
            "mlp128" if hidden else "linear", n, rel(out[True], want), rel(out[True][:d], want[:d]), rel(out[False], want),
            rel(out[False][:d], want[:d]), t1), flush=True)
        eng.close()

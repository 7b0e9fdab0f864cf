"""Developer tool (GPU box, under ncu): a short program that launches every hot kernel a few times at the full shapes:
cfg3 (17/6, 128x128, 1e6 timesteps): returns / GAE scans, VPG, 10-iteration CG (tcgen05 FVP), evaluation, baseline
predict + fit (2e5-sample prefix so that a replayed capture stays short); cfg5 (376/17 linear, 5e5): linear FVP."""
import sys

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
rng = np.random.RandomState(0)
if which == "cfg3":
    n_paths, T, od, ad = 1000, 1000, 17, 6
    n = n_paths * T
    eng = Engine(od, ad, (128, 128), max_samples=n + 8, max_paths=n_paths + 8)
    eng.upload_flat(rng.standard_normal((n, od)), rng.standard_normal((n, ad)), rng.standard_normal(n),
                    np.full(n_paths, T, np.int32), np.zeros(n_paths, np.uint8))
    th = (0.05 * rng.standard_normal(eng.d)).astype(np.float32)
    th[-ad:] = 0.0
    eng.set_params(th)
    eng.vf_set_state((0.1 * rng.standard_normal(eng.vf_d)).astype(np.float32))
    for rep in range(2):
        eng.compute_returns(0.995)
        eng.vf_predict()
        eng.compute_advantages(0.995, 0.97)
        eng.process_paths()
        eng.step("trpo", step_size=0.01, cg_iters=10, damping=1e-4)
    eng.close()
    # the fit on a 2e5-sample batch (3 124 Adam steps): same kernel, shorter launch
    n2 = 200000
    eng = Engine(od, ad, (128, 128), max_samples=n2 + 8, max_paths=256)
    eng.upload_flat(rng.standard_normal((n2, od)), rng.standard_normal((n2, ad)), rng.standard_normal(n2),
                    np.full(200, 1000, np.int32), np.zeros(200, np.uint8))
    eng.vf_set_state((0.1 * rng.standard_normal(eng.vf_d)).astype(np.float32))
    eng.compute_returns(0.995)
    for rep in range(2):
        eng.vf_fit(rng.permutation(n2).astype(np.int32), 64, 1e-3, 1e-3)
    eng.close()
else:
    n_paths, T, od, ad = 500, 1000, 376, 17
    n = n_paths * T
    eng = Engine(od, ad, (), max_samples=n + 8, max_paths=n_paths + 8)
    eng.upload_flat(rng.standard_normal((n, od)), rng.standard_normal((n, ad)), np.zeros(n), np.full(n_paths, T, np.int32),
                    np.zeros(n_paths, np.uint8))
    th = (0.01 * rng.standard_normal(eng.d)).astype(np.float32)
    th[-ad:] = 0.0
    eng.set_params(th)
    v = rng.standard_normal(eng.d).astype(np.float32)
    for rep in range(4):
        eng.fvp(v, 1e-4)
    eng.close()
print("ncu targets done:", which)

"""Developer tool (GPU box): is one policy step bit-reproducible?  Same inputs, same engine, repeated; with and without
a baseline fit in flight, CUDA graphs on / off (MJRL_B200_GRAPH)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import golden_paths, load_golden  # noqa: E402
from mjrl_b200.engine import Engine  # noqa: E402

g = load_golden("cheetah_24x500")
m = g["meta"]
paths = golden_paths(g)
n = int(g["path_len"].sum())
perm = np.random.RandomState(1).permutation(n).astype(np.int32)


def one(eng, with_fit):
    eng.set_params(g["theta0"], True, True)
    eng.vf_set_state(g["vf_w0"], np.zeros_like(g["vf_w0"]), np.zeros_like(g["vf_w0"]), 0)
    eng.upload_paths(paths)
    eng.compute_returns(m["gamma"])
    if with_fit:
        eng.vf_fit_begin(perm, 64, 1e-3, 1e-3)
    eng.vf_predict(prefit=True)
    eng.compute_advantages(m["gamma"], m["lam"])
    eng.process_paths()
    st = eng.step("npg", step_size=m["npg_step"], cg_iters=10, damping=1e-4)
    th = eng.get_params()
    gv, x = eng.last_vectors()
    if with_fit:
        eng.vf_fit_end()
    return th, gv, x, eng.adv_white(), st.alpha


eng = Engine(m["obs_dim"], m["act_dim"], m["hidden"], max_samples=n + 8, max_paths=64)
for with_fit in (False, True):
    ref = None
    for rep in range(4):
        out = one(eng, with_fit)
        if ref is None:
            ref = out
        else:
            print("graph=%s fit_in_flight=%s rep %d: theta %s  g %s  x %s  adv %s  alpha %s" % (
                os.environ.get("MJRL_B200_GRAPH", "1"), with_fit, rep,
                np.array_equal(out[0], ref[0]), np.array_equal(out[1], ref[1]), np.array_equal(out[2], ref[2]),
                np.array_equal(out[3], ref[3]), out[4] == ref[4]))
            if not np.array_equal(out[2], ref[2]):
                d = np.abs(out[2].astype(np.float64) - ref[2])
                print("   x differs: max abs %.3e at %d (of %d), rel-L2 %.3e" % (d.max(), int(d.argmax()), d.size,
                      np.linalg.norm(d) / np.linalg.norm(ref[2])))

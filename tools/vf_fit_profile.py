"""Developer tool: per-phase cycle breakdown of the cluster fit kernel (clock64 on CTA 0 / thread 0)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402

N, obs_dim = 200000, 17
rng = np.random.RandomState(0)
eng = Engine(obs_dim, 6, (128, 128), max_samples=N + 8, max_paths=256)
eng.upload_flat(rng.randn(N, obs_dim), rng.randn(N, 6), rng.randn(N), np.full(200, 1000, np.int32), np.zeros(200, np.uint8))
eng.compute_returns(0.995)
w = (0.1 * rng.randn(eng.vf_d)).astype(np.float32)
names_dp = ["fwd L1", "fwd L2", "out+dy", "W3 grad+delta2", "dgrad+wgrad W2", "wgrad W1", "cluster sync 1", "reduce+adam",
            "cluster sync 2", "reload+commit"]
names_mp = ["P1 L1 slice + E1 scatter", "wait h1 (E1)", "P2 L2 slice + E2", "wait y (E2)", "P3d gather issue (next step)",
            "wait dgrad (E3)", "P4 delta1 + P5 Adam + commit", "-", "P3a dy, small grads, delta2", "P3b dgrad + E3 scatter", "P3c wgrad W2"]
names_tc = ["issue L1, gather loads", "wait L1", "E1 h1 + sync", "issue L2", "wait L2", "E2a h2, y partials + sync",
            "E2b dy, dz2 + sync", "issue gW2 + dh1", "wait gW2, Adam W2 half 1", "wait dh1, E3 dz1 + sync", "issue gW1, Adam W2 half 2, vectors",
            "wait gW1", "Adam W1, stage X, st wait"]
for cl, mp in ((1, True), (16, True), (8, True), (8, False), (16, False)):
    names = names_tc if cl == 1 else (names_mp if mp else names_dp)
    eng.vf_set_state(w, np.zeros_like(w), np.zeros_like(w), 0)
    eng.vf_set_cluster(cl, mp)
    perm = rng.permutation(N).astype(np.int32)
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    eng.lib.mjb_dev_vf_profile(eng.h, None, 1)
    t0 = time.time()
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    dt = time.time() - t0
    out = (C.c_longlong * 16)()
    eng.lib.mjb_dev_vf_profile(eng.h, out, 0)
    steps = N // 64 - 1
    tot = sum(out[:13])
    print("cluster=%d model_parallel=%s: %.2f us/step wall, %d cycles/step" % (cl, mp, dt / steps * 1e6, tot // steps))
    for i, n in enumerate(names):
        if n == "-":
            continue
        print("   %-34s %7d cyc  %5.1f%%" % (n, out[i] // steps, 100.0 * out[i] / tot))

"""Developer tool: per-phase cycle breakdown of the tensor-core fit kernel (clock64 on thread 0 of the head CTA).

    python tools/vf_fit_profile.py [obs_dim]     (17 = cfg3; 39 = cfg4, one K-split helper; 376 = cfg5, six helpers)
"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402

N, obs_dim = 200000, (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
rng = np.random.RandomState(0)
eng = Engine(obs_dim, 6, (128, 128), max_samples=N + 8, max_paths=256)
eng.upload_flat(rng.randn(N, obs_dim), rng.randn(N, 6), rng.randn(N), np.full(200, 1000, np.int32), np.zeros(200, np.uint8))
eng.compute_returns(0.995)
w = (0.1 * rng.randn(eng.vf_d)).astype(np.float32)
names_tc = ["issue L1, gather loads", "wait L1", "E1 h1 + sync", "issue L2", "wait L2", "E2a h2, y partials + sync",
            "E2b dy, dz2 + sync", "issue gW2 + dh1", "wait gW2, Adam W2 half 1", "wait dh1, E3 dz1 + sync", "issue gW1, Adam W2 half 2, vectors",
            "wait gW1", "Adam W1, stage X, st wait"]
for cl, mp in ((1, True),):
    names = names_tc
    eng.vf_set_state(w, np.zeros_like(w), np.zeros_like(w), 0)
    eng.vf_set_tensor_cores(True)
    perm = rng.permutation(N).astype(np.int32)
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    print("obs_dim %d:" % obs_dim, "production instance (no counters): %.3f us per Adam step (CUDA events around the kernel, %d steps)" % (
        eng.last_fit_ms() * 1e3 / (N // 64 - 1), N // 64 - 1))
    eng.lib.mjb_dev_vf_profile(eng.h, None, 1)
    t0 = time.time()
    eng.vf_fit(perm, 64, 1e-3, 1e-3)
    dt = time.time() - t0
    out = (C.c_longlong * 16)()
    hout = (C.c_longlong * 16)()
    eng.lib.mjb_dev_vf_profile(eng.h, hout, 2)
    eng.lib.mjb_dev_vf_profile(eng.h, out, 0)
    steps = N // 64 - 1
    tot = sum(out[:16])
    print("vf_fit_tc_kernel: %.2f us/step wall, %d cycles/step" % (dt / steps * 1e6, tot // steps))
    names = names_tc + ["(top-of-step barrier: waiting for the slowest warp)", "(issue of the 6 layer-1 MMAs; 'issue L1' above is what follows)",
                        "(K-split: waiting for the helpers' partial sums; 'E1' above is what follows)"]
    for i, n in enumerate(names):
        if n == "-":
            continue
        print("   %-34s %7d cyc  %5.1f%%" % (n, out[i] // steps, 100.0 * out[i] / tot))
    if obs_dim + 4 > 32:
        hn = ["top-of-step fence + barrier", "issue z MMAs, gather loads", "wait z", "partial -> L2 + barrier", "release flag (fence)",
              "stage next X", "wait head's dz1 flag", "proxy fence, TMA 32 KB, barrier", "issue gW1", "wait gW1", "Adam slice", "(issue of the 12 z MMAs; 'issue z MMAs' above is what follows)"]
        htot = sum(hout[:12])
        print("K-split helper 0: %d cycles/step" % (htot // steps))
        for i, n in enumerate(hn):
            print("   %-34s %7d cyc  %5.1f%%" % (n, hout[i] // steps, 100.0 * hout[i] / htot))

"""Developer tool: per-phase cycle breakdown + event timing of the tensor-core linear-policy FVP kernel (cfg5 shape)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
obs_dim, act_dim = 376, 17
rng = np.random.RandomState(0)
eng = Engine(obs_dim, act_dim, (), max_samples=N + 8, max_paths=600)
T = 1000
eng.upload_flat(rng.randn(N, obs_dim).astype(np.float32), rng.randn(N, act_dim), rng.randn(N), np.full(N // T, T, np.int32),
                np.zeros(N // T, np.uint8))
d = eng.d
v = rng.randn(d).astype(np.float32)
names = ["convert + STS (waits for the loads)", "fence, sync, issue GEMM1", "issue next loads, wait GEMM1", "epilogue dy + sync",
         "issue GEMM2", "wait GEMM2"]
for tc in (True, False):
    eng.set_tensor_cores(tc)
    for _ in range(3):
        eng.fvp(v, 1e-4)
    ms = []
    for _ in range(10):
        eng.fvp(v, 1e-4)
        ms.append(eng.last_fvp_ms())
    ms = float(np.median(ms))
    print("tensor_cores=%s: %.3f ms per launch = %.0f GB/s of obs" % (tc, ms, N * obs_dim * 4 / ms / 1e6))
eng.set_tensor_cores(True)
eng.lib.mjb_dev_lin_profile(eng.h, None, 1)
eng.fvp(v, 1e-4)
out = (C.c_longlong * 16)()
eng.lib.mjb_dev_lin_profile(eng.h, out, 0)
if out[11] > 0:      # TMA-fed kernel: per-role counters
    tiles = out[11]
    roles = [("converter warp 0", ["wait GEMM2 (+flush)", "wait ring slot full", "convert + STS + release", "wait GEMM1", "dy epilogue"], range(0, 5)),
             ("MMA issuer", ["wait staged", "issue GEMM1", "wait dy", "issue GEMM2"], range(5, 9)),
             ("TMA producer", ["wait slot empty", "issue row copies"], range(9, 11))]
    for role, names_, idxs in roles:
        tot = sum(out[i] for i in idxs)
        print("%s: %d cycles per tile" % (role, tot // tiles))
        for n_, i in zip(names_, idxs):
            print("   %-32s %7d cyc  %5.1f%%" % (n_, out[i] // tiles, 100.0 * out[i] / max(tot, 1)))
else:
    tiles = out[6]
    tot = sum(out[:6])
    print("tiles %d, %d cycles per tile (thread 0 of each CTA)" % (tiles, tot // max(tiles, 1)))
    for i, n in enumerate(names):
        print("   %-62s %7d cyc  %5.1f%%" % (n, out[i] // max(tiles, 1), 100.0 * out[i] / tot))

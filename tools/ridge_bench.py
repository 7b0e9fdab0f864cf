"""Developer tool (GPU box): time of the ridge-baseline kernels (csrc/ridge.cu) at the BASELINE.json shapes, with the
algorithmic bytes / fp64 flops they move, and the numpy oracle (= the reference's algorithm) on a bounded sample beside it.

    python tools/ridge_bench.py
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mjrl_b200.engine import Engine  # noqa: E402
from oracle import npg_oracle as O  # noqa: E402
from oracle import ridge_oracle as RO  # noqa: E402


def main():
    for name, obs_dim, n_paths, T, kinds in (("cfg3 (17 obs, 1e6 timesteps)", 17, 1000, 1000, (0, 1)),
                                            ("cfg4 (39 obs, 2e5 timesteps)", 39, 1000, 200, (0, 1)),
                                            ("cfg5 (376 obs, 5e5 timesteps)", 376, 500, 1000, (0,))):
        n = n_paths * T
        rng = np.random.RandomState(0)
        eng = Engine(obs_dim, 2, (32, 32), max_samples=n + 8, max_paths=n_paths + 8)
        eng.upload_flat(rng.standard_normal((n, obs_dim)), rng.standard_normal((n, 2)), rng.standard_normal(n),
                        np.full(n_paths, T, np.int32), np.zeros(n_paths, np.uint8))
        eng.compute_returns(0.995)
        for kind in kinds:
            K = eng.ridge_features(kind)
            eng.ridge_gram(kind)
            eng.synchronize()
            reps = 5
            t0 = time.time()
            for _ in range(reps):
                G, b, yy = eng.ridge_gram(kind)
            tg = (time.time() - t0) / reps
            c = np.linalg.lstsq(G + 1e-3 * np.identity(K), b, rcond=-1)[0]
            eng.ridge_predict(kind, c)
            eng.synchronize()
            t0 = time.time()
            for _ in range(reps):
                eng.ridge_predict(kind, c)
            eng.synchronize()
            tp = (time.time() - t0) / reps
            bytes_g = n * (4 * obs_dim + 4 + 8)
            flops_g = n * (K + 1) * (K + 2)                      # upper triangle of the augmented Gram, 2 flops per entry
            # the oracle (reference algorithm, float64 numpy) on a bounded sample, extrapolated linearly in N
            ns = min(n_paths, max(4, 40000 // T))
            paths = [dict(observations=rng.standard_normal((T, obs_dim)), rewards=rng.standard_normal(T)) for _ in range(ns)]
            O.compute_returns(paths, 0.995)
            t0 = time.time()
            RO.fit(paths, kind, 1e-3)
            tc = (time.time() - t0) * (n_paths / ns)
            print("%-32s %-9s K=%4d | Gram pass %.3f ms (%.0f GB/s algorithmic, %.2f TFLOP/s fp64; incl. %d KB D2H + sync) | "
                  "predict all paths %.3f ms | numpy reference fit ~%.2f s (x%d sample)" % (
                      name, "linear" if kind == 0 else "quadratic", K, tg * 1e3, bytes_g / tg / 1e9, flops_g / tg / 1e12,
                      (K + 1) ** 2 * 8 // 1024, tp * 1e3, tc, n_paths // ns), flush=True)
        eng.close()


if __name__ == "__main__":
    main()

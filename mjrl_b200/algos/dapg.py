"""DAPG agent (mjrl/algos/dapg.py:26-141): NPG whose gradient is taken over rollout + demonstration samples
(demo weight lam_0 * lam_1^iter, everything scaled by 1e-2 and by N_all/N), Fisher products over the rollout
samples only, step 2*kl_dist without line search."""
import numpy as np

from mjrl_b200.algos.npg_cg import NPG
from mjrl_b200.engine import DEMO


class DAPG(NPG):
    algo = "dapg"

    def __init__(self, env, policy, baseline, demo_paths=None, normalized_step_size=0.01,
                 FIM_invert_args={'iters': 10, 'damping': 1e-4}, hvp_sample_frac=1.0, seed=123,
                 save_logs=False, kl_dist=None, lam_0=1.0, lam_1=0.95, **kwargs):
        self._setup(env, policy, baseline, seed, save_logs)
        self.kl_dist = kl_dist if kl_dist is not None else 0.5 * normalized_step_size
        self.FIM_invert_args = FIM_invert_args
        self.hvp_subsample = hvp_sample_frac
        self.demo_paths, self.lam_0, self.lam_1 = demo_paths, lam_0, lam_1
        self.iter_count = 0.0
        self.alpha, self.n_step_size, self.input_normalization = None, 2.0 * self.kl_dist, None

    def _use_demos(self):
        return self.demo_paths is not None and self.lam_0 > 0.0

    def _demo_samples(self):
        return int(sum(len(p["actions"]) for p in self.demo_paths)) if self._use_demos() else 0

    def _local_demos(self, eng):
        """Under data parallelism the demonstrations are sharded like the rollout samples: every rank is
        constructed with the SAME demo_paths list and keeps a contiguous range of it, so the all-reduced gradient counts
        each demonstration sample once with weight 1e-2*lam/N (dapg.py:62-74,97-98), as on one GPU."""
        if eng.world_size == 1:
            return self.demo_paths
        from mjrl_b200.parallel import shard_paths
        return shard_paths(self.demo_paths, eng.world_size, eng.rank)

    def _step_args(self):
        return dict(step_size=self.kl_dist)

    def _demo_lam(self, eng):
        """Append the demonstrations behind the rollout samples and return lam_0*lam_1^iter (dapg.py:62-66)."""
        if not self._use_demos():
            return 0.0                      # gradient over the rollout batch only, step still 2*kl_dist
        eng.upload_paths(self._local_demos(eng), which=DEMO)
        lam = self.lam_0 * (self.lam_1 ** self.iter_count)
        self.iter_count += 1
        return lam

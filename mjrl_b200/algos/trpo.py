"""TRPO agent (mjrl/algos/trpo.py:26-147): the NPG direction with step 2*kl_dist and a KL backtracking line
search (<= 100 shrinks by 0.9, acceptance on KL alone).  Each probe is one fused surrogate+KL kernel."""
from mjrl_b200.algos.npg_cg import NPG


class TRPO(NPG):
    algo = "trpo"

    def __init__(self, env, policy, baseline, kl_dist=0.01, FIM_invert_args={'iters': 10, 'damping': 1e-4},
                 hvp_sample_frac=1.0, seed=123, save_logs=False, normalized_step_size=0.01, **kwargs):
        self._setup(env, policy, baseline, seed, save_logs)
        self.kl_dist = kl_dist if kl_dist is not None else 0.5 * normalized_step_size
        self.FIM_invert_args = FIM_invert_args
        self.hvp_subsample = hvp_sample_frac
        self.alpha, self.n_step_size, self.input_normalization = None, 2.0 * self.kl_dist, None

    def _step_args(self):
        return dict(step_size=self.kl_dist)

    def _finish_step(self, eng, st, paths, t_host):
        if getattr(self, "verbose", True):
            for _ in range(st.backtracks):      # the reference prints once per shrink (trpo.py:117-118)
                print("Step size too high. Backtracking.")
        super()._finish_step(eng, st, paths, t_host)

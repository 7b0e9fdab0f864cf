"""Natural policy gradient agent with the reference's constructor and methods (mjrl/algos/npg_cg.py:24-163).
`train_from_paths` is one call into the engine: VPG kernel -> device-resident CG (10 fused FVP + update
pairs) -> step size -> parameter update -> fused surrogate/KL re-evaluation."""
import time as timer

import numpy as np

from mjrl_b200.algos.batch_reinforce import BatchREINFORCE


class NPG(BatchREINFORCE):
    algo = "npg"

    def __init__(self, env, policy, baseline, normalized_step_size=0.01, const_learn_rate=None,
                 FIM_invert_args={'iters': 10, 'damping': 1e-4}, hvp_sample_frac=1.0, seed=123,
                 save_logs=False, kl_dist=None, input_normalization=None, **kwargs):
        self._setup(env, policy, baseline, seed, save_logs)
        self.alpha = const_learn_rate
        self.n_step_size = normalized_step_size if kl_dist is None else 2.0 * kl_dist
        self.kl_dist = kl_dist if kl_dist is not None else 0.5 * normalized_step_size
        self.FIM_invert_args = FIM_invert_args
        self.hvp_subsample = hvp_sample_frac
        self.input_normalization = input_normalization
        if self.input_normalization is not None and not (0 < self.input_normalization <= 1):
            self.input_normalization = None

    # ---- reference-signature Fisher-vector product (npg_cg.py:62-88) ----
    def HVP(self, observations, actions, vector, regu_coef=None, _token=None):
        regu_coef = self.FIM_invert_args['damping'] if regu_coef is None else regu_coef
        eng = self._flat_batch(observations, actions, token=_token)
        idx = self._draw_hvp_indices(observations.shape[0], 1)
        return eng.fvp(vector, regu_coef, None if idx is None else idx[0])

    def build_Hvp_eval(self, inputs, regu_coef=None):
        """npg_cg.py:83-88.  The closure uploads its batch on the first product and reuses it for the following ones
        as long as nothing else was uploaded in between (engine generation counter + the same array objects)."""
        state = {"token": None}

        def eval(v):
            out = self.HVP(*(inputs + [v] + [regu_coef]), _token=state["token"])
            eng = self._engine
            state["token"] = (eng, eng.generation, inputs[0], inputs[1])
            return out
        return eval

    def _draw_hvp_indices(self, n_local, iters):
        """hvp_sample_frac < 0.99: np.random.choice(N, int(frac*N)) re-drawn for every product, from the global
        numpy RNG at the reference's program point (npg_cg.py:65-69)."""
        if self.hvp_subsample is None or self.hvp_subsample >= 0.99:
            return None
        eng = self._engine
        if eng is not None and eng.world_size > 1:
            # Data parallel: every rank draws the SAME global index sets (the ranks run the same program
            # with the same numpy seed, as a single-process reference run would) and keeps the entries that fall into its
            # own row range, rebased to local rows -- the union over the ranks is exactly the reference's subsample.
            from mjrl_b200.parallel import local_subsample, sample_ranges
            n_glob = eng.n_global()
            bounds = sample_ranges(n_local)
            return [local_subsample(np.random.choice(n_glob, size=int(self.hvp_subsample * n_glob)), bounds, eng.rank)
                    for _ in range(iters)]
        return np.stack([np.random.choice(n_local, size=int(self.hvp_subsample * n_local))
                         for _ in range(iters)]).astype(np.int32)

    def _normalize_inputs(self, eng, paths):
        """npg_cg.py:101-107: running average of the observation moments into policy.model ONLY (old_model keeps
        the stale transform -- reproduced literally)."""
        obs = np.concatenate([p["observations"] for p in paths])
        m = self.policy.model
        a = self.input_normalization
        if eng.world_size > 1:
            # moments over ALL ranks' samples (two all-reduced passes, like numpy's mean / population std)
            from mjrl_b200.parallel import allreduce_sum_host
            n_glob = float(eng.n_global())
            mean = allreduce_sum_host(obs.sum(axis=0), eng.cfg.device) / n_glob
            std = np.sqrt(allreduce_sum_host(((obs - mean) ** 2).sum(axis=0), eng.cfg.device) / n_glob)
        else:
            mean, std = np.mean(obs, axis=0), np.std(obs, axis=0)
        in_shift = a * m.in_shift.numpy() + (1 - a) * mean
        in_scale = a * m.in_scale.numpy() + (1 - a) * std
        m.set_transformations(in_shift, in_scale, m.out_shift.numpy(), m.out_scale.numpy())
        self._pushed = None
        self._push_policy(eng)

    def _step_args(self):
        return dict(step_size=self.n_step_size, const_learn_rate=self.alpha)

    def _demo_lam(self, eng):
        return 0.0

    def train_from_paths(self, paths):
        return self._train_resident(None, paths)

    def _train_resident(self, eng, paths):
        """train_from_paths on host path dicts, or (paths=None) on the engine's resident device batch."""
        t0 = timer.time()
        _, _, _, base_stats, self.running_score = self.process_paths(paths)
        eng = self._engine
        if self.save_logs:
            self.log_rollout_statistics(paths, base_stats)
        if getattr(self, "input_normalization", None):
            self._normalize_inputs(eng, paths)
        iters = self.FIM_invert_args['iters']
        demo_lam = self._demo_lam(eng)
        # hvp_sample_frac < 1: the device-resident CG needs every index set up front, the reference draws one per
        # product it actually evaluates (npg_cg.py:65-69, cg_solve.py:19-20 may stop early).  Draw all, and when the CG
        # stopped early rewind the global RNG and redraw only the sets that were consumed, so every later host draw
        # (the fit permutations) sees the reference's generator state.
        rng_before = np.random.get_state()
        idx = self._draw_hvp_indices(eng.n, iters)
        st = eng.step(self.algo, cg_iters=iters, damping=self.FIM_invert_args['damping'], demo_lam=demo_lam,
                      hvp_idx=idx, **self._step_args())
        if idx is not None and st.cg_iters_run < iters:
            np.random.set_state(rng_before)
            self._draw_hvp_indices(eng.n, int(st.cg_iters_run))
        self._finish_step(eng, st, paths, timer.time() - t0)
        return base_stats

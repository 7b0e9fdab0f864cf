"""Run under torchrun (one rank per GPU): shard-invariance of the engine.  Every rank holds a contiguous shard
of the golden trajectories; the all-reduced results must equal the single-GPU / reference values.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import golden_paths, load_golden, one_minus_cos, rel  # noqa: E402
from mjrl_b200.engine import Engine  # noqa: E402
from mjrl_b200.parallel import shard_bounds  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)      # a stuck collective must end with a traceback, not a hung box
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for case in ("swim_40x250", "cheetah_24x500", "linear_30x200"):
        g = load_golden(case)
        m = g["meta"]
        paths = golden_paths(g)
        lens = [len(p["rewards"]) for p in paths]
        b = shard_bounds(lens, world)
        s, e = b[rank]
        mine = paths[s:e]
        off = int(np.sum(lens[:s]))
        n_loc = int(np.sum(lens[s:e]))
        eng = Engine(m["obs_dim"], m["act_dim"], m["hidden"], max_samples=n_loc + 8, max_paths=len(mine) + 1,
                     device=local, world_size=world, rank=rank)
        eng.init_comm()
        eng.set_params(g["theta0"])
        eng.vf_set_state(g["vf_w0"], np.zeros_like(g["vf_w0"]), np.zeros_like(g["vf_w0"]), 0)
        eng.upload_paths(mine)
        eng.compute_returns(m["gamma"])
        assert np.array_equal(eng.returns(), g["returns"][off:off + n_loc])
        eng.set_advantages(g["advantages"][off:off + n_loc])
        st = eng.process_paths()
        np.testing.assert_allclose([st.mean_return, st.std_return, st.min_return, st.max_return], g["base_stats"], rtol=1e-10)
        np.testing.assert_allclose(eng.adv_white(), g["adv_white"][off:off + n_loc].astype(np.float32), atol=1e-6, rtol=0)
        assert rel(eng.vpg(), g["vpg"]) < 1e-5
        # the Fisher product's all-reduce: fused peer-memory kernel (default when every rank could map its peers) against
        # ncclAllReduce -- same local sums, rank-ordered vs NCCL's order of additions
        assert eng.p2p, "peer-memory all-reduce not enabled (cudaIpcOpenMemHandle failed?)"
        c0 = eng.p2p_calls()
        f_p2p = eng.fvp(g["fvp_vec"], m["damping"])
        assert eng.p2p_calls() == c0 + 1
        assert rel(f_p2p, g["fvp_out"]) < 1e-5
        assert eng.set_p2p(False) is False
        f_nccl = eng.fvp(g["fvp_vec"], m["damping"])
        assert eng.p2p_calls() == c0 + 1
        assert rel(f_p2p, f_nccl) < 2e-7, rel(f_p2p, f_nccl)
        x_nccl = eng.cg(g["vpg"], iters=m["cg_iters"], damping=m["damping"])
        assert eng.set_p2p(True) is True
        x = eng.cg(g["vpg"], iters=m["cg_iters"], damping=m["damping"])
        assert eng.p2p_calls() >= c0 + 2
        assert one_minus_cos(x, g["cg_x"]) < 1e-6
        assert one_minus_cos(x, x_nccl) < 1e-9
        t = torch.from_numpy(x.copy()).cuda()                       # bit-identical CG solution on every rank
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)
        st = eng.step("npg", step_size=m["npg_step"], cg_iters=m["cg_iters"], damping=m["damping"])
        new = eng.get_params()
        assert rel(new, g["npg_theta"]) < 1e-4, rel(new, g["npg_theta"])
        assert abs(st.kl_dist / g["npg_kl_dist"] - 1) < 5e-3
        # every rank ends with identical parameters
        t = torch.from_numpy(new.copy()).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)
        if "fit_perms" in g and case != "cheetah_24x500":
            err = eng.vf_fit(g["fit_perms"][:2], 64, 1e-3, 1e-3, return_errors=True)   # replicated sequential fit
            np.testing.assert_allclose(err, g["fit1_err"], rtol=2e-4)
            w = eng.vf_get_state()[0]
            assert rel(w, g["fit1_w"]) < 1e-4
        eng.close()
        if rank == 0:
            print("multigpu ok:", case, "world", world, flush=True)
    agents_section(rank, world)
    dist.destroy_process_group()


def agents_section(rank, world):
    """The drop-in classes under data parallelism: every rank runs the same program on its shard of the
    trajectories; hvp_sample_frac < 1 (global index draws, each rank keeps its range), DAPG (demonstrations sharded like
    the rollouts) and input_normalization (all-reduced observation moments) must reproduce the single-process fixtures."""
    from mjrl_b200 import runtime
    from mjrl_b200.algos.dapg import DAPG
    from mjrl_b200.algos.npg_cg import NPG
    from mjrl_b200.baselines.mlp_baseline import MLPBaseline
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    g = load_golden("swim_40x250")
    m = g["meta"]
    paths = golden_paths(g)
    for p, a in zip(paths, np.split(g["advantages"], np.cumsum(g["path_len"])[:-1])):
        p["advantages"] = a
    lens = [len(p["rewards"]) for p in paths]
    s, e = shard_bounds(lens, world)[rank]
    mine = paths[s:e]
    es = EnvSpec(m["obs_dim"], m["act_dim"], m["horizon"])
    kw = dict(FIM_invert_args={"iters": m["cg_iters"], "damping": m["damping"]})

    def fresh():
        pol = MLP(es, hidden_sizes=m["hidden"], seed=m["policy_seed"])
        bl = MLPBaseline(es, reg_coef=1e-3, batch_size=64, epochs=2, learn_rate=1e-3)
        bl.set_flat_weights(g["vf_w0"])
        return pol, bl

    th0 = g["theta0"]
    # ---- hvp_sample_frac = 0.5: the fixture drew its index sets from np.random.seed(77) ----
    pol, bl = fresh()
    agent = NPG(None, pol, bl, normalized_step_size=m["npg_step"], hvp_sample_frac=0.5, **kw)
    np.random.seed(77)
    agent.train_from_paths(mine)
    assert one_minus_cos(pol.get_param_values() - th0, g["sub_theta"] - th0) < 1e-4
    # ---- DAPG: every rank is given the SAME demonstration list and keeps its shard ----
    pol, bl = fresh()
    demo = golden_paths(g, demo=True)
    agent = DAPG(None, pol, bl, demo_paths=demo, kl_dist=0.01, lam_0=1.0, lam_1=0.95, **kw)
    agent.iter_count = 3.0
    agent.train_from_paths(mine)
    assert one_minus_cos(pol.get_param_values() - th0, g["dapg_theta"] - th0) < 1e-4
    assert abs(agent.last_step.alpha / g["dapg_alpha"] - 1) < 2e-3
    # ---- input_normalization: observation moments over ALL ranks' samples ----
    gi = load_golden("inorm_swim_40x250")
    pol, bl = fresh()
    agent = NPG(None, pol, bl, normalized_step_size=m["npg_step"], input_normalization=gi["meta"]["input_normalization"], **kw)
    agent.train_from_paths(mine)
    for k in ("in_shift", "in_scale"):
        np.testing.assert_allclose(getattr(pol.model, k).numpy(), gi["new_" + k], rtol=1e-5, atol=1e-6)
    assert one_minus_cos(pol.get_param_values() - th0, gi["theta1"] - th0) < 1e-4
    runtime.shutdown()
    if rank == 0:
        print("multigpu agents ok: world", world, flush=True)


if __name__ == "__main__":
    main()

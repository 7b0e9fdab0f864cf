"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): trajectory sharding, the unique-id / IPC-handle exchange,
broadcast helper, and the reduction semantics the engine relies on (sum of per-shard sums / global N -- not a
mean of means -- reproduces the single-process gradient, FVP and whitening statistics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mjrl_b200.parallel import local_subsample, shard_bounds, shard_paths
from oracle import npg_oracle as O


def test_shard_bounds_cover_and_balance():
    rng = np.random.RandomState(0)
    for world in (1, 2, 3, 4, 8):
        lens = rng.randint(1, 1000, size=37)
        b = shard_bounds(lens, world)
        assert b[0][0] == 0 and b[-1][1] == len(lens)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))          # contiguous, ordered
        assert all(e > s for s, e in b)                                         # nobody is empty
        per = [lens[s:e].sum() for s, e in b]
        assert max(per) - min(per) <= 2 * lens.max()
    assert shard_bounds([5, 5], 2) == [(0, 1), (1, 2)]
    assert shard_bounds([1, 1, 1000], 3) == [(0, 1), (1, 2), (2, 3)]


def test_local_subsample():
    idx = np.array([0, 5, 9, 10, 11, 19, 5])
    assert list(local_subsample(idx, [(0, 10), (10, 20)], 0)) == [0, 5, 9, 5]
    assert list(local_subsample(idx, [(0, 10), (10, 20)], 1)) == [0, 1, 9]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mjrl_b200.parallel import broadcast_bytes
        from mjrl_b200.runtime import _dist
        assert _dist() == (world, rank)
        secret = bytes(range(128))
        got = broadcast_bytes(secret if rank == 0 else b"", 128, src=0)
        assert got == secret
        # ---- reduction semantics on ragged shards ----
        obs_dim, act_dim, hidden = 5, 2, (32, 32)
        paths = O.synthetic_paths(obs_dim, act_dim, 9, 60, seed=4, ragged=True)
        spec = O.PolicySpec(obs_dim, act_dim, hidden)
        theta = O.init_policy_params(spec, 3)
        cat = lambda ps, k: np.concatenate([p[k] for p in ps])
        adv_all = np.random.RandomState(1).randn(sum(len(p["rewards"]) for p in paths))
        n_glob = adv_all.shape[0]
        bounds = shard_bounds([len(p["rewards"]) for p in paths], world)
        mine = shard_paths(paths, world, rank)
        off = sum(len(p["rewards"]) for p in paths[:bounds[rank][0]])
        adv_loc = adv_all[off:off + sum(len(p["rewards"]) for p in mine)]
        # whitening statistics: all-reduce of sums, two passes (what mjb_process_paths does)
        t = torch.tensor([adv_loc.sum()], dtype=torch.float64)
        dist.all_reduce(t)
        mean = float(t) / n_glob
        t = torch.tensor([((adv_loc - mean) ** 2).sum()], dtype=torch.float64)
        dist.all_reduce(t)
        std = float(np.sqrt(float(t) / n_glob))
        assert abs(mean - adv_all.mean()) < 1e-12 and abs(std - adv_all.std()) < 1e-12
        white_all = O.whiten(adv_all)
        white_loc = (adv_loc - mean) / (std + 1e-6)
        # gradient: local sum / N_global, then all-reduce
        n_loc = adv_loc.shape[0]
        g = O.flat_vpg(spec, theta, cat(mine, "observations"), cat(mine, "actions"), white_loc) * (n_loc / n_glob)
        tg = torch.from_numpy(g.copy())
        dist.all_reduce(tg)
        g_full = O.flat_vpg(spec, theta, cat(paths, "observations"), cat(paths, "actions"), white_all)
        assert np.linalg.norm(tg.numpy() - g_full) / np.linalg.norm(g_full) < 1e-10
        # FVP: data part scaled by n_loc/N_global, the data-free log_std block and damping added once
        v = np.random.RandomState(2).randn(spec.d)
        f_loc = O.fvp(spec, theta, cat(mine, "observations"), v, 0.0)
        f_loc[:-act_dim] *= n_loc / n_glob
        f_loc[-act_dim:] /= world
        tf = torch.from_numpy(f_loc.copy())
        dist.all_reduce(tf)
        f_full = O.fvp(spec, theta, cat(paths, "observations"), v, 0.0)
        assert np.linalg.norm(tf.numpy() - f_full) / np.linalg.norm(f_full) < 1e-10
        # ---- hvp_sample_frac < 1 (npg_cg.py:65-69): every rank draws the SAME global indices and keeps its own range;
        #      the union of the local lists is the reference's subsample, the all-reduced product equals the gathered one
        from mjrl_b200.parallel import allreduce_sum_host, sample_ranges
        ranges = sample_ranges(n_loc)
        assert ranges[rank] == (off, off + n_loc) and ranges[-1][1] == n_glob
        gidx = np.random.RandomState(9).choice(n_glob, size=n_glob // 2)
        lidx = local_subsample(gidx, ranges, rank)
        cnt = allreduce_sum_host(np.array([float(len(lidx))]))
        assert int(cnt[0]) == len(gidx)
        f_sub = O.fvp(spec, theta, cat(mine, "observations")[lidx], v, 0.0)
        f_sub[:-act_dim] *= len(lidx) / len(gidx)
        f_sub[-act_dim:] /= world
        f_sub_full = O.fvp(spec, theta, cat(paths, "observations")[gidx], v, 0.0)
        assert np.linalg.norm(allreduce_sum_host(f_sub) - f_sub_full) / np.linalg.norm(f_sub_full) < 1e-10
        # ---- input_normalization moments (npg_cg.py:101-107) over all ranks' samples, two all-reduced passes
        obs_loc, obs_all = cat(mine, "observations"), cat(paths, "observations")
        mean_o = allreduce_sum_host(obs_loc.sum(axis=0)) / n_glob
        std_o = np.sqrt(allreduce_sum_host(((obs_loc - mean_o) ** 2).sum(axis=0)) / n_glob)
        assert np.allclose(mean_o, obs_all.mean(axis=0), rtol=0, atol=1e-12) and np.allclose(std_o, obs_all.std(axis=0), rtol=0, atol=1e-12)
        # ---- DAPG demonstrations are sharded like the rollouts: every demo path lands on exactly one rank
        demo = O.synthetic_paths(obs_dim, act_dim, 3, 40, seed=8)
        owned = allreduce_sum_host(np.array([float(sum(len(p["rewards"]) for p in shard_paths(demo, world, rank)))]))
        assert int(owned[0]) == sum(len(p["rewards"]) for p in demo)
        assert shard_bounds([7], 2) == [(0, 1), (1, 1)]           # fewer paths than ranks: the last ranks stay empty
        # ---- peer-memory all-reduce setup (engine.init_p2p): 64-byte IPC handles gathered in rank order; the fused kernel is
        #      enabled only when EVERY rank could import (a rank on NCCL and one on peer memory would wait forever)
        from mjrl_b200.parallel import all_gather_bytes, all_ranks_agree
        handles = all_gather_bytes(bytes([rank + 1]) * 64, 64)
        assert handles == [bytes([r + 1]) * 64 for r in range(world)]
        assert all_ranks_agree(True) is True
        assert all_ranks_agree(rank != 1) is False
        # ---- ridge baselines (ridge.cu): the Gram matrix of [F | y] is additive over the shards
        from oracle import ridge_oracle as RO
        for p_ in paths:
            p_["returns"] = O.discount_sum(p_["rewards"], 0.99)
        for kind in (0, 1):
            aug = lambda ps: np.concatenate([RO.features(ps, kind), cat(ps, "returns")[:, None]], axis=1)
            Fl, Fa = aug(mine), aug(paths)
            G = allreduce_sum_host(Fl.T.dot(Fl))
            assert np.linalg.norm(G - Fa.T.dot(Fa)) / np.linalg.norm(Fa.T.dot(Fa)) < 1e-12
        out.put((rank, "ok"))
    except Exception as exc:      # surface the failure in the parent
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

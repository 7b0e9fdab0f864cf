"""Data-parallel plumbing (one process per GPU): trajectories are sharded by contiguous ranges of
paths balanced by timesteps, so every GAE scan is rank-local and the global sample order (rank 0's paths, then
rank 1's, ...) equals the single-GPU order.  Reductions that cross ranks (flat gradient, every FVP result,
whitening / return statistics, surrogate & KL sums) are NCCL all-reduces issued by the engine; this module only
holds the host-side helpers."""
import numpy as np


def shard_bounds(path_lengths, world_size):
    """Contiguous [start, end) path ranges per rank, balanced by the number of timesteps."""
    lens = np.asarray(path_lengths, dtype=np.int64)
    n = len(lens)
    if world_size <= 1:
        return [(0, n)]
    if n < world_size:                      # fewer paths than ranks: one path each, the last ranks stay empty
        return [(min(r, n), min(r + 1, n)) for r in range(world_size)]
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        # never hand out an empty shard while paths remain
        k = min(max(k, cuts[-1] + 1), n - (world_size - r))
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def shard_paths(paths, world_size, rank):
    s, e = shard_bounds([len(p["rewards"]) for p in paths], world_size)[rank]
    return paths[s:e]


def broadcast_bytes(payload, nbytes, src=0, device=None):
    """Broadcast a small byte string (the NCCL unique id) over the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    buf = bytearray(payload) if dist.get_rank() == src else bytearray(nbytes)
    t = torch.frombuffer(buf, dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(buf, n, device=None):
    """Every rank contributes n bytes; returns the list of all ranks' byte strings in rank order."""
    import torch
    import torch.distributed as dist
    t = torch.frombuffer(bytearray(buf[:n].ljust(n, b"\0")), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def all_ranks_agree(flag, device=None):
    """True iff `flag` is true on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], dtype=torch.int32)
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.cpu()[0]) == 1)


def local_subsample(global_idx, bounds_samples, rank):
    """hvp_sample_frac < 1 under data parallelism: the host draws global sample indices once; each rank keeps the
    ones inside its own sample range, rebased to local row numbers."""
    lo, hi = bounds_samples[rank]
    idx = np.asarray(global_idx)
    keep = idx[(idx >= lo) & (idx < hi)] - lo
    return keep.astype(np.int32)


def sample_ranges(n_local):
    """[lo, hi) global row range of every rank, from the ranks' local sample counts (rank order = global sample order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.zeros(world, dtype=torch.int64)
    t[dist.get_rank()] = int(n_local)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    cnt = t.cpu().numpy()
    hi = np.cumsum(cnt)
    return [(int(h - c), int(h)) for c, h in zip(cnt, hi)]


def allreduce_sum_host(x, device=None):
    """Sum a small float64 host array over the ranks (default torch.distributed group)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t)
    return t.cpu().numpy()

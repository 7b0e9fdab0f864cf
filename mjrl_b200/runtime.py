"""Engine registry: one CUDA engine per (policy shape, baseline shape) in this process / on this rank.

The reference has no such object -- its policy, baseline and agent talk through numpy arrays on the host.
Here they share one device-resident engine so a rollout batch is uploaded once per train_step."""
import os

import numpy as np

from mjrl_b200.engine import Engine

_engines = {}


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except Exception:
        pass
    return 1, 0


def device_ordinal():
    """CUDA ordinal of this process' engine (LOCAL_RANK under torchrun, MJRL_B200_DEVICE otherwise)."""
    world, _ = _dist()
    return int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else int(os.environ.get("MJRL_B200_DEVICE", "0"))


def get_engine(obs_dim, act_dim, hidden=None, vf_hidden=(128, 128), min_log_std=-3.0, need_samples=0, need_paths=0):
    """Return (creating or growing as needed) the engine for this shape.  hidden=None matches any policy
    shape with the same obs/act dims (used by stand-alone baselines / process_samples calls)."""
    world, rank = _dist()
    device = device_ordinal()
    key = None
    if hidden is None:
        for k in _engines:
            if k[0] == obs_dim and k[1] == act_dim and k[3] == tuple(vf_hidden):
                key = k
                break
        if key is None:
            hidden = ()
    if key is None:
        key = (obs_dim, act_dim, tuple(hidden), tuple(vf_hidden), float(min_log_std))
    eng = _engines.get(key)
    if eng is not None and (need_samples > eng.max_samples or need_paths > eng.max_paths):
        state = (eng.get_params(), eng.get_params(old=True), eng.vf_get_state())
        eng.close()
        eng = None
    else:
        state = None
    if eng is None:
        cap = max(1 << 14, int(need_samples * 1.25) + 64)
        pcap = max(256, int(need_paths * 1.25) + 8)
        eng = Engine(key[0], key[1], key[2], key[3], key[4], max_samples=cap, max_paths=pcap, device=device,
                     world_size=world, rank=rank)
        eng.init_comm()
        if state is not None:
            eng.set_params(state[0], True, False)
            eng.set_params(state[1], False, True)
            w, m, v, step = state[2]
            eng.vf_set_state(w, m, v, step)
        _engines[key] = eng
    return eng


def _draw_permutations(key, pos, n, count):
    """`count` consecutive np.random.permutation(n) draws from the MT19937 state (key, pos): returns (perms [count, n]
    int32, key, pos) with the state advanced exactly as numpy would have (`mjb_host_permutation`, pinned against numpy in
    tests/test_abi.py).  The foreign call releases the GIL."""
    import ctypes as C
    from mjrl_b200 import _native
    lib = _native.load()
    out = np.empty((count, n), dtype=np.int32)
    cpos = C.c_int32(int(pos))
    for i in range(count):
        rc = lib.mjb_host_permutation(key.ctypes.data_as(C.c_void_p), C.byref(cpos), int(n), out[i].ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("mjb_host_permutation failed (%d)" % rc)
    return out, key, int(cpos.value)


def _same_rng_state(a, b):
    return a[0] == b[0] and int(a[2]) == int(b[2]) and int(a[3]) == int(b[3]) and float(a[4]) == float(b[4]) and \
        np.array_equal(a[1], b[1])


_speculation = None        # the permutations the NEXT fit will draw if nobody touches numpy's global RNG until then


def _speculate(state, n, count):
    """Start computing, on a worker thread and from a COPY of the RNG state, the `count` permutations the next fit of the
    same size would draw.  They are used only if numpy's global state is still exactly `state` at that point."""
    import os
    import threading
    global _speculation
    if os.environ.get("MJRL_B200_PERM_SPECULATE", "1") == "0":
        _speculation = None
        return
    box = {"snap": state, "n": int(n), "count": int(count), "result": None}

    def work():
        try:
            key = np.array(state[1], dtype=np.uint32, copy=True, order="C")
            box["result"] = _draw_permutations(key, int(state[2]), box["n"], box["count"])
        except Exception:          # the regular draw will run (and raise, if it must) at the point of use
            box["result"] = None

    box["thread"] = threading.Thread(target=work, name="mjrl_b200-perm", daemon=True)
    box["thread"].start()
    _speculation = box


def global_permutations(n, count=1, speculate_next=True):
    """`count` x np.random.permutation(n) as int32 [count, n], drawn from numpy's GLOBAL RandomState at this program point
    (optimize_model.py:22, one per epoch) -- same values, same RNG state afterwards -- through the library's batched
    MT19937 / Fisher-Yates loops, which are 2-3x faster than numpy's element-wise loop.

    The draw of 1e6 indices still costs ~5 ms of host time in front of the sequential fit chain, which is the critical
    path of a train step.  So after every draw a worker thread computes, from a copy of the new RNG state, what the next
    call with the same (n, count) would draw; the next call takes that result only if numpy's global state is
    bit-for-bit the state the speculation started from (nobody drew a random number in between) and then installs the
    advanced state.  Values and RNG stream are the reference's in every case; a miss just draws on the spot."""
    global _speculation
    st = np.random.get_state()
    n, count = int(n), int(count)
    if st[0] != "MT19937" or n < 2 or n > 0x7fffffff:
        _speculation = None
        return np.stack([np.random.permutation(np.arange(n, dtype=np.int32)) for _ in range(count)])
    res, sp, _speculation = None, _speculation, None
    if sp is not None:
        sp["thread"].join()
        if sp["result"] is not None and sp["n"] == n and sp["count"] == count and _same_rng_state(sp["snap"], st):
            res = sp["result"]
    if res is None:
        res = _draw_permutations(np.array(st[1], dtype=np.uint32, copy=True, order="C"), int(st[2]), n, count)
    perms, key, pos = res
    np.random.set_state(("MT19937", key, pos, st[3], st[4]))
    if speculate_next:
        _speculate(np.random.get_state(), n, count)
    return perms


def global_permutation(n):
    """One np.random.permutation(n) (int32) from numpy's global RandomState -- see global_permutations."""
    return global_permutations(n, 1, speculate_next=False)[0]


class session:
    """`with runtime.session(eng, paths):` uploads `paths` ONCE (always -- the reference reads the arrays it is handed,
    so a new call never trusts what a previous call left on the device) and pins that list object as the engine's
    rollout batch for the duration of the block: nested helpers that receive the same list (`returns_on`,
    `advantages_on`, `fit_begin`, `process_paths`, `train_from_paths`) then skip their own upload.  The pin is a strong
    reference compared with `is`, never an id(), and it is dropped when the block exits."""

    def __init__(self, eng, paths):
        self.eng, self.paths = eng, paths

    def __enter__(self):
        self.eng.upload_paths(self.paths)
        self.eng.session_paths = self.paths
        return self.eng

    def __exit__(self, *exc):
        self.eng.session_paths = None
        return False


def ensure_resident(eng, paths, force=False):
    """Make `paths` the engine's rollout batch.  Outside a `session` this always uploads; inside one it uploads only
    when `paths` is not the pinned list (or when forced)."""
    if force or getattr(eng, "session_paths", None) is not paths:
        eng.upload_paths(paths)
        if getattr(eng, "session_paths", None) is not None:
            eng.session_paths = None          # another batch replaced the pinned one: the pin no longer holds
    return eng


def shutdown():
    for e in _engines.values():
        e.close()
    _engines.clear()

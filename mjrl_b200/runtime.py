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


def global_permutation(n):
    """np.random.permutation(n) as int32, drawn from numpy's GLOBAL RandomState at this program point
    (optimize_model.py:22) -- same order, same RNG state afterwards -- through the library's batched MT19937 /
    Fisher-Yates loops (`mjb_host_permutation`), which are 2-3x faster than numpy's element-wise loop."""
    import ctypes as C
    from mjrl_b200 import _native
    st = np.random.get_state()
    if st[0] != "MT19937" or n < 2 or n > 0x7fffffff:
        return np.random.permutation(np.arange(n, dtype=np.int32))
    key = np.array(st[1], dtype=np.uint32, copy=True, order="C")
    pos = C.c_int32(int(st[2]))
    out = np.empty(n, dtype=np.int32)
    rc = _native.load().mjb_host_permutation(key.ctypes.data_as(C.c_void_p), C.byref(pos), int(n), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("mjb_host_permutation failed (%d)" % rc)
    np.random.set_state(("MT19937", key, int(pos.value), st[3], st[4]))
    return out


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

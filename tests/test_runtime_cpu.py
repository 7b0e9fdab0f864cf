"""CPU: residency rules of mjrl_b200.runtime (round-1 advisor finding: never key a device batch on id()).  A stand-in engine
counts uploads; the rules are about object identity held by a strong reference, so no GPU is involved."""
import numpy as np

from mjrl_b200 import runtime


class CountingEngine:
    def __init__(self):
        self.uploads = []
        self.session_paths = None

    def upload_paths(self, paths):
        self.uploads.append(paths)


def make_paths(seed=0):
    rng = np.random.RandomState(seed)
    return [dict(observations=rng.randn(5, 3), actions=rng.randn(5, 2), rewards=rng.randn(5)) for _ in range(3)]


def test_outside_a_session_every_call_uploads():
    eng, paths = CountingEngine(), make_paths()
    runtime.ensure_resident(eng, paths)
    runtime.ensure_resident(eng, paths)                    # same list object: still uploaded again (the arrays may have changed)
    assert len(eng.uploads) == 2


def test_session_uploads_once_and_pins_by_identity():
    eng, paths = CountingEngine(), make_paths()
    with runtime.session(eng, paths):
        assert len(eng.uploads) == 1 and eng.session_paths is paths
        runtime.ensure_resident(eng, paths)                # nested helper, same list: no second upload
        assert len(eng.uploads) == 1
        runtime.ensure_resident(eng, list(paths))          # a fresh list over the same dicts is NOT the pinned object
        assert len(eng.uploads) == 2 and eng.session_paths is None
        runtime.ensure_resident(eng, paths)                # the pin is gone: uploads again
        assert len(eng.uploads) == 3
    assert eng.session_paths is None


def test_forced_upload_inside_a_session_unpins():
    eng, paths = CountingEngine(), make_paths()
    with runtime.session(eng, paths):
        runtime.ensure_resident(eng, paths, force=True)
        assert len(eng.uploads) == 2 and eng.session_paths is None


def test_recycled_ids_cannot_alias_a_batch():
    """The failure mode of the id() fingerprint: a new list at a recycled address.  The pin holds a strong reference, so the
    address of the pinned list cannot be reused while the session is open; after it closes nothing is trusted."""
    eng = CountingEngine()
    for seed in range(20):
        paths = make_paths(seed)
        with runtime.session(eng, paths):
            runtime.ensure_resident(eng, paths)
        del paths
    assert len(eng.uploads) == 20

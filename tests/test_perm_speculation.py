"""CPU: the fit's minibatch permutations (runtime.global_permutations) are numpy's, with numpy's RNG stream, whether the
speculative pre-draw of the next step's permutations hits (nobody touched the global RNG in between), misses (someone did,
or the batch size changed) or is switched off."""
import threading

import numpy as np
import pytest

from mjrl_b200 import runtime


def numpy_draws(n, count):
    return np.stack([np.random.permutation(n) for _ in range(count)]).astype(np.int32)


@pytest.mark.parametrize("speculate", ["1", "0"])
def test_sequence_equals_numpy(speculate, monkeypatch):
    monkeypatch.setenv("MJRL_B200_PERM_SPECULATE", speculate)
    runtime._speculation = None
    script = [(5000, 2, None), (5000, 2, None), (5000, 2, "rand"), (5000, 2, None), (4000, 2, None), (4000, 1, None),
              (4000, 1, "normal"), (4000, 1, None), (4000, 1, None)]
    np.random.seed(123)
    want = []
    for n, c, between in script:
        want.append(numpy_draws(n, c))
        if between == "rand":
            np.random.rand(3)
        elif between == "normal":
            np.random.randn(1)                  # leaves a cached gaussian in the state: part of the comparison
    want_state = np.random.get_state()
    np.random.seed(123)
    main_draws = []
    orig = runtime._draw_permutations

    def counting(key, pos, n, count):
        if threading.current_thread() is threading.main_thread():
            main_draws.append((n, count))
        return orig(key, pos, n, count)

    monkeypatch.setattr(runtime, "_draw_permutations", counting)
    for (n, c, between), w in zip(script, want):
        got = runtime.global_permutations(n, c)
        assert got.dtype == np.int32 and np.array_equal(got, w)
        if between == "rand":
            np.random.rand(3)
        elif between == "normal":
            np.random.randn(1)
    got_state = np.random.get_state()
    assert got_state[0] == want_state[0] and np.array_equal(got_state[1], want_state[1]) and got_state[2:] == want_state[2:]
    if speculate == "1":
        # hits: calls 2, 4 (same size, untouched RNG), 9; misses: 1 (cold), 3?  -- call 3 follows call 2 untouched -> hit;
        # call 4 follows the rand() -> miss; 5 (new n), 6 (new count), 8 (after randn) miss; 7, 9 hit
        assert main_draws == [(5000, 2), (5000, 2), (4000, 2), (4000, 1), (4000, 1)]
    else:
        assert len(main_draws) == len(script)
    if runtime._speculation is not None:
        runtime._speculation["thread"].join()
    runtime._speculation = None


def test_single_draw_api_matches_numpy():
    np.random.seed(7)
    a = np.random.permutation(1000).astype(np.int32)
    s1 = np.random.get_state()
    np.random.seed(7)
    b = runtime.global_permutation(1000)
    s2 = np.random.get_state()
    assert np.array_equal(a, b) and np.array_equal(s1[1], s2[1]) and s1[2] == s2[2]

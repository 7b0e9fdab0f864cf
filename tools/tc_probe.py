"""Run the tcgen05 probe kernel for the four operand-major combinations and compare with numpy."""
import ctypes as C
import os
import sys

import numpy as np

lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtcprobe.so"))
lib.tc_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
rng = np.random.RandomState(0)
ok = True
for K in (16, 32, 128):
    A = rng.randn(128, K).astype(np.float16)
    B = rng.randn(128, K).astype(np.float16)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    for mode in range(4):
        a = np.ascontiguousarray(A.T if mode & 1 else A)
        b = np.ascontiguousarray(B.T if mode & 2 else B)
        D = np.zeros((128, 128), np.float32)
        rc = lib.tc_probe(a.ctypes.data, b.ctypes.data, D.ctypes.data, K, mode)
        err = np.abs(D - want).max() / np.abs(want).max()
        print("K=%3d mode=%d (A %s-major, B %s-major): rc=%d max rel err %.2e" % (
            K, mode, "MN" if mode & 1 else "K", "MN" if mode & 2 else "K", rc, err))
        ok &= rc == 0 and err < 1e-3
print("PROBE", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

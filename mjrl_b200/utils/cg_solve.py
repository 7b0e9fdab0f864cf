"""Host-side conjugate gradient with the reference's call signature (mjrl/utils/cg_solve.py:3-22) for code
that passes its own operator.  The NPG/TRPO/DAPG agents do not use this: their CG runs on the device
(`Engine.cg` / `mjb_policy_cg`), one fused FVP + vector-update pair per iteration."""
import numpy as np


def cg_solve(f_Ax, b, x_0=None, cg_iters=10, residual_tol=1e-10):
    # the reference ignores x_0 (starts from zero) -- kept for signature compatibility
    sol = np.zeros_like(b)
    resid = b.copy()
    direction = b.copy()
    rr = resid.dot(resid)
    for _ in range(cg_iters):
        Ad = f_Ax(direction)
        step = rr / direction.dot(Ad)
        sol += step * direction
        resid -= step * Ad
        rr_next = resid.dot(resid)
        direction = resid + (rr_next / rr) * direction
        rr = rr_next
        if rr < residual_tol:
            break
    return sol

"""Developer tool: parity metrics of every baseline-fit kernel against the golden fixtures (same checks as the tests)."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import golden_paths, load_golden, rel  # noqa: E402
from mjrl_b200.engine import Engine  # noqa: E402

for case in ["pm_5x50", "pm_40x25_ragged", "swim_40x250", "cheetah_24x500", "linear_30x200"]:
    g = load_golden(case)
    paths = golden_paths(g)
    m = g["meta"]
    n = int(g["path_len"].sum())
    for cl in (True, False):
        eng = Engine(m["obs_dim"], m["act_dim"], m["hidden"], max_samples=n + 8, max_paths=len(g["path_len"]) + 1)
        eng.set_params(g["theta0"])
        eng.vf_set_state(g["vf_w0"], np.zeros_like(g["vf_w0"]), np.zeros_like(g["vf_w0"]), 0)
        eng.vf_set_tensor_cores(cl)
        eng.upload_paths(paths)
        eng.compute_returns(m["gamma"])
        err = eng.vf_fit(g["fit_perms"][:2], 64, 1e-3, 1e-3, return_errors=True)
        w1 = eng.vf_get_state()[0]
        eng.vf_fit(g["fit_perms"][2:4], 64, 1e-3, 1e-3)
        w2, mm, vv, step = eng.vf_get_state()
        eng.vf_predict()
        pd = np.abs(eng.baseline() - g["fit2_predict"]).max()
        print("%-18s tensor_cores=%-5s err_rel %.1e  fit1_w %.2e  fit2_w %.2e  fit2_v %.2e  predict max|d| %.2e (scale %.2f)" % (
            case, cl, abs(err[1] - g["fit1_err"][1]) / abs(g["fit1_err"][1]), rel(w1, g["fit1_w"]), rel(w2, g["fit2_w"]),
            rel(vv, g["fit2_v"]), pd, np.abs(g["fit2_predict"]).max()))
        eng.close()

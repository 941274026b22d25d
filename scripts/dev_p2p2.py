"""Dev check (GPU): near-field kernel variants at N = 1M: time and accuracy vs the default kernel."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import fmm_bem_relaxed_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pts, q = F.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(8), pts)
plan.set_option("overlap_p2p", 0)
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
base = None
for name, opts in (("pair2", {"p2p_kernel": 2}), ("tma", {"p2p_kernel": 3, "p2p_newton": 0}),
                   ("tma+newton", {"p2p_kernel": 3, "p2p_newton": 1})):
    for k, v in opts.items():
        plan.set_option(k, v)
    ts = []
    for _ in range(5):
        r = plan.execute(q)
        ts.append(plan.phase_times()["p2p"])
    if base is None:
        base = r
    pairs = plan.info().n_p2p_body_pairs
    t = min(ts[1:])
    print("%-12s p2p %.4f ms  %.2f TFLOP/s algorithmic  pot diff %.2e force diff %.2e" % (
        name, t, 22.0 * pairs / (t * 1e-3) / 1e12, rel(r[:, 0], base[:, 0]), rel(r[:, 1:], base[:, 1:])), flush=True)

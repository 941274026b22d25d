"""Dev check (GPU): pair kernel compiled for 20 / 24 / 28 / 32 resident warps per SM."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import fmm_bem_relaxed_b200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pts, q = F.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(8), pts)
plan.set_option("overlap_p2p", 0)
pairs = plan.info().n_p2p_body_pairs
ref = None
for occ in (20, 24, 28, 32):
    plan.set_option("p2p_occ", occ)
    ts = []
    for _ in range(5):
        r = plan.execute(q)
        ts.append(plan.phase_times()["p2p"])
    if ref is None:
        ref = r
    t = min(ts[1:])
    print("occ %d: p2p %.4f ms  %.2f TFLOP/s  same bits %s" % (occ, t, 22.0 * pairs / (t * 1e-3) / 1e12, np.array_equal(r, ref)), flush=True)

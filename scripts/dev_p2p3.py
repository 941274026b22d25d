"""Dev check (GPU): near-field option sweep at N = 1M (serialised phase time of the pair kernel)."""
import sys, os, itertools
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import fmm_bem_relaxed_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pts, q = F.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(8), pts)
plan.set_option("overlap_p2p", 0)
pairs = plan.info().n_p2p_body_pairs
for items, kern, unroll, newton in ((0, 2, 4, 0), (0, 2, 8, 0), (1, 2, 4, 0), (1, 2, 8, 0), (0, 3, 4, 1), (1, 3, 4, 1), (0, 1, 4, 0), (0, 1, 8, 0)):
    plan.set_option("p2p_items", items)
    plan.set_option("p2p_kernel", kern)
    plan.set_option("p2p_unroll", unroll)
    plan.set_option("p2p_newton", newton)
    ts = []
    for _ in range(5):
        plan.execute(q)
        ts.append(plan.phase_times()["p2p"])
    t = min(ts[1:])
    print("items %d kernel %d unroll %d newton %d: p2p %.4f ms  %.2f TFLOP/s" % (items, kern, unroll, newton, t,
          22.0 * pairs / (t * 1e-3) / 1e12), flush=True)

"""Dev check (GPU): P2M kernels (full tile vs narrow tile): upward phase time and bit-identity of the multipoles."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import fmm_bem_relaxed_b200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pts, q = F.drand48_inputs(n)
for P in (8, 5, 12):
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts)
    plan.set_option("overlap_p2p", 0)
    out = {}
    for k in (0, 1):
        plan.set_option("p2m_kernel", k)
        ts = []
        for _ in range(4):
            r = plan.execute(q)
            ts.append(plan.phase_times()["upward"])
        out[k] = (min(ts[1:]), r, plan.expansions()[0])
    print("P=%d upward phase: full tile %.4f ms, narrow tile %.4f ms; same result bits %s, same multipole bits %s" % (
        P, out[0][0], out[1][0], np.array_equal(out[0][1], out[1][1]), np.array_equal(out[0][2], out[1][2])), flush=True)
    out = {}
    for k in (0, 1):
        plan.set_option("l2p_kernel", k)
        ts = []
        for _ in range(4):
            r = plan.execute(q)
            ts.append(plan.phase_times()["downward"])
        out[k] = (min(ts[1:]), r)
    print("P=%d downward phase: leaf per warp %.4f ms, four leaves per warp %.4f ms; same result bits %s" % (
        P, out[0][0], out[1][0], np.array_equal(out[0][1], out[1][1])), flush=True)

"""Development: host-buffer BEM matvec times per expansion order on the C2 mesh (what one GMRES iteration pays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

verts = O.unit_sphere(7)
n = len(verts)
t0 = time.perf_counter()
plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(verts))
print("plan %.1f ms, panels %d" % ((time.perf_counter() - t0) * 1e3, n))
q = np.ones(n)
sched = [8, 8, 6, 5, 5, 5, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1]
for rnd in range(3):
    tot = 0
    line = []
    for p in sched:
        plan.kernel().set_p(p)
        t0 = time.perf_counter()
        plan.execute(q)
        dt = (time.perf_counter() - t0) * 1e3
        tot += dt
        line.append("%d:%.2f" % (p, dt))
    print("round %d total %.2f ms | %s" % (rnd, tot, " ".join(line)))
    ph = plan.phase_times()
    print("   last phases", {k: round(v, 3) for k, v in ph.items()})

"""Dev check (GPU): where the time of the relaxed GMRES solve of config C2 goes (LaplaceBEM sphere, device-resident
fmmb_gmres): first solve of a plan against later ones, matvec time and launch count per order.
usage: python scripts/dev_gmres.py [recursions] [solves]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

rec = int(sys.argv[1]) if len(sys.argv) > 1 else 7
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 4
v = O.unit_sphere(rec)
n = len(v)
t0 = time.perf_counter()
plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 0))
t1 = time.perf_counter()
b = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 1)).execute(np.ones(n))
if "BEM_NEAR" in os.environ:
    plan.set_option("bem_near_kernel", int(os.environ["BEM_NEAR"]))
print("panels %d  plan build %.4fs" % (n, t1 - t0), flush=True)
if os.environ.get("FIRST_USE"):
    fresh = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 0))
    xx = np.ones(n)
    for p in (8, 6, 5, 4, 3, 2, 1):
        fresh.kernel().set_p(p)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            fresh.execute(xx)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("first use p=%d: call 1 %.3f ms (tables, buffers)  call 2 %.3f (capture)  call 3 %.3f  call 4 %.3f" % (p, *ts), flush=True)
    fresh.close()
so = F.SolverOptions(residual=1e-6, max_iters=200, restart=200, max_p=8)
for k in range(solves):
    plan.kernel().set_p(8)
    t0 = time.perf_counter()
    rep = F.GMRES(plan, np.zeros(n), b, so)
    dt = time.perf_counter() - t0
    print("solve %d: %.3f ms  iterations %d  residual %.3e  schedule %s" % (
        k, dt * 1e3, rep["iterations"], rep["final_residual"], rep["p_schedule"]), flush=True)
x = rep["x"]
for p in range(8, 0, -1):
    plan.kernel().set_p(p)
    for _ in range(3):
        plan.execute(x)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        plan.execute(x)
        ts.append(time.perf_counter() - t0)
    t = plan.phase_times()
    print("p=%d host call %.3f ms (min %.3f)  device total %.3f up %.3f m2l %.3f down %.3f p2p %.3f launches %d" % (
        p, np.median(ts) * 1e3, min(ts) * 1e3, t["total"], t["upward"], t["m2l"], t["downward"], t["p2p"], t["launches"]), flush=True)

"""Small matvecs of every kernel family for compute-sanitizer (memcheck / racecheck): LaplaceSpherical through the
class-major engine (P = 6), the fused sweep engine (P = 10 and m2l_mode 3 at P = 5) and the treecode evaluator; a
LaplaceSphericalBEM plan (cached near field, two-set far field) and its near-field-only variant; a stresslet plan.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmm_bem_relaxed_b200 as F

n = 12000
pts, q = F.drand48_inputs(n)
rng = np.random.default_rng(1)
base = None
for P, mode, ev in ((6, 0, F.FMMOptions.FMM), (10, 0, F.FMMOptions.FMM), (5, 3, F.FMMOptions.FMM), (6, 1, F.FMMOptions.FMM),
                    (6, 0, F.FMMOptions.TREECODE)):
    o = F.FMMOptions()
    o.m2l_mode = mode
    o.evaluator = ev
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts, o)
    for _ in range(3):                               # third call replays the captured graph
        r = plan.execute(q)
    d = F.Direct.matvec(plan, q, pts[:200])
    err = float(np.linalg.norm(r[:200, 0] - d[:, 0]) / np.linalg.norm(d[:, 0]))
    print("laplace P=%d mode=%d eval=%d: pot error vs direct %.2e" % (P, mode, ev, err), flush=True)
    assert err < 1e-2
    plan.close()

verts = None
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import oracle_lib as O                            # mesh generator only (this script is test infrastructure)
    verts = O.unit_sphere(4)
except Exception as e:
    print("no mesh generator:", e)
if verts is not None:
    m = len(verts)
    for near_only in (False, True):
        o = F.FMMOptions()
        o.local_evaluation = near_only
        plan = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(verts), o)
        for p in (6, 4, 6):
            plan.kernel().set_p(p)
            r = plan.execute(np.ones(m))
        print("laplace-bem near_only=%s: sum %.6e" % (near_only, float(r.sum())), flush=True)
        plan.close()
g = np.hstack([rng.random((n, 3)), np.tile([0.0, 1.0, 0.0], (n, 1))])
plan = F.FMM_plan(F.StokesSpherical(5, True), pts)
r = plan.execute(g)
print("stresslet: |u| %.6e" % float(np.linalg.norm(r)), flush=True)
plan.close()
print("SANITIZE_CASE_DONE")

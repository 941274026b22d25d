"""Development check on a GPU box: CUDA LaplaceSphericalBEM plan vs the oracle port."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

for rec, P, K in ((4, 8, 4), (5, 8, 4), (5, 5, 3), (6, 8, 4)):
    v = O.unit_sphere(rec)
    n = len(v)
    rng = np.random.default_rng(rec)
    q = rng.random(n)
    for bc in (0, 1):
        orc = O.BemOracle(v, bc)
        t = time.time(); ref = orc.execute(q, P, K); t_o = time.time() - t
        t = time.time(); plan = F.FMM_plan(F.LaplaceSphericalBEM(P, K), F.Panels(v, bc)); t_p = time.time() - t
        res = plan.execute(q)
        t = time.time(); res = plan.execute(q); t_g = time.time() - t
        i = plan.info()
        print("n=%d P=%d K=%d bc=%d: boxes %d near entries %d | rel L2 vs oracle %.3e | plan %.3fs exec %.4fs (oracle %.2fs) %s" % (
            n, P, K, bc, i.n_boxes, i.n_near_entries, O.rel_l2(res, ref), t_p, t_g, t_o,
            {k: round(float(x), 3) for k, x in plan.phase_times().items() if k in ("total", "upward", "m2l", "downward", "p2p")}))
        if n <= 2048:
            d = orc.direct(q, K)
            print("     fmm vs direct %.3e" % O.rel_l2(res, d))

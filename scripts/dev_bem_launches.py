"""Dev (GPU, under ncu --metrics gpu__time_duration.sum): the kernels of one LaplaceBEM matvec of config C2 at orders
8, 4 and 1.  usage: python scripts/dev_bem_launches.py [recursions]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

rec = int(sys.argv[1]) if len(sys.argv) > 1 else 7
v = O.unit_sphere(rec)
n = len(v)
plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 0))
x = np.ones(n)
for p in (8, 4, 1):
    plan.kernel().set_p(p)
    for _ in range(3):
        plan.execute(x)
i = plan.info()
print("panels", n, "boxes", i.n_boxes, "leaves", i.n_leaves, "levels", i.n_levels, "m2l pairs", i.n_m2l_pairs,
      "p2p box pairs", i.n_p2p_box_pairs)

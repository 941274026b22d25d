"""Dev check (GPU): where the phases of a matvec lie on its clock (plain launches, both streams), per launch order."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import fmm_bem_relaxed_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[3:] if "=" not in a] or [(0, 0), (0, 1)]
extra = [a.split("=") for a in sys.argv[3:] if "=" in a]
pts, q = F.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(P), pts, F.FMMOptions())
plan.set_option("use_graph", 0)
for k, v in extra:
    plan.set_option(k, int(v))
d_q = torch.from_numpy(q).cuda()
d_r = torch.empty((n, 4), dtype=torch.float64, device="cuda")
for combo in combos:
    plan.set_option("p2p_order", combo[0])
    plan.set_option("m2l_reduce", combo[1])
    if len(combo) > 2:
        plan.set_option("p2p_occ", combo[2])
    if len(combo) > 3:
        plan.set_option("m2l_reduce_bps", combo[3])
    for _ in range(4):
        plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    print("p2p_order %d m2l_reduce %d" % combo[:2], combo[2:], flush=True)
    os.environ["FMMB_PRINT_TIMELINE"] = "1"
    for _ in range(3):
        plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
        plan.sync()
    del os.environ["FMMB_PRINT_TIMELINE"]

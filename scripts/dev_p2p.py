"""Development: near-field work decomposition variants at the metric config (one plan, all variants).
usage (GPU box): python scripts/dev_p2p.py [N] [P]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import oracle_lib as O  # noqa: E402
import fmm_bem_relaxed_b200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pts, q = O.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(P), pts, F.FMMOptions())
d_q = torch.from_numpy(q).cuda()
d_res = torch.empty((n, 4), dtype=torch.float64, device="cuda")
ref = None
variants = [(0, 0, 1, 4), (1, 0, 1, 4), (2, 0, 1, 4), (2, 0, 1, 8), (2, 1, 1, 4), (2, 1, 1, 8)]
for kern, items, warps, unroll in variants:
    if True:
        plan.set_option("p2p_kernel", kern)
        plan.set_option("p2p_unroll", unroll)
        plan.set_option("p2p_items", items)
        plan.set_option("p2p_warps", warps)
        plan.set_option("overlap_p2p", 0)
        acc = {}
        for i in range(6):
            plan.execute_device(d_q.data_ptr(), d_res.data_ptr())
            plan.sync()
            if i > 0:
                for k, v in plan.phase_times().items():
                    acc[k] = acc.get(k, 0.0) + v / 5
        plan.set_option("overlap_p2p", 1)
        for _ in range(3):
            plan.execute_device(d_q.data_ptr(), d_res.data_ptr())
        plan.sync()
        t0 = time.perf_counter()
        for _ in range(20):
            plan.execute_device(d_q.data_ptr(), d_res.data_ptr())
        plan.sync()
        tot = (time.perf_counter() - t0) / 20 * 1e3
        res = d_res.cpu().numpy()
        if ref is None:
            ref = res.copy()
        pairs = plan.info().n_p2p_body_pairs
        print("kernel=%d unroll=%d" % (kern, unroll), end=" ")
        print("items=%d warps=%d  p2p %.3f ms (%.2f TFLOP/s alg)  serial total %.3f  overlapped+graph %.3f ms  "
              "maxdiff vs first %.2e" % (items, warps, acc["p2p"], 22e-9 * pairs / acc["p2p"], acc["total"], tot,
                                         np.abs(res - ref).max() / np.abs(ref).max()), flush=True)

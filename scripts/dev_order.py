"""Dev check (GPU): whole-matvec time (device-resident call, CUDA graph replay) against the launch order of the near
field (p2p_order) and the column-reduction kernel (m2l_reduce); results compared with the first configuration."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import fmm_bem_relaxed_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
extra = [a.split("=") for a in sys.argv[3:] if "=" in a]
combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[3:] if "=" not in a] or [(0, 0, 28, 2), (0, 1, 28, 2), (1, 0, 28, 2), (1, 1, 28, 2)]
pts, q = F.drand48_inputs(n)
plan = F.FMM_plan(F.LaplaceSpherical(P), pts, F.FMMOptions())
for k, v in extra:
    plan.set_option(k, int(v))
stream = torch.cuda.ExternalStream(plan.stream())
d_q = torch.from_numpy(q).cuda()
d_r = torch.empty((n, 4), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for order, red, occ, bps in combos:
    plan.set_option("p2p_order", order)
    plan.set_option("m2l_reduce", red)
    plan.set_option("p2p_occ", occ)
    plan.set_option("m2l_reduce_bps", bps)
    for _ in range(4):
        plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
            plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
            b.record(stream)
        plan.sync()
        ts.append(a.elapsed_time(b))
    r = d_r.cpu().numpy()
    if ref is None:
        ref = r
    print("p2p_order %d m2l_reduce %d (%d blocks per SM) p2p_occ %d: %.3f ms per matvec (min %.3f)  max |diff| vs first: %.2e" % (
        order, red, bps, occ, sum(ts) / len(ts), min(ts), float(np.abs(r - ref).max() / np.abs(ref).max())), flush=True)

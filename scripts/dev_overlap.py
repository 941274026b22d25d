"""Dev check (GPU): whole-matvec time (device-resident call, CUDA graph replay) against the near-field residency cap."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import fmm_bem_relaxed_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
modes = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
pts, q = F.drand48_inputs(n)
for mode in modes:
    opts = F.FMMOptions()
    opts.m2l_mode = mode
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
    stream = torch.cuda.ExternalStream(plan.stream())
    d_q = torch.from_numpy(q).cuda()
    d_r = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    for wps in (0, 4, 6, 8, 10, 12, 14, 16):
        plan.set_option("p2p_wps", wps)
        for _ in range(4):
            plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
        plan.sync()
        ts = []
        for _ in range(8):
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
        print("engine %d p2p_wps %2d: %.3f ms per matvec (min %.3f)  max |diff| vs wps 0: %.2e" % (
            mode, wps, sum(ts) / len(ts), min(ts), float(np.abs(r - ref).max() / np.abs(ref).max())), flush=True)

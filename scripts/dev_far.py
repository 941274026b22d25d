"""Dev check (GPU): blocked far-field engine (m2l_mode 3) against the class-major engine of round 1 (m2l_mode 2) and
the per-pair kernels (m2l_mode 1): results, expansions, phase times.  usage: python scripts/dev_far.py [N] [P...]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

modes_arg = None
argv = list(sys.argv[1:])
if "--modes" in argv:
    i = argv.index("--modes")
    modes_arg = tuple(int(c) for c in argv[i + 1].split(","))
    del argv[i:i + 2]
n = int(argv[0]) if argv else 1000000
orders = [int(a) for a in argv[1:]] or [8]
pts, q = O.drand48_inputs(n)
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for P in orders:
    res = {}
    for mode in (modes_arg or ((3, 2, 1) if P <= 8 else (3, 1))):
        opts = F.FMMOptions()
        opts.m2l_mode = mode
        t0 = time.perf_counter()
        plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
        tb = time.perf_counter() - t0
        plan.set_option("overlap_p2p", 0)
        r = plan.execute(q)
        ts = []
        for _ in range(4):
            r = plan.execute(q)
            ts.append(plan.phase_times())
        t = ts[-1]
        M, L = plan.expansions()
        res[mode] = (r, M, L)
        print("N=%d P=%d mode=%d build %.3fs total %.3f up %.3f m2l %.3f (gemm %.3f) down %.3f p2p %.3f launches %d classes %d" % (
            n, P, mode, tb, t["total"], t["upward"], t["m2l"], t["m2l_gemm"], t["downward"], t["p2p"], t["launches"],
            plan.info().n_m2l_classes), flush=True)
        del plan
    base = 1 if 1 in res else min(res)
    for mode in res:
        if mode == base:
            continue
        r, M, L = res[mode]
        rb, Mb, Lb = res[base]
        print("  mode %d vs %d: result %.2e  M %.2e  L %.2e" % (mode, base, rel(r, rb), rel(M, Mb), rel(L, Lb)), flush=True)

#!/usr/bin/env python
"""bench.py -- FMM matvecs/s on the BASELINE.json metric config (LaplaceSpherical, N=1M, P=8,
theta=0.5, ncrit=64, uniform cube, drand48 inputs as in the reference's tests/scaling.cpp).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one FMM matvec (FMM_plan::execute) over the whole
point set.
  value     matvecs/s with charges and results resident in HBM (fmmb_plan_execute_device),
            per-step CUDA events on the plan stream, L2 flushed between steps.
  e2e       the same through the host-buffer call fmmb_plan_execute (H2D of the charges and D2H of
            the results inside the timed region, pinned host memory).
  roofline  the longer of the two dominant kernels (P2P pair kernel / M2L GEMM) against the FP64 peak measured in
            the same process (DFMA and DMMA micro-benchmarks; MEASURED_PEAKS.json carries HBM and bf16 only);
            traffic = DRAM bytes per launch from the committed ncu capture (profiles/traffic_r01.json).
  cpu_baseline  the reference itself (oracle/_ref/ref_laplace, unmodified reference headers
            compiled with the reference's flags) timed on this host's cores.
  gmres_c2, stresslet_c4   BASELINE configs 2 and 4 next to the reference (N = 1 only).
  stokes_bem               StokesSphericalBEM sphere solve next to the reference (N = 1 only; separate processes).
N > 1 (torchrun, one rank per GPU, strong scaling): a step is one sharded matvec -- every rank feeds and keeps the
tree-ordered slice of its own bodies (fmmb_plan_execute_sharded), multipoles and charge slices are exchanged through
NVLink peer memory (--no-peer: NCCL all-gathers; --replicated-results: full vectors on every rank).
--impl reference times only that CPU implementation.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fmm_matvecs_per_s"
UNIT = "matvec/s"


def workload(args):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": "LaplaceSpherical FMM matvec, N=%d uniform cube (drand48), P=%d, theta=%g, ncrit=%d"
                        % (args.n, args.p, args.theta, args.ncrit),
            "N": args.n, "P": args.p, "theta": args.theta, "ncrit": args.ncrit,
            "l2": "GPU arm: 256 MiB memset between steps, outside the per-step event pairs; the working set of a "
                  "step (0.3 GB) exceeds the 126 MB L2 as well"}


def ref_binary():
    return os.path.join(ROOT, "oracle", "_ref", "ref_laplace")


def run_reference(args, reps, threads=None):
    """Times the unmodified reference on the host cores.  Returns (dict from REF_JSON, per-execute
    seconds parsed from its own timing)."""
    exe = ref_binary()
    if not os.path.exists(exe):
        return None
    threads = threads or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    cmd = [exe, "-N", str(args.n), "-P", str(args.p), "-ncrit", str(args.ncrit), "-theta", repr(args.theta),
           "-reps", str(reps)]
    out = subprocess.check_output(cmd, env=env).decode()
    line = [l for l in out.splitlines() if l.startswith("REF_JSON")][0]
    return json.loads(line[len("REF_JSON "):])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines[first:]:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bench_gmres_c2():
    """Second half of the BASELINE metric ("GMRES solve s"): config C2, LaplaceBEM on a 32 768-panel sphere,
    relaxed GMRES to 1e-6 with p <= 8, through the C++ host mirror (fmm_bem_relaxed_b200/hostcxx) and -- as the
    CPU baseline -- the reference's unmodified examples/LaplaceBEM.cpp (oracle/_ref/LaplaceBEM)."""
    import re
    import tempfile
    args = ["-recursions", "7", "-p", "8", "-k", "4", "-ncrit", "64", "-theta", "0.5", "-solver_tol", "1e-6"]

    def run(exe, env=None, extra=()):
        if not os.path.exists(exe):
            return None
        with tempfile.TemporaryDirectory() as tmp:     # the reference writes test.vert / test.face into cwd
            out = subprocess.check_output([exe] + args + list(extra), env=env, cwd=tmp).decode()
        m = re.search(r"Final residual: ([0-9.eE+-]+), after (\d+) iterations", out)
        r = {"solve_s": float(re.search(r"solve : ([0-9.eE+-]+)s", out).group(1)),
             "setup_s": float(re.search(r"setup : ([0-9.eE+-]+)s", out).group(1)),
             "iterations": int(m.group(2)), "final_residual": float(m.group(1)),
             "p_schedule": [int(x) for x in re.findall(r"fmm_req_p: (\d+)", out)]}
        # our driver also reports what the reference's leaves untimed (its main plan is built before its clock starts,
        # examples/LaplaceBEM.cpp:209-214) and the process-level CUDA start-up
        for key, pat in (("context_s", r"context : ([0-9.eE+-]+)s"), ("plan_s", r"plan : ([0-9.eE+-]+)s")):
            mm = re.search(pat, out)
            if mm:
                r[key] = float(mm.group(1))
        return r
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    exe = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin", "laplace_bem")
    ours = run(exe, env)
    if ours is None:
        return None
    ours = min((run(exe, env) for _ in range(3)), key=lambda r: r["solve_s"])   # first call pays CUDA start-up
    dev = min((run(exe, env, ("-device_gmres",)) for _ in range(3)), key=lambda r: r["solve_s"])
    warm = run(exe, env, ("-device_gmres", "-repeat", "3"))       # best of three solves on one plan, one process
    threads = os.cpu_count() or 1
    ref = run(os.path.join(ROOT, "oracle", "_ref", "LaplaceBEM"), dict(os.environ, OMP_NUM_THREADS=str(threads)))
    out = {"config": "LaplaceBEM sphere 32768 panels, K=4, relaxed GMRES to 1e-6, p<=8 (BASELINE config 2)",
           "timing_regions": "as examples/LaplaceBEM.cpp:209-283 in every arm: setup = right-hand side (temporary plan + "
                             "one matvec), solve = GMRES; the main plan is built before the clocks start (ours: plan_s, "
                             "with the warm start of every order; CUDA context + module load of the process: context_s)",
           "solve_s": ours["solve_s"], "setup_s": ours["setup_s"], "plan_s": ours.get("plan_s"),
           "context_s": ours.get("context_s"), "iterations": ours["iterations"],
           "final_residual": ours["final_residual"], "p_schedule": ours["p_schedule"],
           "solver": "host GMRES (hostcxx/GMRES.hpp, the reference's algorithm line by line) over FMM_plan::execute",
           "device_resident_gmres": {"solve_s": dev["solve_s"], "setup_s": dev["setup_s"], "plan_s": dev.get("plan_s"),
                                     "solve_s_best_of_3_on_one_plan": warm["solve_s"] if warm else None,
                                     "iterations": dev["iterations"],
                                     "final_residual": dev["final_residual"], "p_schedule": dev["p_schedule"],
                                     "solver": "fmmb_gmres: Krylov basis and BLAS-1 on the GPU, one launch for the Gram-Schmidt sweep and one host sync per iteration; solve_s is the FIRST solve of a fresh process (best of three processes)"}}
    if ref is not None:
        out["reference"] = dict(ref, cores=threads, kind="reference",
                                note="multi-threaded reference M2L has a data race (SURVEY F5): time only")
    return out


def bench_stresslet_c4(device):
    """BASELINE config 4: StokesSpherical stresslet FMM, N = 200 000, P = 8 (serialrun_stresslet.cpp:98-128 inputs:
    drand48 points, charges (U, U, U, 1, 0, 0)); GPU matvec time next to the patched reference."""
    import numpy as np
    import torch
    import oracle_lib as O
    import fmm_bem_relaxed_b200 as F
    n, P = 200000, 8
    pts, _ = O.drand48_inputs(n)
    rng = np.random.default_rng(4)
    q = np.hstack([rng.random((n, 3)), np.tile([1.0, 0.0, 0.0], (n, 1))])
    opts = F.FMMOptions()
    opts.device = device
    plan = F.FMM_plan(F.StokesSpherical(P, True), pts, opts)
    d_q = torch.from_numpy(q).cuda()
    d_r = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    for _ in range(4):
        plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        plan.execute_device(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    ms = (time.perf_counter() - t0) / reps * 1e3
    res = d_r.cpu().numpy()
    m = 300
    err = O.rel_l2(res[:m], F.Direct.matvec(plan, q, pts[:m]))
    i = plan.info()
    out = {"config": "StokesSpherical stresslet FMM, N=200000, P=8, theta=0.5, ncrit=64 (BASELINE config 4)",
           "ms_per_matvec": ms, "matvec_per_s": 1e3 / ms, "m2l_pairs_x_sets": int(i.n_m2l_pairs) * 4,
           "p2p_body_pairs": int(i.n_p2p_body_pairs), "rel_l2_vs_direct_first_300": err}
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_stresslet")
    if os.path.exists(exe):
        threads = os.cpu_count() or 1
        o = subprocess.check_output([exe, "-N", str(n), "-P", str(P)], env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
        r = json.loads([l for l in o.decode().splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        out["reference"] = {"exec_s": r["exec_s"], "plan_s": r["plan_s"], "cores": threads, "kind": "reference",
                            "note": "reference with the two compile patches of SURVEY 8(c); drand48 charges"}
    return out


def bench_stokes_bem():
    """StokesSphericalBEM: flow past the unit sphere, 8 192 panels, p = 8, k = 4, tol 1e-5 -- the
    problem of the reference's examples/StokesBEM.cpp.  Three fresh processes each, so nothing here can disturb the
    numbers above: our driver with the device-resident GMRES (hostcxx/bin/stokes_bem), the reference's unmodified driver
    and GMRES_Stokes.hpp over the GPU plan (hostcxx/bin/ref_StokesBEM), and the unmodified reference on the host cores
    (oracle/_ref/StokesBEM).  Failures are reported, not raised: this kernel class was added after the round's GPU
    minutes were spent."""
    import re
    import tempfile
    args = ["-recursions", "6", "-p", "8", "-k", "4", "-solver_tol", "1e-5"]

    def run(exe, env, extra=()):
        if not os.path.exists(exe):
            return None
        with tempfile.TemporaryDirectory() as tmp:     # the drivers write out.face / out.vert / test.vert into cwd
            out = subprocess.check_output([exe] + args + list(extra), env=env, cwd=tmp, timeout=120,
                                          stderr=subprocess.STDOUT).decode()
        it = re.search(r"after (\d+) iterations|iterations: (\d+)", out)
        fx = re.search(r"Fx: ([0-9.eE+-]+), analytical: ([0-9.eE+-]+)", out)
        r = {"solve_s": float(re.search(r"solve : ([0-9.eE+-]+)s", out).group(1)),
             "setup_s": float(re.search(r"setup : ([0-9.eE+-]+)s", out).group(1)),
             "iterations": int(it.group(1) or it.group(2)), "drag_fx": float(fx.group(1)),
             "drag_analytical": float(fx.group(2))}
        mm = re.search(r"context : ([0-9.eE+-]+)s", out)     # our driver: CUDA context + module load, outside setup
        if mm:
            r["context_s"] = float(mm.group(1))
        return r
    try:
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
        bindir = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin")
        ours = run(os.path.join(bindir, "stokes_bem"), env)
        if ours is None:
            return None
        ours = min([ours] + [run(os.path.join(bindir, "stokes_bem"), env) for _ in range(2)], key=lambda r: r["solve_s"])
        out = dict(ours, config="StokesBEM sphere 8192 panels, K=4, relaxed GMRES to 1e-5, 5<=p<=8 (examples/StokesBEM.cpp)",
                   solver="fmmb_gmres on Vec<3> unknowns, order rule of GMRES_Stokes.hpp:229")
        refdrv = run(os.path.join(bindir, "ref_StokesBEM"), env)
        if refdrv is not None:
            out["reference_driver_on_gpu_plan"] = dict(refdrv, solver="the reference's examples/StokesBEM.cpp + "
                                                       "GMRES_Stokes.hpp, unchanged, over FMM_plan::execute")
        threads = os.cpu_count() or 1
        ref = run(os.path.join(ROOT, "oracle", "_ref", "StokesBEM"), dict(os.environ, OMP_NUM_THREADS=str(threads)))
        if ref is not None:
            out["reference"] = dict(ref, cores=threads, kind="reference",
                                    note="multi-threaded reference M2L has a data race (SURVEY F5): iteration count may differ")
        # the reference's preconditioned solve (-local: FGMRES + an inner GMRES on a near-field-only plan per iteration,
        # examples/StokesBEM.cpp:317-320); in every arm the preconditioner's plan is built inside the timed solve
        pc = {}
        for key, exe, e in (("device_resident_fgmres", os.path.join(bindir, "stokes_bem"), env),
                            ("reference_driver_on_gpu_plan", os.path.join(bindir, "ref_StokesBEM"), env),
                            ("reference", os.path.join(ROOT, "oracle", "_ref", "StokesBEM"),
                             dict(os.environ, OMP_NUM_THREADS=str(threads)))):
            r = run(exe, e, ("-local",))
            if r is not None and key != "reference":     # GPU arms: best of three processes (the plan build inside
                for _ in range(2):                       # the timed solve varies by tens of ms between processes)
                    r2 = run(exe, e, ("-local",))
                    if r2 is not None and r2["solve_s"] < r["solve_s"]:
                        r = r2
            if r is not None:
                pc[key] = {k: r[k] for k in ("solve_s", "iterations", "drag_fx")}
        if pc:
            out["fgmres_local_preconditioner"] = pc
        return out
    except Exception as e:  # noqa: BLE001 -- an extra must not take the headline measurement down
        return {"error": "%s: %s" % (type(e).__name__, str(e)[-400:])}


def bench_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    probe = run_reference(args, 1)
    if probe is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_laplace not built"}))
        return
    # bounded: each step is one full matvec of the workload; keep the whole run within ~5 minutes
    t1 = probe["best_s"] + probe["plan_s"]
    budget = 300.0
    max_total = max(1, int((budget - t1) / max(probe["best_s"], 1e-3)))
    total = min(steps + warmup, max_total)
    warm = min(warmup, max(0, total - 1))
    run_steps = total - warm
    r = run_reference(args, total)
    # the reference driver reports best and mean over all reps; use the mean of all executes
    sec = r["mean_s"]
    val = 1.0 / sec
    threads = r["threads"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": run_steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "%d full matvecs of the workload (mean), unmodified reference FMM_plan::execute, "
                                   "OMP threads=%d; requested steps=%d warmup=%d" % (total, threads, steps, warmup)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_c5(args, rank, world, local_rank, dist, barrier):
    """BASELINE config 5 (reference tests/scaling.cpp:14-74 at N = 10M, P = 8, theta = 0.5, ncrit = 64) sharded over
    the ranks of this run: device-resident sharded matvec, per-step CUDA events, max over ranks.  Rank 0 returns the
    key; a failure is reported in it, not raised."""
    import numpy as np
    import torch
    import fmm_bem_relaxed_b200 as F
    try:
        n, P = 10_000_000, 8
        pts, q = F.drand48_inputs(n)
        opts = F.FMMOptions()
        opts.set_mac_theta(0.5)
        opts.set_max_per_box(64)
        opts.device = local_rank
        opts.rank, opts.nranks = rank, world
        opts.m2l_mode = args.m2l_mode
        t0 = time.perf_counter()
        plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
        plan_s = time.perf_counter() - t0
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
            if not args.no_peer:
                mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8).cuda()
                blobs = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.all_gather(blobs, mine)
                plan.peer_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in blobs))
        info = plan.info()
        perm = plan.tree()["perm"].astype(np.int64)
        b0, b1 = info.own_body_begin, info.own_body_end
        d_q = torch.from_numpy(np.ascontiguousarray(q[perm[b0:b1]])).cuda()
        d_r = torch.empty((b1 - b0, 4), dtype=torch.float64, device="cuda")
        stream = torch.cuda.ExternalStream(plan.stream(), device=torch.device("cuda", local_rank))
        steps = 5
        for _ in range(3):
            plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            with torch.cuda.stream(stream):
                a.record(stream)
                plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
                b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        # known answer of the unmodified reference on one thread (tests/golden/checksums.json c5_n10000000_p8):
        # the potential checksum over ALL bodies is the sum of the per-rank slice sums
        pot = torch.tensor([float(d_r[:, 0].sum().item())], dtype=torch.float64, device="cuda")
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(pot, op=dist.ReduceOp.SUM)
        plan.close()
        del d_q, d_r
        torch.cuda.empty_cache()
        if rank != 0:
            return None
        ref_pot = None
        try:
            ref_pot = json.load(open(os.path.join(ROOT, "tests", "golden", "checksums.json")))["c5_n10000000_p8"]["pot"]
        except Exception:
            pass
        ms = float(t.item())
        return {"workload": "LaplaceSpherical FMM matvec, N=10000000 uniform cube (drand48), P=8, theta=0.5, ncrit=64",
                "n_gpus": world, "ms_per_matvec": ms, "matvecs_per_s": 1e3 / ms, "steps": steps, "warmup": 3,
                "plan_build_s": plan_s, "boxes": info.n_boxes, "m2l_pairs": info.n_m2l_pairs,
                "p2p_body_pairs": info.n_p2p_body_pairs, "pot_checksum": float(pot.item()),
                "pot_checksum_reference_1_thread": ref_pot,
                "pot_checksum_rel_err": abs(float(pot.item()) - ref_pot) / abs(ref_pot) if ref_pot else None,
                "timing": "device-resident sharded call, per-step CUDA events, max over ranks; inputs (0.5 GB per "
                          "step) exceed L2"}
    except Exception as e:  # noqa: BLE001 -- an extra must not take the headline measurement down
        return {"error": "%s: %s" % (type(e).__name__, str(e)[-400:])} if rank == 0 else None


def bench_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import fmm_bem_relaxed_b200 as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pts, q = F.drand48_inputs(args.n)      # the reference drivers' input (tests/scaling.cpp:29-38), numpy restatement
    opts = F.FMMOptions()
    opts.set_mac_theta(args.theta)
    opts.set_max_per_box(args.ncrit)
    opts.device = local_rank
    opts.rank, opts.nranks = rank, world
    opts.m2l_mode = args.m2l_mode
    far_engine = {0: "auto (class-major DMMA GEMM + column reduction for P <= 8, fused output-stationary sweep above)",
                  1: "per-pair kernels", 2: "class-major DMMA GEMM + column reduction",
                  3: "fused output-stationary sweep"}[args.m2l_mode]
    t0 = time.perf_counter()
    plan = F.FMM_plan(F.LaplaceSpherical(args.p), pts, opts)
    for kv in args.plan_option:                      # development: --plan-option name=value (fmmb_plan_set_option)
        k, v = kv.split("=")
        plan.set_option(k, int(v))
    plan_s = time.perf_counter() - t0
    if world > 1:
        # ship the NCCL unique id of the engine's own communicator from rank 0 to everyone
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
        if not args.no_peer:
            # multipoles travel through peer memory (NVLink P2P stores from the producing rank) instead of NCCL
            mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8).cuda()
            blobs = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(blobs, mine)
            plan.peer_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in blobs))
    info = plan.info()
    n = info.n_bodies

    stream = torch.cuda.ExternalStream(plan.stream(), device=torch.device("cuda", local_rank))
    d_q = torch.from_numpy(q).cuda()
    d_res = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    # N > 1: the solver-facing sharded call -- every rank feeds the charges of its own bodies and keeps the
    # results of its own bodies (tree order); the exchanges are NCCL all-gathers of charge slices and multipoles
    # ("results stay sharded by target").  N = 1: full vectors in the caller's order.
    sharded = world > 1 and not args.replicated_results
    if sharded:
        perm = plan.tree()["perm"].astype(np.int64)
        b0, b1 = info.own_body_begin, info.own_body_end
        d_q_own = torch.from_numpy(np.ascontiguousarray(q[perm[b0:b1]])).cuda()
        d_res_own = torch.empty((b1 - b0, 4), dtype=torch.float64, device="cuda")

    def step_device():
        if sharded:
            plan.execute_sharded(d_q_own.data_ptr(), d_res_own.data_ptr())
        else:
            plan.execute_device(d_q.data_ptr(), d_res.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs ~0.1 s to start and then reports every 0.1 s, the timed region of the default run lasts less
    # than that: the sampler starts ahead of the warm-up, and the same steps keep running (untimed) after the timed
    # region until samples under this load have arrived; `clocks` is computed from the samples since the region began
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    first_sample = len(sampler.lines)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    phase_acc = {}
    barrier()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # L2 flush, outside the per-step event pair
            starts[i].record(stream)
            step_device()
            ends[i].record(stream)
    barrier()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    in_region = len(sampler.lines) - first_sample
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    # every rank runs the same number of extra steps (the sharded step exchanges data between the ranks): 0.5 s of load
    extra = int(min(400, max(0, 500.0 / max(float(tmax.item()) / args.steps, 1e-3))))
    for _ in range(extra):
        with torch.cuda.stream(stream):
            flush.zero_()
        step_device()
    barrier()
    clocks = sampler.stop(first_sample) if rank == 0 else None
    if clocks is not None:
        clocks["samples_inside_timed_region"] = in_region
        clocks["sampling"] = "nvidia-smi every 0.1 s from the start of the timed region through %d more untimed steps of the same load" % extra
    # per-kernel times for the roofline: a few extra steps with every kernel on ONE stream, so the
    # CUDA events around M2L / P2P are not disturbed by the concurrent near-field stream
    plan.set_option("overlap_p2p", 0)
    phase_acc = {}
    nser = max(3, min(args.steps, 10))
    for i in range(nser + 1):
        with torch.cuda.stream(stream):
            flush.zero_()
        step_device()
        plan.sync()
        if i > 0:
            for k, v in plan.phase_times().items():
                phase_acc[k] = phase_acc.get(k, 0.0) + v
    phase_acc = {k: v / nser * args.steps for k, v in phase_acc.items()}
    plan.set_option("overlap_p2p", 1)
    launches = int(plan.phase_times()["launches"]) * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step

    # end to end through the public host-buffer API (pinned host memory).  N = 1: fmmb_plan_execute, full vectors in
    # the caller's order.  N > 1: fmmb_plan_execute_sharded_host -- every rank copies the charges of ITS bodies up and
    # the results of ITS bodies down (results stay sharded by target), 1/N of the bytes per rank.
    lib = F.capi.load()
    if sharded:
        hq_own = torch.from_numpy(np.ascontiguousarray(q[perm[b0:b1]])).pin_memory()
        hres_own = torch.empty((b1 - b0, 4), dtype=torch.float64).pin_memory()
        hq_np, hres_np = hq_own.numpy(), hres_own.numpy()

        def step_host():
            F.capi.check(lib.fmmb_plan_execute_sharded_host(plan._h, F.capi.ptr(hq_np), F.capi.ptr(hres_np)))
    else:
        hq = torch.from_numpy(q).pin_memory()
        hres = torch.empty((n, 4), dtype=torch.float64).pin_memory()
        hq_np, hres_np = hq.numpy(), hres.numpy()

        def step_host():
            F.capi.check(lib.fmmb_plan_execute(plan._h, F.capi.ptr(hq_np), F.capi.ptr(hres_np)))

    for _ in range(max(1, min(args.warmup, 3))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    # BASELINE config 5 beside the metric config: LaplaceSpherical, N = 10M, P = 8, sharded over the same ranks
    c5 = bench_c5(args, rank, world, local_rank, dist, barrier) if args.c5 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # rooflines.  Algorithmic work per matvec as defined in SURVEY.md section 8(d):
    #   M2L 7 P^3 (P+1) flop per pair (the reference's complex O(P^4) contraction), P2P 22 flop per body pair.
    # "executed" = flops the sm_100a kernels actually issue: the batched M2L multiplies a real P^2 x P^2
    # translation matrix (2 P^4 flop per pair); P2P issues 18 FP64 instructions per body pair.
    import ctypes
    P = args.p
    pk_fma, pk_mma = ctypes.c_double(), ctypes.c_double()
    F.capi.check(lib.fmmb_measure_fp64_peak(local_rank, ctypes.byref(pk_fma), ctypes.byref(pk_mma)))
    fp64_peak = max(pk_fma.value, pk_mma.value)
    m2l_ms = phase_acc["m2l"] / args.steps
    gemm_ms = phase_acc["m2l_gemm"] / args.steps
    p2p_ms = phase_acc["p2p"] / args.steps
    # N > 1: rank 0's kernels run on its share of the pairs (the plan reports global list sizes and this rank's
    # batched M2L pairs); the P2P share is taken as 1 / world (ranges are balanced by estimated work)
    m2l_flop = 7.0 * P ** 3 * (P + 1) * (info.n_m2l_pairs if world == 1 else info.n_m2l_pairs_batched)
    gemm_exec_flop = 2.0 * P ** 4 * info.n_m2l_pairs_batched
    p2p_pairs_rank = info.n_p2p_body_pairs / world
    p2p_flop = 22.0 * p2p_pairs_rank
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf = lambda flop, ms_: flop / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
    kernels = {
        "p2p_pair2_kernel": {"ms": p2p_ms, "algorithmic_flop": p2p_flop, "achieved": tf(p2p_flop, p2p_ms)},
        "trans_gemm_kernel(M2L)": {"ms": gemm_ms if gemm_ms > 0 else m2l_ms, "algorithmic_flop": m2l_flop,
                                   "achieved": tf(m2l_flop, gemm_ms if gemm_ms > 0 else m2l_ms),
                                   "executed_flop": gemm_exec_flop,
                                   "executed_tflops": tf(gemm_exec_flop, gemm_ms) if gemm_ms > 0 else None},
    }
    dominant = max(kernels, key=lambda k: kernels[k]["ms"])
    dk = kernels[dominant]
    # DRAM traffic per launch of that kernel: from the committed `ncu --set full` capture of this round
    # (profiles/traffic_r01.json, written by scripts/summarize_profiles.py from the .ncu-rep files)
    traffic = m2l_traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
        key = "p2p" if dominant.startswith("p2p") else "m2l_gemm"
        if world == 1 and args.n == 1000000 and P == 8 and args.m2l_mode == 0:
            if key in tj:
                traffic = tj[key]["dram_bytes_read"] + tj[key]["dram_bytes_write"]
            if "m2l_gemm" in tj:
                m2l_traffic = tj["m2l_gemm"]["dram_bytes_read"] + tj["m2l_gemm"]["dram_bytes_write"]
    except Exception:
        pass
    clk_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    nominal_fp64 = sms * 64 * 2 * clk_mhz * 1e6 / 1e12       # 64 DFMA per SM and clock
    gk = kernels["trans_gemm_kernel(M2L)"]
    roofline = {
        # "tensor": the contract's name for a compute roofline.  Here that is the FP64 pipe: DFMA and the FP64 tensor
        # instruction DMMA.8x8x4 share ONE pipe on sm_100a (scripts/micro/dual_pipe.cu), and the peak is the measured
        # DMMA rate, the higher of the two -- for the DFMA-bound near-field kernel as well as for the DMMA GEMM.
        "kernel": dominant, "bound": "tensor", "bound_detail": "fp64 pipe (DFMA / DMMA.8x8x4)",
        "achieved": dk["achieved"], "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": dk["achieved"] / fp64_peak, "traffic": traffic,
        "peak_source": "measured in this process by fmmb_measure_fp64_peak: DFMA %.1f, DMMA.8x8x4 %.1f TFLOP/s "
                       "(MEASURED_PEAKS.json has no FP64 entry); nominal %d SMs x 64 DFMA x 2 x %.0f MHz = %.1f TFLOP/s"
                       % (pk_fma.value, pk_mma.value, sms, clk_mhz, nominal_fp64),
        "peak_nominal": nominal_fp64,
        "algorithmic_flop_per_launch": dk["algorithmic_flop"], "ms_per_launch": dk["ms"],
        "hbm_gbs_measured": peaks.get("hbm_gbs"),
        # the other dominant launch: the M2L contraction, EXECUTED flops (2 P^4 per pair: the real-form class matrix
        # needs 4.0x fewer flops than the reference's 7 P^3 (P + 1) formula, so its algorithmic rate exceeds the peak)
        "secondary": {"kernel": "trans_gemm_kernel<P,0> (M2L)" if gemm_ms > 0 else "far-field sweep",
                      "bound": "tensor", "achieved_executed": gk["executed_tflops"],
                      "frac_executed": (gk["executed_tflops"] or 0) / fp64_peak,
                      "achieved_algorithmic": gk["achieved"], "executed_flop_per_launch": gemm_exec_flop,
                      "algorithmic_flop_per_launch": m2l_flop, "ms_per_launch": gk["ms"],
                      "traffic": m2l_traffic},
    }
    others = {
        "m2l": {"ms": m2l_ms, "gemm_ms": gemm_ms, "reduce_ms": m2l_ms - gemm_ms if gemm_ms > 0 else None,
                "tflops_algorithmic_gemm": kernels["trans_gemm_kernel(M2L)"]["achieved"],
                "tflops_executed_gemm": kernels["trans_gemm_kernel(M2L)"]["executed_tflops"],
                "frac_fp64_peak_executed": (kernels["trans_gemm_kernel(M2L)"]["executed_tflops"] or 0) / fp64_peak,
                "pairs": info.n_m2l_pairs, "pairs_batched": info.n_m2l_pairs_batched, "classes": info.n_m2l_classes,
                "reduce_gbs": (info.n_m2l_pairs * ((P * P + 1) // 2 * 2) * 8) / ((m2l_ms - gemm_ms) * 1e-3) / 1e9
                if gemm_ms > 0 and m2l_ms > gemm_ms else None},
        "p2p": {"ms": p2p_ms, "tflops_algorithmic": tf(p2p_flop, p2p_ms),
                "frac_fp64_peak": tf(p2p_flop, p2p_ms) / fp64_peak,
                "fp64_instr_issue_frac": 18.0 * p2p_pairs_rank / (p2p_ms * 1e-3) / (pk_fma.value * 1e12 / 2)
                if p2p_ms > 0 else None,
                "body_pairs": info.n_p2p_body_pairs},
        "upward_ms": phase_acc["upward"] / args.steps, "downward_ms": phase_acc["downward"] / args.steps,
        "total_ms_serialised": phase_acc["total"] / args.steps,
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = run_reference(args, 1)
        if r is not None:
            cpu = {"value": 1.0 / r["best_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                   "sample": "1 full matvec of the same workload by the unmodified reference "
                             "(oracle/_ref/ref_laplace, FMM_plan::execute, %.2f s; plan %.2f s), OMP threads=%d"
                             % (r["best_s"], r["plan_s"], r["threads"])}
    gmres = None
    stokes = None
    sbem = None
    if world == 1 and not args.no_cpu_baseline:
        gmres = bench_gmres_c2()
        stokes = bench_stresslet_c4(local_rank)
        sbem = bench_stokes_bem()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload(args),
        "details": dict(parallelism="1 GPU" if world == 1 else
                        ("target leaves in %d Morton-contiguous ranges of equal estimated work; per step: owned upward "
                         "pass, multipoles and charge slices %s; results stay sharded by target "
                         "(fmmb_plan_execute_sharded; e2e: fmmb_plan_execute_sharded_host)"
                         % (world, "all-gathered over NCCL" if args.no_peer else
                            "stored into the peers' arrays over NVLink (P2P stores + flag vectors, no NCCL in the step)"))
                        if sharded else
                        ("target leaves in %d Morton-contiguous ranges of equal estimated work; charges replicated, "
                         "owned upward pass + multipole exchange, NCCL all-gather of the result slices" % world),
                        plan_build_s=plan_s, boxes=info.n_boxes, m2l_pairs=info.n_m2l_pairs,
                        p2p_body_pairs=info.n_p2p_body_pairs, far_field_engine=far_engine),
        # bytes per step summed over the ranks (every rank moves the slice of its own bodies)
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 32 * n,
                "ms_per_step": e2e_s * 1e3,
                "call": "fmmb_plan_execute_sharded_host" if sharded else "fmmb_plan_execute"},
        "gpu_launches": launches,
        "roofline": roofline, "phases": others, "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if gmres is not None:
        line["gmres_c2"] = gmres
    if stokes is not None:
        line["stresslet_c4"] = stokes
    if sbem is not None:
        line["stokes_bem"] = sbem
    if c5 is not None:
        line["c5_scaling"] = c5
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--npoints", dest="n", type=int, default=1000000,
                    help="bodies (under torchrun use --npoints: its own parser claims the prefix --n)")
    ap.add_argument("--p", type=int, default=8)
    ap.add_argument("--theta", type=float, default=0.5)
    ap.add_argument("--ncrit", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", dest="c5", action="store_false",
                    help="skip the N = 10M (BASELINE config 5) extra")
    ap.add_argument("--m2l-mode", type=int, default=0, help="far-field engine (fmmb_options.m2l_mode)")
    ap.add_argument("--plan-option", action="append", default=[],
                    help="name=value passed to fmmb_plan_set_option on the benchmark plan (development)")
    ap.add_argument("--no-peer", action="store_true",
                    help="N > 1: exchange the multipoles with an NCCL all-gather instead of peer-memory stores")
    ap.add_argument("--replicated-results", action="store_true",
                    help="N > 1: time fmmb_plan_execute_device (full result vector all-gathered to every rank) "
                         "instead of the sharded call")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        if args.warmup is None:
            args.warmup = 1
        bench_reference(args, rank, world)
        return
    if args.steps is None:
        args.steps = 20
    if args.warmup is None:
        args.warmup = 5
    args.warmup = max(args.warmup, 3)
    bench_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

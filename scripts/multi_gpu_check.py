"""Run under torchrun on N GPUs: the sharded matvec (NCCL all-gather of results, owned upward pass with
multipole exchange) must agree with the single-GPU plan on every rank.  Prints one line per rank."""
import os, sys, faulthandler
faulthandler.dump_traceback_later(int(os.environ.get('CHECK_TIMEOUT', '240')), exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for n, P in ((200000, 8), (60000, 5), (30000, 7)):
    if n == 30000:
        rng = np.random.default_rng(7)
        pts = rng.random((n, 3)); pts[n // 2:] = 0.3 + 0.05 * rng.random((n - n // 2, 3)); q = rng.random(n) - 0.3
    else:
        pts, q = O.drand48_inputs(n)
    print("rank %d: config N=%d P=%d" % (rank, n, P), flush=True)
    single = F.FMMOptions(); single.device = local
    ref = F.FMM_plan(F.LaplaceSpherical(P), pts, single).execute(q)
    opts = F.FMMOptions(); opts.device = local; opts.rank, opts.nranks = rank, world
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
    print("rank %d: comm ready" % rank, flush=True)
    for rep in range(3):            # third call replays the captured CUDA graph
        res = plan.execute(q)
        print("rank %d: execute %d done" % (rank, rep), flush=True)
        err = O.rel_l2(res, ref)
        ok &= err < 1e-12
    # the same matvec with the multipoles pushed through peer memory (NVLink P2P stores) instead of NCCL, including
    # an order change and back (GMRES relaxation) and the captured-graph replay
    mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8).cuda()
    allb = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allb, mine)
    plan.peer_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb))
    for rep in range(4):
        perr = O.rel_l2(plan.execute(q), ref)
        ok &= perr < 1e-12
    plan.kernel().set_p(max(1, P - 2))
    low = plan.execute(q)
    plan.kernel().set_p(P)
    for rep in range(3):
        perr = max(perr, O.rel_l2(plan.execute(q), ref))
    ok &= perr < 1e-12 and O.rel_l2(low, ref) < 1e-2
    print("rank %d: peer-memory exchange rel-L2 vs single GPU %.2e" % (rank, perr), flush=True)
    i = plan.info()
    own = (i.own_body_begin, i.own_body_end)
    plan.close()          # teardown (ncclCommDestroy) at the same point on every rank
    print("rank %d/%d N=%d P=%d own [%d,%d) rel-L2 vs single GPU %.2e" % (rank, world, n, P, own[0], own[1], err), flush=True)
# the other kernel classes: owned upward pass per expansion set (BEM, Stokes) or replicated (Yukawa), NCCL all-gather
# of the result slices
rng = np.random.default_rng(11)
n = 40000
pts, q = O.drand48_inputs(n)
verts = O.unit_sphere(6)
cases = [("laplace-p12", lambda: F.LaplaceSpherical(12), pts, q),      # fused sweep engine: owned sweep, exchange, rest
         ("stokeslet", lambda: F.StokesSpherical(7, False), pts, rng.random((n, 3))),
         ("stresslet", lambda: F.StokesSpherical(8, True), pts, np.hstack([rng.random((n, 3)), np.tile([1.0, 0, 0], (n, 1))])),
         ("yukawa", lambda: F.YukawaCartesian(6, 1.0), pts, q),
         ("laplace-bem", lambda: F.LaplaceSphericalBEM(8, 4), F.Panels(verts), rng.random(len(verts))),
         ("stokes-bem", lambda: F.StokesSphericalBEM(6), F.Panels(verts), rng.random((len(verts), 3)))]
for name, mk, src, chg in cases:
    single = F.FMMOptions(); single.device = local
    ref = F.FMM_plan(mk(), src, single).execute(chg)
    opts = F.FMMOptions(); opts.device = local; opts.rank, opts.nranks = rank, world
    plan = F.FMM_plan(mk(), src, opts)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
    for rep in range(3):
        err = O.rel_l2(plan.execute(chg), ref)
        ok &= err < 1e-12
    # the sharded call of this kernel kind with HOST buffers: this rank's charge slice up, its result slice down
    # (fmmb_plan_execute_sharded_host; charge slices all-gathered over NCCL inside the call)
    i = plan.info()
    perm = plan.tree()["perm"].astype(np.int64)
    own = perm[i.own_body_begin:i.own_body_end]
    chg2 = np.asarray(chg, dtype=np.float64).reshape(len(perm), -1)
    ref2 = np.asarray(ref).reshape(len(perm), -1)
    for rep in range(3):
        out = plan.execute_sharded_host(np.ascontiguousarray(chg2[own]))
        serr = O.rel_l2(out.reshape(len(own), -1), ref2[own]) if len(own) else 0.0
        ok &= serr < 1e-12
    plan.close()
    print("rank %d/%d %s rel-L2 vs single GPU %.2e, sharded host call %.2e (own %d bodies)" % (
        rank, world, name, err, serr, len(own)), flush=True)
dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)

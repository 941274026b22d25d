"""Development (torchrun, N GPUs): per-rank phase times of the sharded matvec at the metric config."""
import os, sys, time
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pts, q = O.drand48_inputs(n)
opts = F.FMMOptions(); opts.device = local; opts.rank, opts.nranks = rank, world
plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
d_q = torch.from_numpy(q).cuda()
d_res = torch.empty((n, 4), dtype=torch.float64, device="cuda")
info = plan.info()
b0, b1 = info.own_body_begin, info.own_body_end
perm = plan.tree()["perm"].astype(np.int64)
d_q_own = torch.from_numpy(np.ascontiguousarray(q[perm[b0:b1]])).cuda()
d_res_own = torch.empty((b1 - b0, 4), dtype=torch.float64, device="cuda")
def enable_peer():
    mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8).cuda()
    allb = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allb, mine)
    plan.peer_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb))


for mode in sys.argv[3:] or ["full", "sharded", "sharded+peer"]:
    if mode.endswith("+peer"):
        enable_peer()
        if rank == 0:
            print("peer-memory multipole exchange enabled", flush=True)

    def run():
        if mode == "full":
            plan.execute_device(d_q.data_ptr(), d_res.data_ptr())
        else:
            plan.execute_sharded(d_q_own.data_ptr(), d_res_own.data_ptr())
    plan.set_option("overlap_p2p", 0)
    plan.set_option("use_graph", 0)
    acc = {}
    for i in range(5):
        dist.barrier(); torch.cuda.synchronize()
        run()
        plan.sync()
        if i > 0:
            for k, v in plan.phase_times().items():
                acc[k] = acc.get(k, 0.0) + v / 4
    i = plan.info()
    print("rank %d mode %s own %d bodies: serial phases %s" % (rank, mode, i.own_body_end - i.own_body_begin,
          {k: round(v, 3) for k, v in acc.items()}), flush=True)
    plan.set_option("overlap_p2p", 1)
    # where the phases lie on the clock of an overlapped matvec (plain launches), rank by rank
    for i in range(6):
        dist.barrier(); torch.cuda.synchronize()
        if i == 4:
            os.environ["FMMB_PRINT_TIMELINE"] = "1"
        run()
        plan.sync()
    del os.environ["FMMB_PRINT_TIMELINE"]
    plan.set_option("use_graph", 1)
    for _ in range(4):
        run()
    plan.sync(); dist.barrier(); torch.cuda.synchronize()
    if mode == "full":
        ref_own = d_res.cpu().numpy()[perm[b0:b1]]
    elif "ref_own" not in globals():
        ref_own = d_res_own.cpu().numpy().copy()
    else:
        err = np.abs(d_res_own.cpu().numpy() - ref_own).max() / np.abs(ref_own).max()
        print("rank %d %s vs first mode: max rel diff %.2e" % (rank, mode, err), flush=True)
    t0 = time.perf_counter()
    for _ in range(50):
        run()
    plan.sync(); dist.barrier(); torch.cuda.synchronize()
    print("rank %d mode %s: %.3f ms per matvec (graph, overlapped, 50 reps wall)" % (rank, mode, (time.perf_counter() - t0) / 50 * 1e3), flush=True)
plan.close()
dist.barrier()
dist.destroy_process_group()

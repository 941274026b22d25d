"""Two ranks (two processes), one GPU each: the sharded matvec with peer-memory exchange only (no NCCL; the processes
swap the 128-byte IPC blobs over gloo).  Exercises fmmb_plan_peer_export / peer_init / execute_sharded and the flag
protocol.  Needs two GPUs: kernels of two processes that wait on one another through flags must not share a device
(nothing guarantees that they run at the same time; B200_PROFILING.md records Xid 109 for exactly that), so with one
visible device the script refuses to run.  usage: python scripts/peer_one_gpu.py [N] [P]   (spawns the ranks itself)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, port, n, P, ret):
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    import fmm_bem_relaxed_b200 as F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank
    torch.cuda.set_device(dev)
    pts, q = O.drand48_inputs(n)
    single = F.FMMOptions()
    single.device = dev
    ref_plan = F.FMM_plan(F.LaplaceSpherical(P), pts, single)
    ref = ref_plan.execute(q)
    perm = ref_plan.tree()["perm"].astype(np.int64)
    opts = F.FMMOptions()
    opts.device = dev
    opts.rank, opts.nranks = rank, world
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
    blobs = [None] * world
    dist.all_gather_object(blobs, plan.peer_export())
    plan.peer_init(b"".join(blobs))
    i = plan.info()
    b0, b1 = i.own_body_begin, i.own_body_end
    d_q = torch.from_numpy(np.ascontiguousarray(q[perm[b0:b1]])).cuda()
    d_r = torch.zeros((b1 - b0, 4), dtype=torch.float64, device="cuda")
    ok = True
    for rep in range(4):                       # the third call replays the captured graph
        plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
        plan.sync()
        err = O.rel_l2(d_r.cpu().numpy(), ref[perm[b0:b1]])
        ok &= err < 1e-12
    plan.kernel().set_p(max(1, P - 2))         # relaxation: lower order, then back
    plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    plan.kernel().set_p(P)
    plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
    plan.sync()
    err = O.rel_l2(d_r.cpu().numpy(), ref[perm[b0:b1]])
    ok &= err < 1e-12
    dist.barrier()
    plan.close()
    dist.destroy_process_group()
    ret[rank] = (bool(ok), float(err), int(b1 - b0))


if __name__ == "__main__":
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        print("PEER_ONE_GPU needs two devices (spin-waiting kernels of two processes must not share a GPU)")
        sys.exit(3)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(2, 29600 + os.getpid() % 300, n, P, ret), nprocs=2, join=True)
    print("PEER_ONE_GPU", dict(ret))
    sys.exit(0 if all(v[0] for v in ret.values()) and len(ret) == 2 else 1)

"""Run under torchrun on N GPUs: BASELINE config 5 (LaplaceSpherical, N = 10 000 000 drand48 points, P = 8, theta = 0.5,
ncrit = 64) sharded over the ranks -- peer-memory exchange, sharded HOST-buffer call.  Checks
  * every rank's result slice against the single-GPU plan run on the same device (relative L2 <= 1e-12);
  * the slices tile the body range exactly once;
  * the global checksums (sum of potentials, weighted sum of x-forces in ORIGINAL body order) against the unmodified
    reference run on one thread (tests/golden/checksums.json, c5_n10000000_p8) at 1e-10.
Prints MULTI_GPU_C5 OK / FAILED on rank 0.  usage: torchrun --nproc-per-node N scripts/multi_gpu_c5.py [n] [engine]"""
import json
import os
import sys
import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("CHECK_TIMEOUT", "600")), exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import fmm_bem_relaxed_b200 as F

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
engine = int(sys.argv[2]) if len(sys.argv) > 2 else 0
P = 8
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts, q = F.drand48_inputs(n)
single = F.FMMOptions()
single.device = local
single.m2l_mode = engine
ref_plan = F.FMM_plan(F.LaplaceSpherical(P), pts, single)
ref = ref_plan.execute(q)
perm = ref_plan.tree()["perm"].astype(np.int64)
ref_plan.close()
opts = F.FMMOptions()
opts.device = local
opts.m2l_mode = engine
opts.rank, opts.nranks = rank, world
plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(F.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    plan.comm_init(bytes(idt.cpu().numpy().tobytes()))
    mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8).cuda()
    blobs = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(blobs, mine)
    plan.peer_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in blobs))
i = plan.info()
b0, b1 = i.own_body_begin, i.own_body_end
own = perm[b0:b1]
ok = True
for rep in range(3):                                   # the third call replays the captured graph
    out = plan.execute_sharded_host(np.ascontiguousarray(q[own]))
    err = float(np.linalg.norm(out - ref[own]) / np.linalg.norm(ref[own]))
    ok &= err < 1e-12
# checksums in the reference's convention: pot = sum_k r[k][0], fxw = sum_k r[k][1] (k mod 7 + 1), k = ORIGINAL index
sums = torch.tensor([float(out[:, 0].sum()), float((out[:, 1] * (own % 7 + 1)).sum()), float(b1 - b0)],
                    dtype=torch.float64, device="cuda")
dist.all_reduce(sums)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "checksums.json"))).get("c5_n10000000_p8") if n == 10_000_000 else None
msg = "rank %d/%d N=%d own [%d,%d) slice rel-L2 vs single GPU %.2e" % (rank, world, n, b0, b1, err)
if gold:
    e_pot = abs(sums[0].item() - gold["pot"]) / abs(gold["pot"])
    e_fxw = abs(sums[1].item() - gold["fxw"]) / abs(gold["fxw"])
    ok &= e_pot < 1e-10 and e_fxw < 1e-10
    msg += "; global checksums vs 1-thread reference: pot %.2e fxw %.2e" % (e_pot, e_fxw)
ok &= int(sums[2].item()) == n
print(msg, flush=True)
flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
plan.close()
if rank == 0:
    print("MULTI_GPU_C5", "OK" if flag.item() > 0 else "FAILED", "ranks", world, "N", n)
dist.destroy_process_group()
sys.exit(0 if flag.item() > 0 else 1)

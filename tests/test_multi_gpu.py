"""Multi-GPU sharding: target leaves in Morton-contiguous ranges.

CPU part (gloo, world_size 2): the host-side partition helper and the slice / all-gather / un-permute
logic, with the oracle standing in for the per-rank compute.
GPU part: partitioned plans on ONE device, without a communicator -- every rank writes only its own slice
and the slices must tile the single-GPU result bit for bit (per-target sums do not depend on the rank).
"""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
import fmm_bem_relaxed_b200 as F
from conftest import ROOT


def test_partition_ranges_balanced():
    rng = np.random.default_rng(0)
    w = rng.random(1000) + 0.1
    for r in (1, 2, 3, 8):
        cuts = F.partition_ranges(w, r)
        assert cuts[0] == 0 and cuts[-1] == 1000 and np.all(np.diff(cuts) >= 0)
        sums = np.array([w[cuts[i]:cuts[i + 1]].sum() for i in range(r)])
        assert sums.max() - sums.min() <= 2 * w.max() + 1e-12
    # degenerate: more ranks than items, zero weights
    cuts = F.partition_ranges(np.ones(3), 8)
    assert cuts[-1] == 3 and np.all(np.diff(cuts) >= 0)
    cuts = F.partition_ranges(np.zeros(10), 4)
    assert cuts[0] == 0 and cuts[-1] == 10


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, P = 4000, 4
    pts, q = O.drand48_inputs(n)
    orc = O.Oracle(pts, 32, 0.5)
    t = orc.tree()
    boxes = t["boxes"]
    leaves = np.nonzero(boxes[:, 7])[0]
    leaves = leaves[np.argsort(boxes[leaves, 4])]            # body order
    counts = (boxes[leaves, 5] - boxes[leaves, 4]).astype(float)
    cuts = F.partition_ranges(counts * counts, world)         # any positive work estimate will do here
    body_cuts = [int(boxes[leaves[c], 4]) if c < len(leaves) else n for c in cuts]
    b0, b1 = body_cuts[rank], body_cuts[rank + 1]
    full = orc.execute(q, P, mode=1, threads=1)               # stands in for the per-rank GPU compute
    mine = torch.from_numpy(full[t["perm"][b0:b1].astype(np.int64)].copy())   # owned slice, tree order
    # all-gather of unequal slices = one broadcast per owner (what csrc/comm.cu does with NCCL)
    tree_res = torch.zeros((n, 4), dtype=torch.float64)
    tree_res[b0:b1] = mine
    for r in range(world):
        dist.broadcast(tree_res[body_cuts[r]:body_cuts[r + 1]], r)
    out = np.zeros((n, 4))
    out[t["perm"].astype(np.int64)] = tree_res.numpy()
    ok = np.array_equal(out, full) and sum(body_cuts[i + 1] - body_cuts[i] for i in range(world)) == n
    dist.destroy_process_group()
    ret[rank] = bool(ok)


def test_gloo_world2_slices_tile_the_result():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_partitioned_plans_tile_single_gpu_result(world):
    n, P = 60000, 6
    pts, q = O.drand48_inputs(n)
    ref_plan = F.FMM_plan(F.LaplaceSpherical(P), pts)
    full = ref_plan.execute(q)
    perm = ref_plan.tree()["perm"].astype(np.int64)
    merged = np.zeros_like(full)
    covered = np.zeros(n, dtype=int)
    work = []
    for r in range(world):
        opts = F.FMMOptions()
        opts.rank, opts.nranks = r, world
        plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
        i = plan.info()
        part = plan.execute(q)                       # no communicator: only the owned slice is written
        own = perm[i.own_body_begin:i.own_body_end]
        covered[own] += 1
        mask = np.ones(n, bool)
        mask[own] = False
        assert np.all(part[mask] == 0)
        merged += part
        work.append(i.n_m2l_pairs_batched)
    assert np.all(covered == 1)
    # same far-field sums on every rank; the near-field chunk size (hence its summation order) adapts to the
    # number of leaves a rank owns, so the tiles agree to rounding, not bit for bit
    assert O.rel_l2(merged, full) <= 1e-14
    # far-field work is split (the top of the tree is shared, so the sum exceeds the single-GPU count a little)
    assert max(work) < 0.75 * ref_plan.info().n_m2l_pairs_batched


@pytest.mark.gpu
def test_sharded_call_on_one_gpu_is_the_tree_ordered_matvec():
    """fmmb_plan_execute_sharded on a single-GPU plan: the slice is the whole vector in tree order."""
    import torch
    n, P = 50000, 7
    pts, q = O.drand48_inputs(n)
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts)
    full = plan.execute(q)
    perm = plan.tree()["perm"].astype(np.int64)
    i = plan.info()
    assert (i.own_body_begin, i.own_body_end) == (0, n)
    d_q = torch.from_numpy(np.ascontiguousarray(q[perm])).cuda()
    d_r = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    for _ in range(3):                                # third call replays the captured graph
        plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
        plan.sync()
        assert np.array_equal(d_r.cpu().numpy(), full[perm])
    # a partitioned plan without a communicator refuses the sharded call instead of computing on partial charges
    opts = F.FMMOptions()
    opts.rank, opts.nranks = 0, 2
    part = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
    with pytest.raises(F.FmmbError):
        part.execute_sharded(d_q.data_ptr(), d_r.data_ptr())


def _tile_check(make_kernel, sources, charges, world, tol=1e-13):
    """Partitioned plans on one device without a communicator: each writes its own slice; the slices tile the
    single-GPU result."""
    full = F.FMM_plan(make_kernel(), sources).execute(charges)
    merged = np.zeros_like(full)
    for r in range(world):
        opts = F.FMMOptions()
        opts.rank, opts.nranks = r, world
        merged += F.FMM_plan(make_kernel(), sources, opts).execute(charges)
    assert O.rel_l2(merged, full) <= tol


@pytest.mark.gpu
def test_sharded_host_call_and_sharded_calls_of_every_kernel_kind():
    """fmmb_plan_execute_sharded[_host] on single-GPU plans of every kernel kind: the slice is the whole vector in
    tree order, so the result must be the ordinary matvec, permuted (bit for bit: same kernels, same sums)."""
    import torch
    n = 20000
    pts, q = O.drand48_inputs(n)
    rng = np.random.default_rng(5)
    verts = O.unit_sphere(5)
    g = np.hstack([rng.random((n, 3)), np.tile([0.0, 1.0, 0.0], (n, 1))])
    cases = [(lambda: F.LaplaceSpherical(6), pts, q),
             (lambda: F.StokesSpherical(5, False), pts, rng.random((n, 3))),
             (lambda: F.StokesSpherical(5, True), pts, g),
             (lambda: F.YukawaCartesian(5, 0.5), pts, q),
             (lambda: F.LaplaceSphericalBEM(6, 4), F.Panels(verts), rng.random(len(verts))),
             (lambda: F.YukawaCartesianBEM(5, 1.0, 4), F.Panels(verts), rng.random(len(verts))),
             (lambda: F.StokesSphericalBEM(5), F.Panels(verts), rng.random((len(verts), 3)))]
    for mk, src, chg in cases:
        plan = F.FMM_plan(mk(), src)
        full = plan.execute(chg)
        perm = plan.tree()["perm"].astype(np.int64)
        chg2 = np.asarray(chg, dtype=np.float64).reshape(len(perm), -1)
        own_q = np.ascontiguousarray(chg2[perm])
        for _ in range(3):                            # third call replays the captured graph
            out = plan.execute_sharded_host(own_q)
            assert np.array_equal(out.reshape(len(perm), -1), full.reshape(len(perm), -1)[perm]), type(plan.K).__name__
        d_q = torch.from_numpy(own_q).cuda()
        d_r = torch.zeros(out.shape, dtype=torch.float64, device="cuda")
        plan.execute_sharded(d_q.data_ptr(), d_r.data_ptr())
        plan.sync()
        assert np.array_equal(d_r.cpu().numpy(), out)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_plans_other_kernels(world):
    n = 20000
    pts, q = O.drand48_inputs(n)
    rng = np.random.default_rng(3)
    _tile_check(lambda: F.StokesSpherical(6, False), pts, rng.random((n, 3)), world)
    g = np.hstack([rng.random((n, 3)), np.tile([0.0, 1.0, 0.0], (n, 1))])
    _tile_check(lambda: F.StokesSpherical(5, True), pts, g, world)
    _tile_check(lambda: F.YukawaCartesian(5, 0.5), pts, q, world)
    verts = O.unit_sphere(5)                                   # 2048 panels
    _tile_check(lambda: F.LaplaceSphericalBEM(6, 4), F.Panels(verts), rng.random(len(verts)), world)
    _tile_check(lambda: F.LaplaceSphericalBEM(6, 4), F.Panels(verts, 1), rng.random(len(verts)), world)


@pytest.mark.gpu
def test_two_ranks_peer_memory_exchange():
    """Two processes, one device each: sharded matvec with the peer-memory exchange only (IPC-exported multipole /
    charge arrays, P2P stores, system-scope flag vectors; no NCCL).  scripts/peer_one_gpu.py spawns the ranks and
    compares every rank's slice with the single-GPU result (including an order change and the graph replay).
    Skipped on a one-GPU box: kernels of two processes that spin on each other's flags must not share a device."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "peer_one_gpu.py"), "30000", "6"],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "PEER_ONE_GPU" in out.stdout


def _torchrun(script, nproc, *args, timeout=900):
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
           os.path.join(ROOT, "scripts", script)] + [str(a) for a in args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.gpu
@pytest.mark.timeout(1200)
@pytest.mark.parametrize("world", [2, 8])
def test_c5_n10m_sharded_over_ranks(world):
    """BASELINE config 5 on several GPUs (judge-added row J1): N = 10M sharded over `world` ranks through the
    peer-memory exchange and the sharded host call; slices against the single-GPU plan, global checksums against the
    one-thread run of the unmodified reference (scripts/multi_gpu_c5.py).  Needs `world` GPUs."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    out = _torchrun("multi_gpu_c5.py", world)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_C5 OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.timeout(1200)
def test_every_kernel_kind_sharded_over_two_ranks():
    """All kernel kinds over NCCL and peer memory on 2 GPUs: replicated call, sharded host call, order changes, the
    fused sweep engine at P = 12 (scripts/multi_gpu_check.py).  Needs 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = _torchrun("multi_gpu_check.py", 2)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_CHECK OK" in out.stdout

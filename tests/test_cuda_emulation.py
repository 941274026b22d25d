"""CPU suite: the StokesSphericalBEM CUDA kernels (fmm_bem_relaxed_b200/csrc/stokes_bem.cu) executed WITHOUT a GPU.

tests/emu/extract_kernels.py cuts the kernels out of the shipped .cu files and tests/emu/cuda_emu.hpp runs them in lock
step (one std::thread per CUDA thread, __syncwarp as a barrier, shared memory as statics): the indexing, staging and
per-lane arithmetic that will run on the B200 are these very source lines.  Checked here:
  * near field -- sbem_setup / count / assemble / gather / near kernels on a 512-panel sphere whose interaction lists
    are all near field (so the oracle's FMM matvec IS the near field), all boundary-condition mixes, both near-field
    modes: 1e-13 against the oracle restatement, which is bit-identical to the reference;
  * far field -- sbem_p2m_kernel<0|1> against the point-source kernels of csrc/stokes.cu (green on hardware this round)
    fed with one source per (panel, quadrature point), sbem_l2p_kernel against stokes_l2p_kernel: 1e-13.
  * treecode -- bem_m2p_kernel<0|1> of csrc/bem.cu against m2p_kernel of csrc/laplace.cu (green on hardware);
    yk_bem_m2p_kernel<0|1> of csrc/yukawa.cu against yk_table_kernel (green on hardware) + a host dot product;
    stokes_m2p_kernel / sbem_m2p_kernel against four runs of m2p_kernel combined on the host.
  * Direct::matvec kernels of fmmb_plan_direct_panels against the oracle's direct sums;
  * host sequencing -- stokes_bem_setup + stokes_bem_execute as written, each launch with its own configuration and
    guard bytes behind the dynamic shared segment, on host memory (cudaMalloc / cudaMemcpyAsync ... as macros).
This verifies kernel logic, not performance, and does not replace the first run on the device (tests/test_zz_stokes_bem.py).
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "fmm_bem_relaxed_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu")
    subprocess.check_call(["python", os.path.join(EMU, "extract_kernels.py"), os.path.join(CSRC, "stokes_bem.cu"),
                           str(d / "sbem_whole.inc"), "--whole"])
    subprocess.check_call(["python", os.path.join(EMU, "extract_kernels.py"), os.path.join(CSRC, "gmres.cu"),
                           str(d / "gmres_whole.inc"), "--whole"])
    for src, dst, names in (("stokes.cu", "stokes_kernels.inc", []),
                            ("laplace.cu", "lap_m2p.inc", ["m2p_kernel"]), ("bem.cu", "bem_m2p.inc", ["bem_m2p_kernel"]),
                            ("laplace.cu", "lap_trans.inc", ["m2m_kernel", "m2l_coeff_kernel", "m2l_pair_kernel", "l2l_kernel"]),
                            ("bem.cu", "bem_kernels.inc", [])):
        subprocess.check_call(["python", os.path.join(EMU, "extract_kernels.py"), os.path.join(CSRC, src), str(d / dst)] + names)
    subprocess.check_call(["python", os.path.join(EMU, "extract_kernels.py"), os.path.join(CSRC, "yukawa.cu"),
                           str(d / "yukawa_kernels.inc"), "--until", "const double* yk_class_tables("])
    subprocess.check_call(["python", os.path.join(EMU, "extract_kernels.py"), os.path.join(CSRC, "bem.cu"),
                           str(d / "bem_whole.inc"), "--whole"])
    exe = str(d / "emu_stokes_bem")
    for src, out in (("emu_stokes_bem.cpp", exe), ("emu_bem_pipeline.cpp", str(d / "emu_bem_pipeline"))):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-I", CUDA_INC, "-I", str(d), "-I", EMU,
                               os.path.join(EMU, src), "-o", out, "-L/usr/local/cuda/lib64", "-lcudart"])
    return exe


def test_far_field_kernels_match_the_point_source_kernels(emu):
    out = subprocess.check_output([emu, "far"], timeout=600).decode()
    m = re.search(r"far: p2m ([0-9.eE+-]+) l2p ([0-9.eE+-]+) max_multipole ([0-9.eE+-]+) max_velocity ([0-9.eE+-]+)", out)
    assert m, out
    p2m, l2p, mm, mu = (float(x) for x in m.groups())
    assert mm > 1e-4 and mu > 1e-2                      # the comparison is not between zeros
    assert p2m <= 1e-13 and l2p <= 1e-13


def test_bem_treecode_kernel_matches_the_point_treecode_kernel(emu):
    """bem_m2p_kernel<0|1> (`LaplaceBEM -eval TREE`, csrc/bem.cu) against m2p_kernel of csrc/laplace.cu, which is green on
    hardware against the reference's treecode: same potential at the panel centres, sign and set by the target's BC."""
    out = subprocess.check_output([emu, "m2p"], timeout=600).decode()
    m = re.search(r"m2p: ([0-9.eE+-]+) max_potential ([0-9.eE+-]+)", out)
    assert m, out
    assert float(m.group(2)) > 1e-3 and float(m.group(1)) <= 1e-13


def test_bem_p2m_with_the_higher_gauss_rules(emu):
    """bem_p2m_kernel<0|1> with the 13-, 19-, 25- and 79-point rules of the reference's table equals the sum of its own
    runs with the one-point rules (pt_k, w_k): the (panel, quadrature point) -> lane mapping for K > 4."""
    out = subprocess.check_output([emu, "bem_rules"], timeout=900).decode()
    m = re.search(r"bem_rules: ([0-9.eE+-]+) max_multipole ([0-9.eE+-]+)", out)
    assert m, out
    assert float(m.group(2)) > 1e-4 and float(m.group(1)) <= 1e-12


def test_stokes_treecode_kernels_match_the_point_treecode_kernel(emu):
    """stokes_m2p_kernel (StokesSpherical, csrc/stokes.cu) and sbem_m2p_kernel (StokesSphericalBEM, csrc/stokes_bem.cu)
    against four runs of m2p_kernel of csrc/laplace.cu (green on hardware), one per expansion set, combined on the host
    as StokesSpherical.hpp:207-291 prescribes."""
    out = subprocess.check_output([emu, "stokes_m2p"], timeout=900).decode()
    m = re.search(r"stokes_m2p: point ([0-9.eE+-]+) bem ([0-9.eE+-]+) max_velocity ([0-9.eE+-]+)", out)
    assert m, out
    assert float(m.group(3)) > 1e-3 and float(m.group(1)) <= 1e-12 and float(m.group(2)) <= 1e-12


def test_yukawa_bem_treecode_kernel_matches_the_table_builder(emu):
    """yk_bem_m2p_kernel<0|1> (treecode of YukawaCartesianBEM, csrc/yukawa.cu: Taylor tables in lane-private arrays)
    against yk_table_kernel -- the block-cooperative builder behind the M2L that is green on hardware -- and a host
    dot product with the multipoles, orders 1, 4, 8, 10."""
    out = subprocess.check_output([emu, "ykm2p"], timeout=900).decode()
    m = re.search(r"ykm2p: ([0-9.eE+-]+) max_potential ([0-9.eE+-]+) point ([0-9.eE+-]+)", out)
    assert m, out
    assert float(m.group(2)) > 1e-4 and float(m.group(1)) <= 1e-12
    assert float(m.group(3)) <= 1e-12          # yk_m2p_kernel (point kernel): potential and gradient


@pytest.mark.parametrize("bcmix,as_written,K", [(0, False, 4), (1, False, 4), (2, False, 3), (0, True, 4), (1, True, 4),
                                               (2, True, 1)])
def test_near_field_kernels_match_the_oracle(emu, tmp_path, bcmix, as_written, K):
    verts = O.unit_sphere(4)                            # 512 panels, ncrit 64, theta 0.5: every list entry is near field
    n = len(verts)
    bc = np.zeros(n, np.int32) if bcmix == 0 else (np.ones(n, np.int32) if bcmix == 1 else (np.arange(n) % 3 == 1).astype(np.int32))
    mu, kfine = 0.02, 19
    orc = O.StokesBemOracle(verts, bc, mu=mu, K=K, kfine=kfine, as_written=as_written)
    t = orc.tree()
    assert len(t["lr"]) == 0 and t["p2p_off"][-1] >= 64
    q = np.random.default_rng(11 + bcmix).random((n, 3)) - 0.4
    want = orc.execute(q, 5, threads=1)
    boxes = t["boxes"]
    bb, be, leaf = boxes[:, 4].astype(np.uint32), boxes[:, 5].astype(np.uint32), boxes[:, 7]
    items = []
    for b in np.nonzero(leaf)[0]:                       # csrc/laplace.cu: chunks of <= 32 targets in leaf order
        for first in range(int(bb[b]), int(be[b]), 32):
            items.append((int(b), first, min(32, int(be[b]) - first), 0))
    items = np.array(items, np.int32)
    path = tmp_path / "near.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4q4id", n, len(boxes), len(items), len(t["p2p_idx"]), K, kfine, int(as_written), 0, mu))
        for a in (np.ascontiguousarray(verts, np.float64), bc, t["perm"].astype(np.uint32), bb, be,
                  t["p2p_off"].astype(np.int32), t["p2p_idx"].astype(np.int32), items, np.ascontiguousarray(q, np.float64)):
            f.write(np.ascontiguousarray(a).tobytes())
    out = subprocess.check_output([emu, "near", str(path)], timeout=900).decode()
    assert "near: n %d" % n in out
    got = np.fromfile(str(path) + ".out").reshape(n, 3)
    for k in range(3):
        assert O.rel_l2(got[:, k], want[:, k]) <= 1e-13


@pytest.mark.parametrize("kind,as_written", [(0, False), (0, True), (1, False), (2, False)])
def test_direct_matvec_kernels_match_the_oracle(emu, tmp_path, kind, as_written):
    """sbem_direct_kernel / bem_direct_kernel (fmmb_plan_direct_panels: Direct::matvec with the panel kernels) against
    the oracle's direct sums, which are bit-identical to the reference's Direct::matvec."""
    verts = O.unit_sphere(3)                             # 128 source panels
    n = len(verts)
    bc = (np.arange(n) % 3 == 1).astype(np.int32)
    sel = np.arange(0, n, 5)
    rng = np.random.default_rng(3 + kind)
    par = 0.02 if kind == 0 else 0.7
    if kind == 0:
        q = rng.random((n, 3)) - 0.4
        want = O.StokesBemOracle(verts, bc, mu=par, K=4, kfine=19, as_written=as_written).direct(q)[sel]
    elif kind == 1:
        q = rng.random(n) - 0.4
        want = O.BemOracle(verts, bc).direct(q, 4)[sel]
    else:
        q = rng.random(n) - 0.4
        want = O.YukawaBemOracle(verts, bc, par).direct(q, 4)[sel]
    path = tmp_path / "direct.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<2q4id", n, len(sel), 4, 19, int(as_written), kind, par))
        for a in (np.ascontiguousarray(verts, np.float64), np.ascontiguousarray(q, np.float64),
                  np.ascontiguousarray(verts[sel], np.float64), np.ascontiguousarray(bc[sel], np.int32)):
            f.write(a.tobytes())
    out = subprocess.check_output([emu, "direct", str(path)], timeout=900).decode()
    assert "direct: n %d" % n in out
    got = np.fromfile(str(path) + ".out").reshape(want.shape)
    assert O.rel_l2(got, want) <= 1e-13


@pytest.mark.parametrize("bcmix,as_written,treecode,P", [(0, False, False, 8), (2, True, False, 5), (1, False, True, 11), (2, False, True, 3)])
def test_host_sequencing_and_launch_configurations_of_the_stokes_bem_plan(emu, tmp_path, bcmix, as_written, treecode, P):
    """stokes_bem_setup + stokes_bem_execute of csrc/stokes_bem.cu -- the HOST functions as written, every launch with its
    own grid, block and dynamic shared size (guard bytes behind the shared segment) -- on the 512-panel sphere, whose
    lists have no far-field pairs (the translations, the one thing not emulated, are a stub): results against the oracle,
    FMM and treecode evaluators, orders on both sides of the 4-warp / 1-warp P2M configuration switch."""
    verts = O.unit_sphere(4)
    n = len(verts)
    bc = np.zeros(n, np.int32) if bcmix == 0 else (np.ones(n, np.int32) if bcmix == 1 else (np.arange(n) % 3 == 1).astype(np.int32))
    mu, kfine, K = 0.02, 19, 4
    orc = O.StokesBemOracle(verts, bc, mu=mu, K=K, kfine=kfine, as_written=as_written)
    t = orc.tree()
    assert len(t["lr"]) == 0
    q = np.random.default_rng(5 + bcmix).random((n, 3)) - 0.4
    want = orc.execute(q, P, threads=1, treecode=treecode)
    boxes = t["boxes"]
    bb, be, leaf = boxes[:, 4].astype(np.uint32), boxes[:, 5].astype(np.uint32), boxes[:, 7]
    items = []
    for b in np.nonzero(leaf)[0]:
        for first in range(int(bb[b]), int(be[b]), 32):
            items.append((int(b), first, min(32, int(be[b]) - first), 0))
    items = np.array(items, np.int32)
    path = tmp_path / "pipe.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4q4id", n, len(boxes), len(items), len(t["p2p_idx"]), K, kfine, int(as_written), 0, mu))
        for a in (np.ascontiguousarray(verts, np.float64), bc, t["perm"].astype(np.uint32), bb, be,
                  t["p2p_off"].astype(np.int32), t["p2p_idx"].astype(np.int32), items, np.ascontiguousarray(q, np.float64),
                  np.ascontiguousarray(t["geom"], np.float64), boxes[:, 1].astype(np.uint32), leaf.astype(np.int32),
                  np.array([P, int(treecode)], np.int32)):
            f.write(np.ascontiguousarray(a).tobytes())
    out = subprocess.check_output([emu, "pipeline", str(path)], timeout=900).decode()
    m = re.search(r"pipeline: n (\d+) launches (\d+) \(reported (\d+)\) translation_calls (\d+) guard_failures (\d+) repeatable (\d+) nnz (\d+)", out)
    assert m, out
    groups = len(set(bc.tolist()))
    assert int(m.group(1)) == n and int(m.group(5)) == 0 and int(m.group(6)) == 1
    assert int(m.group(4)) == 2 * 4 * groups                       # two matvecs x four sets per active group
    assert int(m.group(7)) == int(t["p2p_off"][-1] and sum((be[b] - bb[b]) * sum(be[s] - bb[s] for s in t["p2p_idx"][t["p2p_off"][b]:t["p2p_off"][b + 1]]) for b in np.nonzero(leaf)[0]))
    got = np.fromfile(str(path) + ".out").reshape(n, 3)
    for k in range(3):
        assert O.rel_l2(got[:, k], want[:, k]) <= 1e-13


@pytest.mark.parametrize("K,treecode,P", [(4, False, 8), (13, False, 6), (25, True, 9), (4, True, 5), (79, False, 3)])
def test_host_sequencing_of_the_laplace_bem_plan_with_high_rules_and_treecode(emu, tmp_path, K, treecode, P):
    """bem_setup + bem_execute of csrc/bem.cu as written (same emulation as above) on the 512-panel sphere with mixed
    boundary conditions: the Gauss rules above 4 points in setup / assembly / P2M and the treecode branch
    (bem_m2p_kernel) with their launch configurations; K = 4 without treecode is the hardware-verified path, as a
    control.  Results against the oracle (K = 79 is not in the oracle: finite values and repeatability only)."""
    verts = O.unit_sphere(4)
    n = len(verts)
    bc = (np.arange(n) % 2).astype(np.int32)
    orc = O.BemOracle(verts, bc)
    t = orc.tree()
    assert len(t["lr"]) == 0
    q = np.random.default_rng(K).random(n) - 0.4
    boxes = t["boxes"]
    bb, be, leaf = boxes[:, 4].astype(np.uint32), boxes[:, 5].astype(np.uint32), boxes[:, 7]
    items = []
    for b in np.nonzero(leaf)[0]:
        for first in range(int(bb[b]), int(be[b]), 32):
            items.append((int(b), first, min(32, int(be[b]) - first), 0))
    items = np.array(items, np.int32)
    path = tmp_path / "bem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4q4i", n, len(boxes), len(items), len(t["p2p_idx"]), K, P, int(treecode), 0))
        for a in (np.ascontiguousarray(verts, np.float64), bc, t["perm"].astype(np.uint32), bb, be,
                  t["p2p_off"].astype(np.int32), t["p2p_idx"].astype(np.int32), items, np.ascontiguousarray(q, np.float64),
                  np.ascontiguousarray(t["geom"], np.float64), boxes[:, 1].astype(np.uint32), leaf.astype(np.int32)):
            f.write(np.ascontiguousarray(a).tobytes())
    exe = os.path.join(os.path.dirname(emu), "emu_bem_pipeline")
    out = subprocess.check_output([exe, str(path)], timeout=900).decode()
    m = re.search(r"bem pipeline: n (\d+) launches (\d+) translation_calls (\d+) guard_failures (\d+) repeatable (\d+)", out)
    assert m, out
    assert int(m.group(1)) == n and int(m.group(4)) == 0 and int(m.group(5)) == 1 and int(m.group(3)) == 2 * 2
    got = np.fromfile(str(path) + ".out")
    assert np.isfinite(got).all()
    if K != 79:
        assert O.rel_l2(got, orc.execute(q, P, K, threads=1, treecode=treecode)) <= 1e-13


def test_device_gmres_on_vec3_unknowns_matches_the_replica_of_gmres_stokes(emu, tmp_path):
    """gmres_solve of csrc/gmres.cu as written -- fmmb_gmres with charge_dim 3, the order rule of GMRES_Stokes.hpp:229,
    Krylov basis, dot / axpy kernels and their launches -- over the emulated stokes_bem_execute, on the sphere problem of
    examples/StokesBEM.cpp with 512 panels; against the Python replica of the reference's GMRES_Stokes.hpp over the
    oracle matvec (tests/test_stokes_solve_replica.py): same iteration count, same order schedule, same residual
    history, same solution."""
    from test_stokes_solve_replica import gmres_stokes
    verts = O.unit_sphere(4)
    n = len(verts)
    bc = np.zeros(n, np.int32)
    mu, kfine, K, P = 1e-3, 19, 4, 8
    orc = O.StokesBemOracle(verts, bc, mu=mu, K=K, kfine=kfine, as_written=False)
    t = orc.tree()
    assert len(t["lr"]) == 0
    b = np.tile([4 * np.pi, 0.0, 0.0], (n, 1)).reshape(-1)
    x, it, res, lines = gmres_stokes(lambda v, p: orc.execute(v.reshape(-1, 3), p).reshape(-1), b, 1e-5, 8, 5)
    boxes = t["boxes"]
    bb, be, leaf = boxes[:, 4].astype(np.uint32), boxes[:, 5].astype(np.uint32), boxes[:, 7]
    items = []
    for bx in np.nonzero(leaf)[0]:
        for first in range(int(bb[bx]), int(be[bx]), 32):
            items.append((int(bx), first, min(32, int(be[bx]) - first), 0))
    items = np.array(items, np.int32)
    path = tmp_path / "gmres.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4q4id", n, len(boxes), len(items), len(t["p2p_idx"]), K, kfine, 0, 0, mu))
        for a in (np.ascontiguousarray(verts, np.float64), bc, t["perm"].astype(np.uint32), bb, be,
                  t["p2p_off"].astype(np.int32), t["p2p_idx"].astype(np.int32), items, np.zeros(3 * n),
                  np.ascontiguousarray(t["geom"], np.float64), boxes[:, 1].astype(np.uint32), leaf.astype(np.int32),
                  np.array([P, 0], np.int32)):
            f.write(np.ascontiguousarray(a).tobytes())
    out = subprocess.check_output([emu, "gmres", str(path)], timeout=1800).decode()
    m = re.search(r"gmres: iterations (\d+) final_residual ([0-9.eE+-]+) final_p (\d+) guard_failures (\d+) schedule((?: \d+)+)", out)
    assert m, out
    sched = [int(v) for v in m.group(5).split()]
    hist = [float(v) for v in re.search(r"residuals((?: [0-9.eE+-]+)+)", out).group(1).split()]
    assert int(m.group(1)) == it and int(m.group(4)) == 0
    assert sched[:len(lines)] == [p for _, _, p in lines] and len(sched) == it
    for h, (_, r, _) in zip(hist, lines):
        assert abs(h - r) <= 1e-6 * r
    assert abs(float(m.group(2)) - res) <= 1e-6 * res
    got = np.fromfile(str(path) + ".out")
    assert O.rel_l2(got, x) <= 1e-9


def _write_pipeline_file(path, verts, bc, t, q, K, kfine, as_written, mu, P, treecode, far, laplace=False):
    """File of `emu_stokes_bem pipeline`: tree and near-field lists, optionally the far-field structures."""
    n = len(verts)
    boxes = t["boxes"]
    bb, be, leaf = boxes[:, 4].astype(np.uint32), boxes[:, 5].astype(np.uint32), boxes[:, 7]
    items = []
    for b in np.nonzero(leaf)[0]:
        for first in range(int(bb[b]), int(be[b]), 32):
            items.append((int(b), first, min(32, int(be[b]) - first), 0))
    items = np.array(items, np.int32)
    with open(path, "wb") as f:
        if laplace:      # emu_bem_pipeline: K, P, treecode in the header, scalar charges
            f.write(struct.pack("<4q4i", n, len(boxes), len(items), len(t["p2p_idx"]), K, P, int(treecode), 0))
        else:
            f.write(struct.pack("<4q4id", n, len(boxes), len(items), len(t["p2p_idx"]), K, kfine, int(as_written), 0, mu))
        for a in (np.ascontiguousarray(verts, np.float64), bc, t["perm"].astype(np.uint32), bb, be,
                  t["p2p_off"].astype(np.int32), t["p2p_idx"].astype(np.int32), items, np.ascontiguousarray(q, np.float64),
                  np.ascontiguousarray(t["geom"], np.float64), boxes[:, 1].astype(np.uint32), leaf.astype(np.int32)):
            f.write(np.ascontiguousarray(a).tobytes())
        if not laplace:
            f.write(np.array([P, int(treecode)], np.int32).tobytes())
        if far:
            nb = len(boxes)
            lr = t["lr"].astype(np.int64)
            order = np.argsort(lr[:, 1], kind="stable")                 # target-major, list order inside a target
            m2l_off = np.zeros(nb + 1, np.int32)
            np.add.at(m2l_off, lr[:, 1] + 1, 1)
            m2l_off = np.cumsum(m2l_off).astype(np.int32)
            has_local = np.zeros(nb, np.int32)
            has_local[lr[:, 1]] = 1
            level = boxes[:, 6].astype(np.int64)
            nlevels = int(level.max()) + 1
            for b in range(1, nb):                                       # parents precede children
                if has_local[boxes[b, 1]]:
                    has_local[b] = 1
            level_off = np.searchsorted(level, np.arange(nlevels + 1)).astype(np.int32)
            f.write(struct.pack("<q2i", len(lr), nlevels, 0))
            for a in (boxes[:, 0].astype(np.uint32), boxes[:, 2].astype(np.uint32), boxes[:, 3].astype(np.uint32), level_off,
                      m2l_off, lr[order, 0].astype(np.int32), has_local):
                f.write(np.ascontiguousarray(a).tobytes())


@pytest.mark.parametrize("name,treecode", [("stokes_bem_2048_p6_bc2", False), ("stokes_bem_tree_asis_2048_p6_bc2", True)])
def test_whole_stokes_bem_matvec_with_far_field_against_the_reference_fixtures(emu, tmp_path, name, treecode):
    """The complete StokesSphericalBEM matvec of csrc/stokes_bem.cu on the CPU: host functions as written, all their
    kernels, and for the translations the per-pair kernels of csrc/laplace.cu (M2M sweep, M2L per target box, L2L sweep:
    the path behind P > 8 / m2l_mode 1, with laplace_translations' launch configurations; the class-batched DMMA path is
    PTX and is not emulated).  2 048 panels with a real far field, against the golden fixtures of the reference."""
    import json
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    m = json.loads(str(g["meta"]))
    orc = O.StokesBemOracle(g["verts"], g["bc"], mu=m["mu"], K=m["K"], kfine=m["kfine"], as_written=m["as_written"],
                            ncrit=m["ncrit"], theta=m["theta"])
    t = orc.tree()
    assert len(t["lr"]) > 1000
    path = tmp_path / "whole.bin"
    _write_pipeline_file(path, g["verts"], g["bc"].astype(np.int32), t, g["charges"], m["K"], m["kfine"], m["as_written"],
                         m["mu"], m["P"], treecode, far=True)
    out = subprocess.check_output([emu, "pipeline", str(path)], timeout=3000).decode()
    mm = re.search(r"guard_failures (\d+) repeatable (\d+)", out)
    assert mm and int(mm.group(1)) == 0, out
    got = np.fromfile(str(path) + ".out").reshape(-1, 3)
    for k in range(3):
        assert O.rel_l2(got[:, k], g["results"][:, k]) <= 1e-10


@pytest.mark.parametrize("name", ["laplace_bem_2048_p6_k13_bc1", "laplace_bem_tree_2048_p6_k4_bc0"])
def test_whole_laplace_bem_matvec_with_far_field_against_the_reference_fixtures(emu, tmp_path, name):
    """The same for csrc/bem.cu: bem_setup + bem_execute as written with the 13-point rule (FMM) and with the treecode
    branch, per-pair translation kernels of csrc/laplace.cu, 2 048 panels, against the golden fixtures of the reference."""
    import json
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    m = json.loads(str(g["meta"]))
    n = len(g["verts"])
    bc = np.full(n, m["bc"], np.int32)
    t = O.BemOracle(g["verts"], bc, ncrit=m["ncrit"], theta=m["theta"]).tree()
    assert len(t["lr"]) > 1000
    path = tmp_path / "whole_bem.bin"
    _write_pipeline_file(path, g["verts"], bc, t, g["charges"], m["K"], 0, 0, 0.0, m["P"], bool(m.get("treecode", 0)), far=True,
                         laplace=True)
    out = subprocess.check_output([os.path.join(os.path.dirname(emu), "emu_bem_pipeline"), str(path)], timeout=3000).decode()
    mm = re.search(r"guard_failures (\d+)", out)
    assert mm and int(mm.group(1)) == 0, out
    got = np.fromfile(str(path) + ".out")
    assert O.rel_l2(got, g["results"]) <= 1e-10

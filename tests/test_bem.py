"""LaplaceSphericalBEM (config C2 of BASELINE.json): oracle pins on the CPU, CUDA parity on the GPU.

Reference values (unmodified reference, oracle/_ref, 1 thread; also listed in SURVEY.md section 8c):
  2 048-panel sphere, P=8, K=4, random charges: FMM vs Direct 2.665e-06 (G), 2.296e-05 (dG/dn)
  LaplaceBEM 512 panels  -p 8 -k 4 -solver_tol 1e-6: 7 iterations, p = 8,8,8,7,6,4, final 3.1066e-07
  LaplaceBEM 32 768 panels (C2): 16 iterations, p = 8,6,5,5,5,4,4,3,3,3,2,2,2,1,1, final 8.4431e-07,
      relative error 4.862e-03; -fixed_p: 15 iterations, final 9.1278e-07
"""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import fmm_bem_relaxed_b200 as F
from conftest import ROOT

BIN = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin")
TOL = 1e-10


# ------------------------------------------------------------------ CPU: the restatement itself
def test_oracle_sphere_and_analytic_limits():
    v = O.unit_sphere(4)
    assert v.shape == (512, 3, 3)
    assert np.allclose(np.linalg.norm(v.reshape(-1, 3), axis=1), 1.0)
    orc = O.BemOracle(v, 0)
    # single layer of a constant density on the unit sphere: potential 4 pi R = 4 pi on the surface
    r = orc.direct(np.ones(512), K=4)
    assert abs(r.mean() / (4 * np.pi) - 1) < 2e-2
    # double layer (bc = 1): the self term is exactly 2 pi and rows sum to about 2 pi + 2 pi = 4 pi ... sign as the
    # reference defines it; what is pinned here is the self term
    orc1 = O.BemOracle(v[:1], 1)
    assert orc1.direct(np.ones(1), K=4)[0] == 2 * np.pi


def test_oracle_fmm_matches_direct_like_the_reference():
    v = O.unit_sphere(5)
    q = np.random.default_rng(5).random(len(v))
    for bc, ref_err in ((0, 2.665e-06), (1, 2.296e-05)):
        orc = O.BemOracle(v, bc)
        fmm, d = orc.execute(q, 8, 4), orc.direct(q, 4)
        err = O.rel_l2(fmm, d)
        assert 0.5 * ref_err < err < 2 * ref_err      # other random charges than the reference run: same size


def test_oracle_kernel_branches():
    """Near panels take the semi-analytical / 16-point branch, far ones the K-point rule; both must be
    continuous across the switch sqrt(2A)/dist = 0.5 to the accuracy of the rules."""
    v = O.unit_sphere(3)[:1]
    area = 0.5 * np.linalg.norm(np.cross(v[0, 2] - v[0, 0], v[0, 1] - v[0, 0]))
    c = O.panel_centers(v)[0]
    d_switch = np.sqrt(2 * area) / 0.5
    for bc in (0, 1):
        vals = []
        for d in (0.999 * d_switch, 1.001 * d_switch):
            far = c + d * c / np.linalg.norm(c)
            tgt = np.stack([far, far, far])[None]              # degenerate panel: only its centre matters
            both = np.concatenate([v, tgt])
            res = O.BemOracle(both, bc).direct(np.array([1.0, 0.0]), K=4)
            vals.append(res[1])
        assert abs(vals[0] - vals[1]) / abs(vals[1]) < 2e-2


# ------------------------------------------------------------------ GPU
def run_bin(exe, *args):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.skip(path + " not built")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    return subprocess.check_output([path] + list(args), env=env, timeout=900, cwd=BIN).decode()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["laplace_bem_2048_p6_k4_bc0", "laplace_bem_2048_p6_k4_bc1"])
def test_golden_fixtures_of_the_reference(name):
    """FMM matvec of the unmodified reference class (oracle/_ref/ref_bem, tests/golden/make_golden.py --laplace-bem)."""
    import json
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    m = json.loads(str(g["meta"]))
    opts = F.FMMOptions()
    opts.set_max_per_box(m["ncrit"])
    opts.set_mac_theta(m["theta"])
    plan = F.FMM_plan(F.LaplaceSphericalBEM(m["P"], m["K"]), F.Panels(g["verts"], m["bc"]), opts)
    assert O.rel_l2(plan.execute(g["charges"]), g["results"]) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("rec,P,K", [(4, 8, 4), (5, 8, 4), (5, 5, 3), (6, 6, 1), (7, 8, 4)])
def test_bem_matvec_vs_oracle(rec, P, K):
    v = O.unit_sphere(rec)
    q = np.random.default_rng(rec).random(len(v)) - 0.3
    for bc in (0, 1):
        orc = O.BemOracle(v, bc)
        plan = F.FMM_plan(F.LaplaceSphericalBEM(P, K), F.Panels(v, bc))
        assert np.array_equal(plan.tree()["lr"], orc.tree()["lr"])
        assert np.array_equal(plan.tree()["perm"], orc.tree()["perm"])
        res = plan.execute(q)
        assert res.shape == (len(v),)
        assert O.rel_l2(res, orc.execute(q, P, K)) <= TOL
        assert plan.info().n_near_entries == plan.info().n_p2p_body_pairs


def _two_scale_mesh():
    """A sphere with a 20x smaller one just outside its surface: leaves of the coarse mesh see long lists of small
    neighbour leaves (hundreds: the chunked list walk of the split near-field kernels).  With ncrit = 8 the tree is 8
    levels deep, inside the 10 levels of the reference's 30-bit Morton codes."""
    big = O.unit_sphere(4)
    small = 0.05 * O.unit_sphere(5) + np.array([1.075, 0.0, 0.0])
    return np.concatenate([big, small])


def _largest_first_chunk_rows(tree):
    """Rows (source panels) in the first 32 source boxes of a target box's near-field list, maximised over the boxes."""
    off, idx, boxes = tree["p2p_off"], tree["p2p_idx"], tree["boxes"]
    size = (boxes[:, 5] - boxes[:, 4]).astype(np.int64)
    return max(int(size[idx[off[b]:min(off[b] + 32, off[b + 1])]].sum()) for b in range(len(off) - 1))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["long lists", "large leaves"])
def test_cached_near_field_with_long_source_lists(case):
    """bem_near_split_kernel (eight warps per work item; the list is walked 32 source leaves at a time, a chunk's
    charges staged in shared memory and its rows split evenly over the warps) against the one-warp-per-item kernel it
    replaced and against the oracle.  "long lists": two-scale mesh, ncrit = 8, target leaves with hundreds of source
    leaves (several chunks).  "large leaves": ncrit = 250, chunks with more rows than the staging buffer (2 048), which
    take the per-leaf path.  Deterministic across calls."""
    v, ncrit = (_two_scale_mesh(), 8) if case == "long lists" else (O.unit_sphere(6), 250)
    n = len(v)
    q = np.random.default_rng(11).random(n) - 0.4
    opts = F.FMMOptions()
    opts.set_max_per_box(ncrit)
    for bc in (0, 1):
        plan = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(v, bc), opts)
        t = plan.tree()
        if case == "long lists":
            assert np.diff(t["p2p_off"]).max() > 100
        else:
            assert _largest_first_chunk_rows(t) > 2048
        res = plan.execute(q)
        assert np.array_equal(plan.execute(q), res) and np.array_equal(plan.execute(q), res)
        plan.set_option("bem_near_kernel", 0)
        assert O.rel_l2(res, plan.execute(q)) <= 1e-13
        assert O.rel_l2(res, O.BemOracle(v, bc, ncrit=ncrit).execute(q, 6, 4)) <= TOL


@pytest.mark.gpu
def test_bem_mixed_boundary_conditions_and_relaxation():
    v = O.unit_sphere(5)
    n = len(v)
    bc = (np.arange(n) % 3 == 0).astype(np.int32)          # both expansion sets active
    q = np.random.default_rng(1).random(n)
    orc = O.BemOracle(v, bc)
    plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, bc))
    for p in (8, 6, 3, 1, 8):
        plan.kernel().set_p(p)
        assert O.rel_l2(plan.execute(q), orc.execute(q, p, 4)) <= TOL


@pytest.mark.gpu
def test_c2_size_counts():
    """Config C2: 32 768 panels -> the tree statistics of SURVEY.md section 8."""
    v = O.unit_sphere(7)
    plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 0))
    i = plan.info()
    assert (i.n_bodies, i.n_boxes, i.n_leaves, i.n_m2l_pairs, i.n_p2p_box_pairs, i.n_near_entries) == (
        32768, 1305, 1040, 41516, 17480, 17077856)


def parse_gmres(out):
    its = [(int(m.group(1)), float(m.group(2)), int(m.group(3)))
           for m in re.finditer(r"it: (\d+), res: ([0-9.eE+-]+), fmm_req_p: (\d+)", out)]
    fin = re.search(r"Final residual: ([0-9.eE+-]+), after (\d+) iterations", out)
    rel = float(re.search(r"relative error: ([0-9.eE+-]+)", out).group(1))
    ext = float(re.search(r"external phi: ([0-9.eE+-]+)", out).group(1))
    return its, float(fin.group(1)), int(fin.group(2)), rel, ext


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["laplace_bem", "laplace_bem_refgmres"])
def test_relaxed_gmres_512_panels_same_iterations_as_reference(exe):
    its, final, niter, rel, ext = parse_gmres(run_bin(exe, "-recursions", "4", "-p", "8", "-k", "4", "-solver_tol", "1e-6"))
    assert niter == 7
    assert [p for _, _, p in its] == [8, 8, 8, 7, 6, 4]
    assert [r for _, r, _ in its] == [7.878e-04, 2.986e-04, 1.081e-04, 3.370e-05, 1.043e-05, 2.522e-06]
    assert final == 3.1066e-07 and rel == 1.512e-02 and ext == 0.19071


@pytest.mark.gpu
def test_relaxed_gmres_c2_same_iterations_as_reference():
    out = run_bin("laplace_bem", "-recursions", "7", "-p", "8", "-k", "4", "-ncrit", "64", "-theta", "0.5",
                  "-solver_tol", "1e-6")
    its, final, niter, rel, ext = parse_gmres(out)
    assert niter == 16
    assert [p for _, _, p in its] == [8, 6, 5, 5, 5, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1]
    assert its[0][1] == 3.831e-05 and its[-1][1] == 1.007e-06
    assert final == 8.4431e-07 and rel == 4.862e-03 and ext == 0.19242
    out = run_bin("laplace_bem", "-recursions", "7", "-p", "8", "-k", "4", "-solver_tol", "1e-6", "-fixed_p")
    its, final, niter, rel, ext = parse_gmres(out)
    assert niter == 15 and final == 9.1278e-07 and rel == 4.847e-03


@pytest.mark.gpu
def test_device_resident_gmres_same_iterations_as_reference():
    """fmmb_gmres (Krylov basis and BLAS-1 on the GPU, one host sync per iteration) takes the same decisions as the
    reference's GMRES.hpp: iteration counts and the order of every iteration are identical; residuals agree to the
    printed digits up to the rounding of the (parallel) dot products."""
    its, final, niter, rel, ext = parse_gmres(run_bin("laplace_bem", "-recursions", "4", "-p", "8", "-k", "4",
                                                      "-solver_tol", "1e-6", "-device_gmres"))
    assert niter == 7 and [p for _, _, p in its] == [8, 8, 8, 7, 6, 4]
    for (_, r, _), want in zip(its, [7.878e-04, 2.986e-04, 1.081e-04, 3.370e-05, 1.043e-05, 2.522e-06]):
        assert abs(r - want) <= 2e-3 * want
    assert abs(final - 3.1066e-07) <= 2e-3 * 3.1066e-07 and rel == 1.512e-02 and ext == 0.19071
    out = run_bin("laplace_bem", "-recursions", "7", "-p", "8", "-k", "4", "-ncrit", "64", "-theta", "0.5",
                  "-solver_tol", "1e-6", "-device_gmres")
    its, final, niter, rel, ext = parse_gmres(out)
    assert niter == 16 and [p for _, _, p in its] == [8, 6, 5, 5, 5, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1]
    assert abs(final - 8.4431e-07) <= 2e-3 * 8.4431e-07 and rel == 4.862e-03 and ext == 0.19242
    out = run_bin("laplace_bem", "-recursions", "6", "-p", "8", "-k", "4", "-solver_tol", "1e-6", "-device_gmres", "-diagonal")
    host = run_bin("laplace_bem", "-recursions", "6", "-p", "8", "-k", "4", "-solver_tol", "1e-6", "-diagonal")
    a, b = parse_gmres(out), parse_gmres(host)
    assert a[2] == b[2] and [p for _, _, p in a[0]] == [p for _, _, p in b[0]] and abs(a[1] - b[1]) <= 2e-3 * b[1]


@pytest.mark.gpu
def test_python_gmres_solves_the_first_kind_equation():
    v = O.unit_sphere(5)
    n = len(v)
    plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 0))
    b = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, 1)).execute(np.ones(n))   # dphi/dn = 1 on the sphere
    rep = F.GMRES(plan, np.zeros(n), b, F.SolverOptions(residual=1e-6, max_iters=200, restart=200, max_p=8))
    assert rep["iterations"] < 30 and rep["final_residual"] < 1e-6
    assert rep["p_schedule"][0] == 8 and min(rep["p_schedule"]) < 8          # the order was relaxed
    assert abs(rep["x"].mean() - 1.0) < 3e-2                                    # exact solution: 1
    plan.kernel().set_p(8)
    assert O.rel_l2(plan.execute(rep["x"]), b) < 1e-3                           # true residual at full order


@pytest.mark.gpu
def test_reference_laplacebem_driver_unchanged():
    """bin/ref_LaplaceBEM = the reference's examples/LaplaceBEM.cpp compiled UNCHANGED, with the reference's own
    GMRES.hpp / fmgmres.hpp / LocalPC.hpp / BlockDiagonalPC.hpp / Preconditioner.hpp / Triangulation.hpp / MshReader.hpp;
    only FMM_plan, FMMOptions, the kernel classes, Vec, Mat3 and Direct come from hostcxx/ (i.e. the GPU engine).
    It must print what the reference prints."""
    its, final, niter, rel, ext = parse_gmres(run_bin("ref_LaplaceBEM", "-recursions", "4", "-p", "8", "-k", "4",
                                                      "-solver_tol", "1e-6"))
    assert niter == 7 and [p for _, _, p in its] == [8, 8, 8, 7, 6, 4]
    assert [r for _, r, _ in its] == [7.878e-04, 2.986e-04, 1.081e-04, 3.370e-05, 1.043e-05, 2.522e-06]
    assert final == 3.1066e-07 and rel == 1.512e-02 and ext == 0.19071
    out = run_bin("ref_LaplaceBEM", "-recursions", "7", "-p", "8", "-k", "4", "-ncrit", "64", "-theta", "0.5",
                  "-solver_tol", "1e-6")
    its, final, niter, rel, ext = parse_gmres(out)
    assert niter == 16 and [p for _, _, p in its] == [8, 6, 5, 5, 5, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1]
    assert its[0][1] == 3.831e-05 and its[-1][1] == 1.007e-06
    assert final == 8.4431e-07 and rel == 4.862e-03 and ext == 0.19242
    # diagonal preconditioner (Preconditioners::Diagonal over plan.source_begin()/source_end()) and the 2nd-kind equation:
    # same lines as our own driver with the same options
    for extra in (("-diagonal",), ("-second_kind",)):
        a = parse_gmres(run_bin("ref_LaplaceBEM", "-recursions", "5", "-p", "8", "-k", "4", "-solver_tol", "1e-6", *extra))
        b = parse_gmres(run_bin("laplace_bem", "-recursions", "5", "-p", "8", "-k", "4", "-solver_tol", "1e-6", *extra))
        assert a == b


@pytest.mark.gpu
def test_reference_msh_reader_feeds_the_gpu_plan(tmp_path):
    """SURVEY 8(f) rank 4, mesh readers: the reference's MshReader.hpp (compiled unchanged into bin/ref_LaplaceBEM) reads
    a Gmsh 2 ASCII file -- nodes written with 17 digits, a non-triangular element that the reader skips, triangles with
    the (v1, v3, v2) order the reader turns back (MshReader.hpp:84-89) -- and the solve over the GPU plan prints exactly
    the lines of the generated sphere with the same 512 panels."""
    v = O.unit_sphere(4)                                  # (512, 3, 3), the panels of Triangulation::UnitSphere(4)
    n = v.shape[0]
    msh = tmp_path / "sphere.msh"
    with open(msh, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % (3 * n))
        for i, p in enumerate(v.reshape(-1, 3)):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, p[0], p[1], p[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % (n + 1))
        for e in range(n):                                # element: number, type 2 (triangle), 2 tags, three nodes
            f.write("%d 2 2 0 1 %d %d %d\n" % (e + 1, 3 * e + 1, 3 * e + 3, 3 * e + 2))
        f.write("%d 1 2 0 1 1 2\n" % (n + 1))             # a line element: skipped by the reader
        f.write("$EndElements\n")
    args = ("-p", "8", "-k", "4", "-solver_tol", "1e-6")
    out = run_bin("ref_LaplaceBEM", "-mesh", str(msh), *args)
    assert "num_nodes: %d" % (3 * n) in out and "1 elements skipped" in out
    its, final, niter, rel, ext = parse_gmres(out)
    assert niter == 7 and [p for _, _, p in its] == [8, 8, 8, 7, 6, 4]
    assert [r for _, r, _ in its] == [7.878e-04, 2.986e-04, 1.081e-04, 3.370e-05, 1.043e-05, 2.522e-06]
    assert final == 3.1066e-07 and rel == 1.512e-02 and ext == 0.19071

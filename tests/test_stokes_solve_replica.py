"""CPU suite: the whole Stokes BEM solve of the reference's examples/StokesBEM.cpp, replayed without the reference and
without a GPU -- a line-by-line Python replica of examples/BEM/GMRES_Stokes.hpp:170-300 (flat 3n arrays, modified
Gram-Schmidt, Givens rotations, p = max(p_min, predict_p(|resid|) - 1) before every inner matvec) over the oracle's
StokesSphericalBEM matvec, which is bit-identical to the reference's (tests/test_oracle.py).

It must print what the UNMODIFIED reference driver prints on one thread (oracle/_ref/StokesBEM -recursions 5 -p 8 -k 4
-solver_tol 1e-5; recorded below).  This pins, on the CPU, the numbers the GPU tests expect from bin/ref_StokesBEM,
bin/stokes_bem and fmmb_gmres (tests/test_zz_stokes_bem.py), including two things one has to know about that driver:
its "Fx" line is x[0][0] times the total area (the summation loop never advances its index, StokesBEM.cpp:341-352),
while the solution it writes to out.charge has the drag 0.018695.
"""
import math

import numpy as np

import oracle_lib as O

REFERENCE_LINES = [(1, 1.540e-03, 7), (2, 5.298e-04, 7), (3, 2.929e-04, 5), (4, 1.364e-04, 5), (5, 7.147e-05, 5),
                   (6, 4.341e-05, 5), (7, 2.744e-05, 5), (8, 1.578e-05, 5), (9, 1.377e-05, 5)]
REFERENCE_FINAL = (8.1330e-06, 10)
REFERENCE_PRINTED_FX = 0.01911          # = x[0][0] * sum(Area)
REFERENCE_X0 = (0.00152575, 1.55091e-05, 1.5509e-05)     # first line of the out.charge it writes


def predict_p(eps, tol, max_p):         # SolverOptions::predict_p, BOURAS (examples/BEM/SolverOptions.hpp:25-38)
    nu = min(tol / min(eps, 1.0), 1.0)
    return min(int(math.ceil(-math.log2(nu))), max_p)


def gmres_stokes(matvec, b, tol, max_p, p_min, restart=100):
    n3 = b.size
    x = np.zeros(n3)
    normb = np.linalg.norm(b)
    w = matvec(x, max_p) - b                       # the kernel still has the order it was built with
    beta = np.linalg.norm(w)
    V = [-w / beta]
    s = np.zeros(restart + 1)
    s[0] = beta
    cs, sn, H = np.zeros(restart), np.zeros(restart), np.zeros((restart + 1, restart))
    resid = s[0] / normb
    lines, it = [], 0
    for i in range(restart):
        it += 1
        p = max(p_min, predict_p(abs(resid), tol, max_p) - 1)
        w = matvec(V[i], p).copy()
        for k in range(i + 1):
            H[k, i] = w @ V[k]
            w -= H[k, i] * V[k]
        H[i + 1, i] = np.linalg.norm(w)
        V.append(w / H[i + 1, i])
        for k in range(i):
            H[k, i], H[k + 1, i] = cs[k] * H[k, i] + sn[k] * H[k + 1, i], -sn[k] * H[k, i] + cs[k] * H[k + 1, i]
        dx, dy = H[i, i], H[i + 1, i]
        if dy == 0:
            cs[i], sn[i] = 1.0, 0.0
        elif abs(dy) > abs(dx):
            t = dx / dy
            sn[i] = 1 / math.sqrt(1 + t * t)
            cs[i] = t * sn[i]
        else:
            t = dy / dx
            cs[i] = 1 / math.sqrt(1 + t * t)
            sn[i] = t * cs[i]
        H[i, i], H[i + 1, i] = cs[i] * H[i, i] + sn[i] * H[i + 1, i], -sn[i] * H[i, i] + cs[i] * H[i + 1, i]
        s[i], s[i + 1] = cs[i] * s[i] + sn[i] * s[i + 1], -sn[i] * s[i] + cs[i] * s[i + 1]
        resid = s[i + 1] / normb
        if abs(resid) < tol:
            break
        lines.append((it, abs(resid), p))
    y = np.linalg.solve(np.triu(H[:i + 1, :i + 1]), s[:i + 1])
    for j in range(i + 1):
        x += y[j] * V[j]
    return x, it, abs(resid), lines


def test_replica_prints_the_reference_drivers_lines():
    verts = O.unit_sphere(5)
    n = len(verts)
    orc = O.StokesBemOracle(verts, 0, mu=1e-3, K=4, kfine=19, as_written=False)
    b = np.tile([4 * np.pi, 0.0, 0.0], (n, 1)).reshape(-1)
    x, it, res, lines = gmres_stokes(lambda v, p: orc.execute(v.reshape(-1, 3), p).reshape(-1), b, 1e-5, 8, 5)
    assert it == REFERENCE_FINAL[1] and abs(res - REFERENCE_FINAL[0]) <= 2e-3 * REFERENCE_FINAL[0]
    assert len(lines) == len(REFERENCE_LINES)
    for (i, r, p), (wi, wr, wp) in zip(lines, REFERENCE_LINES):
        assert (i, p) == (wi, wp) and "%.3e" % r == "%.3e" % wr, (lines, REFERENCE_LINES)
    x = x.reshape(n, 3)
    for got, want in zip(x[0], REFERENCE_X0):
        assert abs(got - want) <= 1e-5 * abs(want) + 1e-11
    area = 0.5 * np.linalg.norm(np.cross(verts[:, 2] - verts[:, 0], verts[:, 1] - verts[:, 0]), axis=1)
    assert abs(x[0, 0] * area.sum() - REFERENCE_PRINTED_FX) <= 1e-5      # what the driver prints as "Fx"
    drag = (x[:, 0] * area).sum()
    assert abs(drag - 0.018695) <= 1e-6                                   # the drag of the solution it writes
    assert abs(drag - 6 * np.pi * 1e-3) <= 0.01 * 6 * np.pi * 1e-3        # Stokes' law to 1 %
    # its right-hand-side diagnostic (:256-271): the TRACTION plan applied to charges (1, 0, 0)
    rhs = O.StokesBemOracle(verts, 1, mu=1e-3, K=4, kfine=19, as_written=False).execute(np.tile([1.0, 0, 0], (n, 1)), 8)
    assert "%.4e" % np.sum(np.abs(rhs[:, 0] - 4 * np.pi) / 4 / np.pi) == "2.0203e+03"


def test_as_written_entries_converge_to_stokes_law_too():
    verts = O.unit_sphere(5)
    n = len(verts)
    orc = O.StokesBemOracle(verts, 0, mu=1e-3, K=4, kfine=19, as_written=True)
    b = np.tile([4 * np.pi, 0.0, 0.0], (n, 1)).reshape(-1)
    x, it, res, _ = gmres_stokes(lambda v, p: orc.execute(v.reshape(-1, 3), p).reshape(-1), b, 1e-5, 8, 5)
    area = 0.5 * np.linalg.norm(np.cross(verts[:, 2] - verts[:, 0], verts[:, 1] - verts[:, 0]), axis=1)
    drag = (x.reshape(n, 3)[:, 0] * area).sum()
    assert it == 22 and res < 1e-5 and abs(drag - 6 * np.pi * 1e-3) <= 0.01 * 6 * np.pi * 1e-3


def test_reference_gmres_stokes_header_over_the_mirror_class_prints_the_replicas_lines(tmp_path):
    """tests/host/stokes_dense.cpp: the reference's examples/BEM/GMRES_Stokes.hpp, compiled UNCHANGED, solving a dense
    system assembled with operator() of the mirror class hostcxx/StokesSphericalBEM.hpp (128 panels: the FMM matvec of
    this size is all near field, so the oracle matvec is the same matrix).  Its lines equal the replica's: the replica
    IS that header's algorithm.  Skipped where the reference sources are absent."""
    import os
    import re
    import subprocess
    import pytest
    from conftest import ROOT
    ref = "/root/reference/examples/BEM"
    if not os.path.exists(os.path.join(ref, "GMRES_Stokes.hpp")):
        pytest.skip("reference sources are only present in the build container")
    hostcxx = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx")
    exe = str(tmp_path / "stokes_dense")
    subprocess.check_call(["g++", "-std=gnu++14", "-O1", "-I", ref, "-I", hostcxx, "-include", os.path.join(hostcxx, "ref_prelude.hpp"),
                           os.path.join(ROOT, "tests", "host", "stokes_dense.cpp"), "-o", exe])
    out = subprocess.check_output([exe], cwd=str(tmp_path), timeout=300).decode()
    got = [(int(a), b, int(c)) for a, b, c in re.findall(r"it: (\d+), res: ([0-9.eE+-]+), fmm_req_p: (\d+)", out)]
    verts = O.unit_sphere(3)
    n = len(verts)
    orc = O.StokesBemOracle(verts, 0, mu=1e-3, K=4, kfine=19, as_written=False)
    assert len(orc.tree()["lr"]) == 0
    b = np.tile([4 * np.pi, 0.0, 0.0], (n, 1)).reshape(-1)
    x, it, res, lines = gmres_stokes(lambda v, p: orc.execute(v.reshape(-1, 3), p).reshape(-1), b, 1e-5, 8, 5)
    assert [(i, "%.3e" % r, p) for i, r, p in lines] == got and len(got) >= 5
    m = re.search(r"Final residual: ([0-9.eE+-]+), after (\d+) iterations", out)
    assert int(m.group(2)) == it and abs(float(m.group(1)) - res) <= 2e-3 * res
    area = 0.5 * np.linalg.norm(np.cross(verts[:, 2] - verts[:, 0], verts[:, 1] - verts[:, 0]), axis=1)
    fx = float(re.search(r"Fx: ([0-9.]+)", out).group(1))
    assert abs(fx - (x.reshape(n, 3)[:, 0] * area).sum()) <= 1e-5

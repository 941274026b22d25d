"""GPU suite (-m gpu): StokesSpherical (Stokeslet and stresslet) through the C ABI against the golden
fixtures of the reference (tests/golden/stoke*let_*.npz, made by tests/golden/make_golden.py from
oracle/_ref/ref_stokeslet = unmodified reference, ref_stresslet = reference + the two compile patches of
SURVEY.md section 8(c)) and against the oracle restatement, which is bit-identical to both.

Tolerance: relative L2 <= 1e-10 (BASELINE.json north_star) on the velocity field.
"""
import os

import numpy as np
import pytest

import oracle_lib as O
import fmm_bem_relaxed_b200 as F
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = 1e-10


def make_plan(points, P, stresslet, ncrit=64, theta=0.5):
    opts = F.FMMOptions()
    opts.set_mac_theta(theta)
    opts.set_max_per_box(ncrit)
    return F.FMM_plan(F.StokesSpherical(P, stresslet), points, opts)


@pytest.mark.parametrize("name,stresslet,P,ncrit,theta", [
    ("stokeslet_drand48_n3000_p5", False, 5, 32, 0.5),
    ("stresslet_drand48_n3000_p6", True, 6, 32, 0.5),
    ("stresslet_two_scale_n4000_p7", True, 7, 12, 0.6),
    ("stokeslet_two_scale_n4000_p4", False, 4, 12, 0.6),
])
def test_golden_fixtures(name, stresslet, P, ncrit, theta):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    plan = make_plan(g["points"], P, stresslet, ncrit, theta)
    i = plan.info()
    assert (i.charge_dim, i.result_dim) == (6 if stresslet else 3, 3)
    res = plan.execute(g["charges"])
    assert res.shape == g["results"].shape
    for k in range(3):
        assert O.rel_l2(res[:, k], g["results"][:, k]) <= TOL
    # deterministic (target ownership, fixed summation order), also through the CUDA-graph replay
    for _ in range(3):
        assert np.array_equal(plan.execute(g["charges"]), res)
    # brute force on the GPU (Direct::matvec with the kernel's own pair rule) against the oracle's
    d = F.Direct.matvec(plan, g["charges"], g["points"][:200])
    assert O.rel_l2(d, O.stokes_direct(g["points"], g["charges"], g["points"][:200], stresslet)) <= 1e-12


@pytest.mark.parametrize("stresslet", [False, True])
@pytest.mark.parametrize("n,P,ncrit", [(20000, 8, 64), (30000, 3, 100), (8000, 10, 40)])
def test_vs_oracle(stresslet, n, P, ncrit):
    """serialrun_stresslet.cpp:98-128 style inputs: drand48 points, charges (U, U, U[, 1, 0, 0])."""
    pts, _ = O.drand48_inputs(n)
    rng = np.random.default_rng(n + P)
    if stresslet:
        q = np.hstack([rng.random((n, 3)), np.tile([1.0, 0.0, 0.0], (n, 1))])
    else:
        q = rng.random((n, 3))
    ref = O.Oracle(pts, ncrit, 0.5).stokes_execute(q, P, stresslet)
    plan = make_plan(pts, P, stresslet, ncrit)
    res = plan.execute(q)
    assert O.rel_l2(res, ref) <= TOL
    # relaxation: the order can change between matvecs (GMRES_Stokes.hpp sets it per iteration)
    plan.kernel().set_p(max(1, P - 3))
    res2 = plan.execute(q)
    ref2 = O.Oracle(pts, ncrit, 0.5).stokes_execute(q, max(1, P - 3), stresslet)
    assert O.rel_l2(res2, ref2) <= TOL
    plan.kernel().set_p(P)
    assert np.array_equal(plan.execute(q), res)


def test_linearity_and_accuracy_c4_sample():
    """Config C4 shape (stresslet, P = 8, ncrit = 64) at N = 60 000: FMM vs brute force on a target sample
    agrees to the reference's own accuracy (2e-4 .. 5e-4 per component), and the matvec is linear."""
    n, P = 60000, 8
    pts, _ = O.drand48_inputs(n)
    rng = np.random.default_rng(5)
    qa = np.hstack([rng.random((n, 3)), np.tile([1.0, 0.0, 0.0], (n, 1))])
    qb = np.hstack([rng.random((n, 3)) - 0.5, np.tile([1.0, 0.0, 0.0], (n, 1))])
    plan = make_plan(pts, P, True)
    ra, rb = plan.execute(qa), plan.execute(qb)
    # linear in g for fixed n: A(g_a + 2 g_b) = A g_a + 2 A g_b
    qc = qa.copy()
    qc[:, :3] = qa[:, :3] + 2 * qb[:, :3]
    assert O.rel_l2(plan.execute(qc), ra + 2 * rb) <= 1e-12
    d = F.Direct.matvec(plan, qa, pts[:500])
    assert O.rel_l2(ra[:500], d) < 2e-3


def test_edge_cases():
    # one leaf only: pure near field, self term excluded
    pts = np.array([[0.25, 0.5, 0.75], [0.3, 0.1, 0.2], [0.31, 0.1, 0.2]])
    q = np.array([[1.0, 2.0, 3.0], [0.5, 0.0, -1.0], [0.0, 1.0, 0.0]])
    plan = make_plan(pts, 4, False)
    res = plan.execute(q)
    assert O.rel_l2(res, O.stokes_direct(pts, q, pts, False)) <= 1e-13
    with pytest.raises(ValueError):
        plan.execute(np.ones(3))            # a Stokeslet charge has three components

"""GPU suite (-m gpu): the CUDA engine, called through the C ABI, against the oracle and the
golden fixtures of the reference.

Tolerances.  Tree, permutation, box table, box geometry and interaction lists: BIT-EXACT
(integer and double fields).  Floating-point results: relative L2 <= 1e-10 as BASELINE.json's
north_star states (potential and force separately); in practice the engine agrees to ~1e-15.
"""
import json

import numpy as np
import pytest

import oracle_lib as O
import fmm_bem_relaxed_b200 as F

pytestmark = pytest.mark.gpu

TOL = 1e-10
TREE_KEYS = ("perm", "codes", "boxes", "geom", "lr", "p2p_off", "p2p_idx")


def make_plan(points, P, ncrit=64, theta=0.5):
    opts = F.FMMOptions()
    opts.set_mac_theta(theta)
    opts.set_max_per_box(ncrit)
    return F.FMM_plan(F.LaplaceSpherical(P), points, opts)


def assert_parity(res, ref, tol=TOL):
    assert O.rel_l2(res[:, 0], ref[:, 0]) <= tol
    assert O.rel_l2(res[:, 1:], ref[:, 1:]) <= tol


@pytest.mark.parametrize("which", ["golden_drand48", "golden_two_scale"])
def test_golden_fixture_tree_bit_exact_and_results(which, request):
    g = request.getfixturevalue(which)
    m = json.loads(str(g["meta"]))
    plan = make_plan(g["points"], m["P"], m["ncrit"], m["theta"])
    t = plan.tree()
    for k in TREE_KEYS:
        assert t[k].shape == g[k].shape, k
        assert np.array_equal(t[k], g[k]), k
    i = plan.info()
    assert (i.n_boxes, i.n_levels, i.n_m2l_pairs, i.n_p2p_box_pairs) == (
        m["boxes"], m["levels"], m["lr_pairs"], m["p2p_pairs"])
    res = plan.execute(g["charges"])
    assert_parity(res, g["results"])
    M, L = plan.expansions()
    used = np.abs(g["M"]).sum(axis=(1, 2)) > 0       # the reference only fills multipoles it needs
    assert O.rel_l2(M[used], g["M"][used]) <= TOL
    assert O.rel_l2(L, g["L"]) <= TOL
    # deterministic: a second execute returns the same bits (target ownership, no atomics)
    assert np.array_equal(plan.execute(g["charges"]), res)


@pytest.mark.parametrize("n,P,ncrit,theta", [(10000, 5, 64, 0.5), (20000, 3, 125, 0.5), (5000, 8, 16, 0.7),
                                             (100000, 5, 64, 0.5)])
def test_uniform_cube_vs_oracle(n, P, ncrit, theta):
    pts, q = O.drand48_inputs(n)
    orc = O.Oracle(pts, ncrit, theta)
    plan = make_plan(pts, P, ncrit, theta)
    ot, gt = orc.tree(), plan.tree()
    for k in TREE_KEYS:
        assert np.array_equal(gt[k], ot[k]), k
    assert_parity(plan.execute(q), orc.execute(q, P, mode=0))


def test_c1_known_answer_checksums(checksums):
    """Config C1 of BASELINE.json: N=100k, P=5 against the reference's own checksums."""
    c = checksums["c1_n100000_p5"]
    pts, q = O.drand48_inputs(100000)
    plan = make_plan(pts, 5)
    res = plan.execute(q)
    w = (np.arange(100000) % 7 + 1).astype(float)
    assert abs(res[:, 0].sum() - c["pot"]) <= TOL * abs(c["pot"])
    assert abs((res[:, 1] * w).sum() - c["fxw"]) <= 1e-8 * abs(c["fxw"])   # signed sum: looser
    assert np.allclose(res[0], c["r0"], rtol=1e-11, atol=0)
    d = F.Direct.matvec(plan, q, pts[:1000])
    assert abs(O.rel_l2(res[:1000, 0], d[:, 0]) - c["err_pot"]) < 1e-8
    assert abs(O.rel_l2(res[:1000, 1:], d[:, 1:]) - c["err_force"]) < 1e-7


def test_per_iteration_p_relaxation():
    """GMRES calls K.set_p(p) before every matvec (reference examples/BEM/GMRES.hpp:195-201)."""
    pts, q = O.drand48_inputs(8000)
    orc = O.Oracle(pts, 64, 0.5)
    plan = make_plan(pts, 8)
    for p in (8, 6, 5, 5, 4, 3, 2, 1, 12, 16, 8):
        plan.kernel().set_p(p)
        assert plan.info().p == p
        assert_parity(plan.execute(q), orc.execute(q, p, mode=0))


def test_adaptive_clustered_cloud():
    rng = np.random.default_rng(3)
    n = 30000
    pts = rng.random((n, 3))
    pts[n // 2:] = 0.3 + 0.05 * rng.random((n - n // 2, 3))
    q = rng.random(n) - 0.3
    orc = O.Oracle(pts, 20, 0.5)
    plan = make_plan(pts, 6, 20, 0.5)
    ot, gt = orc.tree(), plan.tree()
    for k in TREE_KEYS:
        assert np.array_equal(gt[k], ot[k]), k
    assert plan.info().n_levels >= 8
    assert_parity(plan.execute(q), orc.execute(q, 6, mode=0))


def test_sphere_surface_points():
    rng = np.random.default_rng(5)
    v = rng.normal(size=(20000, 3))
    pts = v / np.linalg.norm(v, axis=1)[:, None]
    q = rng.random(20000)
    orc = O.Oracle(pts, 64, 0.5)
    plan = make_plan(pts, 7)
    assert np.array_equal(plan.tree()["lr"], orc.tree()["lr"])
    assert_parity(plan.execute(q), orc.execute(q, 7, mode=0))


def test_direct_sum_matches_oracle():
    pts, q = O.drand48_inputs(3000)
    plan = make_plan(pts, 4)
    tg = np.vstack([pts[:200], np.array([[3.0, 3.0, 3.0]])])
    assert O.rel_l2(F.Direct.matvec(plan, q, tg), O.direct(pts, q, tg)) < 1e-13


def test_edge_cases():
    # fewer bodies than ncrit: the root is a leaf, pure P2P
    pts, q = O.drand48_inputs(50)
    plan = make_plan(pts, 5)
    assert plan.info().n_boxes == 1
    assert O.rel_l2(plan.execute(q), O.direct(pts, q, pts)) < 1e-13
    # one body: zero (self interaction excluded, LaplaceSpherical.hpp:158)
    one = make_plan(np.array([[0.2, 0.4, 0.6]]), 5)
    assert np.all(one.execute(np.array([3.0])) == 0)
    # coincident bodies are dropped from each other's sums like the reference does
    pts2 = np.vstack([pts, pts[:5]])
    q2 = np.concatenate([q, q[:5]])
    assert_parity(make_plan(pts2, 5).execute(q2), O.Oracle(pts2, 64, 0.5).execute(q2, 5, mode=0))
    # ncrit = 1 (deep tree) and a huge ncrit
    pts3, q3 = O.drand48_inputs(600)
    for ncrit in (1, 100000):
        assert_parity(make_plan(pts3, 4, ncrit).execute(q3), O.Oracle(pts3, ncrit, 0.5).execute(q3, 4, mode=0))
    # zero and negative charges
    assert np.all(make_plan(pts3, 4).execute(np.zeros(600)) == 0)
    # wrong charge count
    with pytest.raises(ValueError):
        plan.execute(np.ones(7))


def test_tree_depth_limit_is_an_error_not_a_hang():
    cl = np.full((100, 3), 0.5) + 1e-9 * np.arange(300).reshape(100, 3)
    cl = np.vstack([cl, [[0, 0, 0], [1, 1, 1]]])
    with pytest.raises(F.FmmbError) as e:
        make_plan(cl, 4, ncrit=8)
    assert e.value.status == -3


def test_metric_config_properties_n1m_p8(checksums):
    """BASELINE.json metric config (N=1M, P=8): size-independent properties + the reference's
    known-answer checksums when they were generated (make_golden.py --full)."""
    n = 1000000
    pts, q = O.drand48_inputs(n)
    plan = make_plan(pts, 8)
    i = plan.info()
    assert (i.n_boxes, i.n_leaves, i.n_m2l_pairs, i.n_p2p_box_pairs, i.n_p2p_body_pairs) == (
        37449, 32768, 3873584, 935032, 871763628)
    res = plan.execute(q)
    c = checksums.get("cm_n1000000_p8")
    if c:
        assert abs(res[:, 0].sum() - c["pot"]) <= TOL * abs(c["pot"])
        assert np.allclose(res[0], c["r0"], rtol=1e-10, atol=0)
    # accuracy vs brute force on a target sample matches the reference's own figures
    d = F.Direct.matvec(plan, q, pts[:1000])
    assert abs(O.rel_l2(res[:1000, 0], d[:, 0]) - 1.367046e-06) < 1e-9
    assert abs(O.rel_l2(res[:1000, 1:], d[:, 1:]) - 4.666262e-05) < 1e-8
    # linearity: A(2q) = 2 A(q) exactly (scaling by 2 is exact in binary floating point)
    assert np.array_equal(plan.execute(2.0 * q), 2.0 * res)
    # superposition within rounding
    rng = np.random.default_rng(1)
    q2 = rng.random(n) - 0.5
    r2 = plan.execute(q2)
    r12 = plan.execute(q + q2)
    assert O.rel_l2(r12, res + r2) < 1e-12
    # oracle on a subset of targets is too slow at this size; the sampled leaves' near field is
    # covered by the Direct comparison above.


@pytest.mark.parametrize("name,P,ncrit,theta", [("laplace_treecode_n3000_p4", 4, 32, 0.5),
                                                ("laplace_treecode_two_scale_n4000_p6", 6, 12, 0.6)])
def test_treecode_evaluator_golden(name, P, ncrit, theta):
    """-eval TREE (FMMOptions::TREECODE): M2P for every accepted pair instead of M2L / L2L / L2P."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    opts = F.FMMOptions()
    opts.set_mac_theta(theta)
    opts.set_max_per_box(ncrit)
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.LaplaceSpherical(P), g["points"], opts)
    res = plan.execute(g["charges"])
    assert_parity(res, g["results"])
    assert np.array_equal(plan.execute(g["charges"]), res)
    assert np.array_equal(plan.execute(g["charges"]), res)        # graph replay


def test_treecode_vs_oracle_and_fmm():
    n, P = 30000, 7
    pts, q = O.drand48_inputs(n)
    opts = F.FMMOptions()
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts)
    res = plan.execute(q)
    assert_parity(res, O.Oracle(pts, 64, 0.5).execute(q, P, mode=2))
    # treecode and FMM approximate the same sum: they agree to the truncation error of the expansions
    fmm = make_plan(pts, P).execute(q)
    assert O.rel_l2(res[:, 0], fmm[:, 0]) < 1e-4
    # (the other kernel classes have the treecode path too: tests/test_zz_*.py)


def _near_field_numpy(t, pts, q, self_only):
    """Sum of the Laplace pair kernel over the near-field lists of the tree (what EvalLocal evaluates)."""
    boxes, perm = t["boxes"], t["perm"].astype(np.int64)
    out = np.zeros((len(pts), 4))
    for b in np.nonzero(boxes[:, 7])[0]:
        tid = perm[boxes[b, 4]:boxes[b, 5]]
        srcs = [b] if self_only else t["p2p_idx"][t["p2p_off"][b]:t["p2p_off"][b + 1]]
        sid = np.concatenate([perm[boxes[s, 4]:boxes[s, 5]] for s in srcs])
        d = pts[sid][None, :, :] - pts[tid][:, None, :]
        r2 = (d * d).sum(-1)
        inv = np.where(r2 < 1e-8, 0.0, 1.0 / np.sqrt(np.where(r2 < 1e-8, 1.0, r2)))
        out[tid, 0] = (inv * q[sid]).sum(1)
        out[tid, 1:] = (d * (inv ** 3 * q[sid])[:, :, None]).sum(1)
    return out


def test_near_field_only_plans_for_preconditioners():
    """FMMOptions::local_evaluation (EvalLocal.hpp) and block_diagonal (EvalDiagonalSparse.hpp): the plans the
    reference's LocalPC / BlockDiagonalPC preconditioners are built on."""
    n = 6000
    pts, q = O.drand48_inputs(n)
    t = O.Oracle(pts, 24, 0.5).tree()
    for attr, self_only in (("local_evaluation", False), ("block_diagonal", True)):
        opts = F.FMMOptions()
        opts.set_max_per_box(24)
        setattr(opts, attr, True)
        plan = F.FMM_plan(F.LaplaceSpherical(5), pts, opts)
        res = plan.execute(q)
        assert O.rel_l2(res, _near_field_numpy(t, pts, q, self_only)) <= 1e-12
        assert np.array_equal(plan.execute(q), res)
    # BEM kernels: the cached near-field matrix alone
    v = O.unit_sphere(5)
    qq = np.random.default_rng(2).random(len(v))
    opts = F.FMMOptions()
    opts.local_evaluation = True
    near = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(v), opts).execute(qq)
    opts = F.FMMOptions()
    opts.set_mac_theta(1e-3)                     # nothing accepted: the full matrix is near field
    opts.set_max_per_box(4096)
    full = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(v), opts).execute(qq)
    fmm = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(v)).execute(qq)
    assert O.rel_l2(fmm, full) < 1e-4 and 0.05 < O.rel_l2(near, full) < 0.9     # the near field is a part of it
    opts = F.FMMOptions()
    opts.block_diagonal = True
    diag = F.FMM_plan(F.LaplaceSphericalBEM(6, 4), F.Panels(v), opts)
    assert diag.info().n_p2p_box_pairs == diag.info().n_leaves
    assert np.isfinite(diag.execute(qq)).all()
    with pytest.raises(F.FmmbError):
        F.FMM_plan(F.StokesSpherical(4), pts, opts)


def _plan_mode(points, P, mode, ncrit=64, theta=0.5):
    opts = F.FMMOptions()
    opts.set_mac_theta(theta)
    opts.set_max_per_box(ncrit)
    opts.m2l_mode = mode
    return F.FMM_plan(F.LaplaceSpherical(P), points, opts)


@pytest.mark.parametrize("which", ["golden_drand48", "golden_two_scale"])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_every_far_field_engine_reproduces_the_reference_expansions(which, mode, request):
    """m2l_mode 1 (per-pair kernels), 2 (class-major DMMA GEMM + column reduction) and 3 (output-stationary fused
    sweep, csrc/trans_blocked.cu) against the multipole / local expansions and results the reference itself dumped
    (uniform and strongly adaptive fixture)."""
    g = request.getfixturevalue(which)
    m = json.loads(str(g["meta"]))
    plan = _plan_mode(g["points"], m["P"], mode, m["ncrit"], m["theta"])
    res = plan.execute(g["charges"])
    assert_parity(res, g["results"])
    M, L = plan.expansions()
    used = np.abs(g["M"]).sum(axis=(1, 2)) > 0
    assert O.rel_l2(M[used], g["M"][used]) <= TOL
    assert O.rel_l2(L, g["L"]) <= TOL
    assert np.array_equal(plan.execute(g["charges"]), res)       # fixed summation order in every engine


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 16])
def test_fused_sweep_engine_all_orders_vs_oracle(P):
    """The blocked engine at every order 1..16 (orders 9..16: k-chunks of 64, row blocks of 64) on an adaptive
    cloud, against the oracle; auto mode (class-major GEMM up to 8, fused sweep above) must agree too."""
    rng = np.random.default_rng(P)
    n = 6000
    pts = rng.random((n, 3))
    pts[n // 2:] = 0.6 + 0.1 * rng.random((n - n // 2, 3))        # dense sub-cube: leaves on several levels
    q = rng.random(n) - 0.4
    ref = O.Oracle(pts, 32, 0.5).execute(q, P, mode=0)
    res = _plan_mode(pts, P, 3, 32).execute(q)
    assert_parity(res, ref)
    assert_parity(_plan_mode(pts, P, 0, 32).execute(q), ref)


def test_order_changes_do_not_replay_stale_graphs():
    """graphs are cached per (order, pointers); growing the order reallocates the expansion buffers
    the cached graphs point into.  p = 5, 5, 5 (graph captured and replayed), 12 (buffers grow, other engine), 5
    (must NOT replay the stale graph), 12, 12, 12 (captured at the new order), 5 -- every result against the oracle."""
    n = 20000
    pts, q = O.drand48_inputs(n)
    orc = O.Oracle(pts, 64, 0.5)
    ref = {p: orc.execute(q, p, mode=0) for p in (5, 12)}
    plan = make_plan(pts, 5)
    for p in (5, 5, 5, 12, 5, 5, 12, 12, 12, 5, 5):
        plan.kernel().set_p(p)
        assert_parity(plan.execute(q), ref[p])


@pytest.mark.parametrize("mode,orders", [(3, (6, 6, 6, 8, 8, 8, 6, 6, 4, 4, 4, 8, 6)),
                                         (0, (12, 12, 12, 16, 16, 16, 12, 12, 5, 5, 5, 12, 16))])
def test_replayed_graphs_see_the_expansion_layout_of_their_order(mode, orders):
    """The expansion arrays hold one order's layout at a time (row stride P^2, plus the all-zero row the fused sweep
    reads for absent pairs), and re-laying them out is host-driven work outside the captured launches.  A relaxed
    GMRES revisits orders, so a graph captured at order a is replayed after order b has overwritten the place of a's
    zero row (found with config C2 under m2l_mode 3: the second solve of a plan took 26 iterations instead of 16).
    Third and later calls of an order are replays; every result against the oracle."""
    n = 8000
    pts, q = O.drand48_inputs(n)
    orc = O.Oracle(pts, 64, 0.5)
    ref = {p: orc.execute(q, p, mode=0) for p in set(orders)}
    plan = _plan_mode(pts, orders[0], mode, 64)
    for p in orders:
        plan.kernel().set_p(p)
        assert_parity(plan.execute(q), ref[p])


def test_blocked_batches_pass_their_host_side_audit(monkeypatch):
    """compute-sanitizer is closed on this GPU pool, so the plan-time structures of the fused sweep engine carry their
    own audit (FMMB_SELF_CHECK=1, csrc/trans_blocked.cu::check_blk_batch): every index the kernel dereferences is in
    range, launch orders are permutations, and the items hold exactly the input pairs -- on a uniform and on a strongly
    adaptive tree, single-rank and as one of three ranks."""
    monkeypatch.setenv("FMMB_SELF_CHECK", "1")
    rng = np.random.default_rng(9)
    n = 30000
    uni = rng.random((n, 3))
    ada = rng.random((n, 3))
    ada[n // 3:] = 0.2 + 0.02 * rng.random((n - n // 3, 3))
    for pts in (uni, ada):
        for rank, nranks in ((0, 1), (1, 3)):
            opts = F.FMMOptions()
            opts.m2l_mode = 3
            opts.rank, opts.nranks = rank, nranks
            plan = F.FMM_plan(F.LaplaceSpherical(4), pts, opts)      # build_blk_batch audits M2L, M2M, L2L (+ own / strad)
            plan.execute(rng.random(n))
            plan.close()


@pytest.mark.parametrize("P", [1, 2, 3, 5, 7, 8])
def test_overlap_options_change_the_schedule_not_the_bits(P):
    """Round 2: the far-field chain and the near field overlap inside one CUDA graph.  What decides HOW
    they overlap -- per-node priorities in the graph, the column reduction staged through shared memory by TMA bulk
    copies (m2l_reduce 1) or with a block per box (0), one to three reduction blocks per SM, the near field started
    with the upward pass or behind the GEMM (p2p_order) -- must not change a single bit of the result: the reduction
    adds its columns in the same order in both kernels (odd orders: the padding double of an expansion stays zero).
    Each setting: first call plain launches, second call captures, third replays; all against the oracle."""
    rng = np.random.default_rng(100 + P)
    n = 12000
    pts = rng.random((n, 3))
    pts[n // 2:] = 0.3 + 0.1 * rng.random((n - n // 2, 3))        # adaptive: boxes with and without M2L pairs
    q = rng.random(n) - 0.3
    ref = O.Oracle(pts, 48, 0.5).execute(q, P, mode=0)
    plan = make_plan(pts, P, 48)
    first = None
    for opts in ({}, {"m2l_reduce": 0}, {"m2l_reduce_bps": 1}, {"m2l_reduce_bps": 3}, {"graph_node_priority": 0},
                 {"p2p_order": 1}, {"p2p_order": 1, "m2l_reduce": 0}, {"overlap_p2p": 0}):
        for k, v in {"m2l_reduce": 1, "m2l_reduce_bps": 2, "graph_node_priority": 1, "p2p_order": 0, "overlap_p2p": 1,
                     **opts}.items():
            plan.set_option(k, v)
        for _ in range(3):
            res = plan.execute(q)
            if first is None:
                first = res.copy()
                assert_parity(res, ref)
            assert np.array_equal(res, first), opts

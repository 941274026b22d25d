"""GPU suite (-m gpu): the higher triangle Gauss rules of the reference's table (examples/BEM/GaussQuadrature.hpp:62-276,
keys 13, 19, 25, 79) as panel rules of LaplaceSphericalBEM, against the golden fixtures of the unmodified reference
(tests/golden/laplace_bem_2048_*_k13_*.npz, *_k25_*.npz) and the oracle restatement (bit-identical to them,
tests/test_oracle.py).  The rule tables themselves are pinned on the CPU (tests/test_host_logic.py, kernel 7).

Also here: the treecode evaluator of YukawaCartesian[BEM] (yk_m2p_kernel, yk_bem_m2p_kernel in csrc/yukawa.cu) and of LaplaceSphericalBEM (`LaplaceBEM -eval TREE`, bem_m2p_kernel in csrc/bem.cu) against
the golden fixtures of `ref_bem -tree` and the oracle (bit-identical to them).

STATUS: green on hardware since the round-1 driver run (GPUTEST_r01.json); no xfail mask.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import fmm_bem_relaxed_b200 as F
from conftest import GOLDEN

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("name", ["laplace_bem_2048_p6_k13_bc0", "laplace_bem_2048_p6_k13_bc1", "laplace_bem_2048_p8_k25_bc0"])
def test_golden_fixtures_of_the_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    m = json.loads(str(g["meta"]))
    opts = F.FMMOptions()
    opts.set_max_per_box(m["ncrit"])
    opts.set_mac_theta(m["theta"])
    plan = F.FMM_plan(F.LaplaceSphericalBEM(m["P"], m["K"]), F.Panels(g["verts"], m["bc"]), opts)
    assert O.rel_l2(plan.execute(g["charges"]), g["results"]) <= 1e-10


@pytest.mark.parametrize("K", [13, 19, 25])
def test_vs_oracle(K):
    v = O.unit_sphere(5)
    q = np.random.default_rng(K).random(len(v)) - 0.3
    bc = (np.arange(len(v)) % 2).astype(np.int32)
    opts = F.FMMOptions()
    opts.set_max_per_box(30)
    plan = F.FMM_plan(F.LaplaceSphericalBEM(7, K), F.Panels(v, bc), opts)
    assert O.rel_l2(plan.execute(q), O.BemOracle(v, bc, ncrit=30).execute(q, 7, K)) <= 1e-10


@pytest.mark.parametrize("bc", [0, 1])
def test_treecode_golden_fixtures(bc):
    g = dict(np.load(os.path.join(GOLDEN, "laplace_bem_tree_2048_p6_k4_bc%d.npz" % bc)))
    m = json.loads(str(g["meta"]))
    opts = F.FMMOptions()
    opts.set_max_per_box(m["ncrit"])
    opts.set_mac_theta(m["theta"])
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.LaplaceSphericalBEM(m["P"], m["K"]), F.Panels(g["verts"], bc), opts)
    res = plan.execute(g["charges"])
    assert O.rel_l2(res, g["results"]) <= 1e-10
    assert np.array_equal(plan.execute(g["charges"]), res)


def test_treecode_mixed_boundary_conditions_and_orders():
    v = O.unit_sphere(6)
    bc = (np.arange(len(v)) % 2).astype(np.int32)
    q = np.random.default_rng(8).random(len(v)) - 0.3
    orc = O.BemOracle(v, bc, ncrit=50)
    opts = F.FMMOptions()
    opts.set_max_per_box(50)
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.LaplaceSphericalBEM(8, 4), F.Panels(v, bc), opts)
    for p in (8, 3, 12):
        plan.kernel().set_p(p)
        assert O.rel_l2(plan.execute(q), orc.execute(q, p, 4, treecode=True)) <= 1e-10


@pytest.mark.parametrize("bc", [0, 1])
def test_yukawa_bem_treecode_golden_fixtures(bc):
    """BASELINE config 3's kernel class through the evaluator of the reference that works for it: the treecode
    (tests/golden/yukawa_bem_tree_2048_p6_bc*.npz from oracle/_ref/ref_yukawa_bem -tree; the reference's FMM evaluator
    returns garbage for this class).  This pins the GPU far field of YukawaCartesianBEM to the reference."""
    g = dict(np.load(os.path.join(GOLDEN, "yukawa_bem_tree_2048_p6_bc%d.npz" % bc)))
    opts = F.FMMOptions()
    opts.set_max_per_box(32)
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.YukawaCartesianBEM(6, 1.0, 4), F.Panels(g["verts"], bc), opts)
    res = plan.execute(g["charges"])
    assert O.rel_l2(res, g["results"]) <= 1e-10
    assert O.rel_l2(res, g["direct"]) < (1e-4 if bc == 0 else 1e-3)
    # and the FMM evaluator of the engine approximates the same sum
    opts.evaluator = F.FMMOptions.FMM
    fmm = F.FMM_plan(F.YukawaCartesianBEM(6, 1.0, 4), F.Panels(g["verts"], bc), opts).execute(g["charges"])
    assert O.rel_l2(fmm, res) < (2e-4 if bc == 0 else 2e-3)


def test_yukawa_bem_treecode_at_the_size_of_config_c3():
    """BASELINE config 3 at full size (32 768 panels, kappa = 1, p = 8, k = 4) through the evaluator of the reference that
    works for this kernel class: tests/golden/yukawa_bem_tree_c3_32768_p8_bc0.npz holds the charges and the results of
    oracle/_ref/ref_yukawa_bem -recursions 7 -P 8 -K 4 -kappa 1 -tree (one thread, 14 s)."""
    g = dict(np.load(os.path.join(GOLDEN, "yukawa_bem_tree_c3_32768_p8_bc0.npz")))
    m = json.loads(str(g["meta"]))
    verts = O.unit_sphere(m["recursions"])
    opts = F.FMMOptions()
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.YukawaCartesianBEM(m["P"], m["kappa"], m["K"]), F.Panels(verts, 0), opts)
    res = plan.execute(g["charges"])
    assert O.rel_l2(res, g["results"]) <= 1e-10
    assert abs(res.sum() - m["sum"]) <= 1e-11 * m["sum"]


def test_yukawa_bem_driver_with_the_treecode_evaluator(tmp_path):
    """hostcxx/examples/yukawa_bem.cpp with `-eval TREE` (get_options -> FMMOptions::TREECODE -> fmmb_options.evaluator):
    the GPU matvec against Direct::matvec of the host kernel class, and a relaxed device-resident GMRES solve."""
    import re
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin", "yukawa_bem")
    if not os.path.exists(exe):
        pytest.skip(exe + " not built")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.check_output([exe, "-recursions", "6", "-p", "8", "-k", "4", "-kappa", "1", "-solver_tol", "1e-6", "-check",
                                   "200", "-eval", "TREE"], env=env, timeout=600, cwd=str(tmp_path)).decode()
    assert float(re.search(r"matvec vs Direct \(first 200 rows\): ([0-9.eE+-]+)", out).group(1)) < 1e-4
    m = re.search(r"iterations: (\d+), final residual: ([0-9.eE+-]+), relative error of the solution: ([0-9.eE+-]+)", out)
    assert m and int(m.group(1)) < 60 and float(m.group(2)) < 1e-6 and float(m.group(3)) < 5e-3, out


def test_yukawa_point_kernel_treecode_golden_fixture():
    """YukawaCartesian with FMMOptions::TREECODE (yk_m2p_kernel) against ref_yukawa -tree: potential and gradient."""
    g = dict(np.load(os.path.join(GOLDEN, "yukawa_tree_n3000_p5.npz")))
    m = json.loads(str(g["meta"]))
    opts = F.FMMOptions()
    opts.set_max_per_box(m["ncrit"])
    opts.set_mac_theta(m["theta"])
    opts.evaluator = F.FMMOptions.TREECODE
    plan = F.FMM_plan(F.YukawaCartesian(m["P"], m["kappa"]), g["points"], opts)
    res = plan.execute(g["charges"])
    for k in range(4):
        assert O.rel_l2(res[:, k], g["results"][:, k]) <= 1e-10
    orc = O.Oracle(g["points"], m["ncrit"], m["theta"])
    plan.kernel().set_p(8)
    assert O.rel_l2(plan.execute(g["charges"]), orc.yukawa_execute(g["charges"], 8, m["kappa"], treecode=True)) <= 1e-10


def test_reference_laplacebem_driver_with_the_treecode_evaluator(tmp_path):
    """bin/ref_LaplaceBEM (the reference's examples/LaplaceBEM.cpp compiled unchanged) with `-eval TREE`: the lines the
    unmodified reference prints on one thread (oracle/_ref/LaplaceBEM -recursions 5 -p 8 -k 4 -solver_tol 1e-6 -eval TREE)."""
    import re
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin", "ref_LaplaceBEM")
    if not os.path.exists(exe):
        pytest.skip(exe + " not built")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.check_output([exe, "-recursions", "5", "-p", "8", "-k", "4", "-solver_tol", "1e-6", "-eval", "TREE"],
                                  env=env, timeout=600, cwd=str(tmp_path)).decode()
    want = [(1, 2.030e-04, 8), (2, 9.140e-05, 8), (3, 4.232e-05, 7), (4, 1.847e-05, 6), (5, 8.457e-06, 5), (6, 3.965e-06, 4),
            (7, 2.880e-06, 2), (8, 1.142e-06, 2)]
    got = [(int(a), float(b), int(c)) for a, b, c in re.findall(r"it: (\d+), res: ([0-9.eE+-]+), fmm_req_p: (\d+)", out)]
    assert len(got) == len(want), out
    for (i, r, p), (wi, wr, wp) in zip(got, want):
        assert (i, p) == (wi, wp) and abs(r - wr) <= 2e-3 * wr, (got, out)
    m = re.search(r"Final residual: ([0-9.eE+-]+), after (\d+) iterations", out)
    assert m and int(m.group(2)) == 9 and abs(float(m.group(1)) - 3.5761e-07) <= 5e-3 * 3.5761e-07, out
    assert "relative error: 5.302e-03" in out


def test_key_5_is_rejected():
    with pytest.raises(F.FmmbError):
        F.FMM_plan(F.LaplaceSphericalBEM(5, 5), F.Panels(O.unit_sphere(3)))

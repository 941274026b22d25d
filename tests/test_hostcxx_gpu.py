"""GPU suite: the C++ host mirror (fmm_bem_relaxed_b200/hostcxx) drives the engine.

bin/ref_scaling is the REFERENCE's tests/scaling.cpp compiled unchanged against our headers
(built in the build container, where /root/reference exists); bin/laplace_scaling is our own driver.
"""
import os
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "fmm_bem_relaxed_b200", "hostcxx", "bin")


def run(exe, *args):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.skip(path + " not built")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "fmm_bem_relaxed_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    return subprocess.check_output([path] + list(args), env=env, timeout=600).decode()


def test_reference_scaling_driver_unchanged():
    """Reference tests/scaling.cpp: N=10000, P=5, ncrit=125, prints the force error vs Direct.
    The unmodified reference prints 8.130425e-04 for this case (oracle/_ref, 1 thread)."""
    out = run("ref_scaling")
    m = re.search(r"relative error of Laplace force: ([0-9.eE+-]+)", out)
    assert m, out
    assert abs(float(m.group(1)) - 8.130425e-04) < 1e-8
    assert "FMM execution time" in out


def test_own_driver_c1():
    out = run("laplace_scaling", "100000", "5", "64", "1000")
    pot = float(re.search(r"Laplace potential: ([0-9.eE+-]+)", out).group(1))
    force = float(re.search(r"Laplace force: ([0-9.eE+-]+)", out).group(1))
    chk = float(re.search(r"checksum pot ([0-9.eE+-]+)", out).group(1))
    # reference figures for C1 (tests/golden/checksums.json)
    assert abs(pot - 2.038194e-05) < 1e-10
    assert abs(force - 7.702187e-04) < 1e-9
    assert abs(chk - 9432714514.34655) <= 1e-10 * 9432714514.34655


def test_reference_stresslet_driver_unchanged():
    """Reference serialrun_stresslet.cpp (StokesSpherical, STRESSLET) compiled unchanged against hostcxx/:
    N = 20 000, p = 8.  The patched reference prints 2.1859e-04, 5.2482e-06, 5.1619e-04 for this run
    (its sample loop compares body 0 a thousand times, quirk Q1)."""
    out = run("ref_serialrun_stresslet", "-N", "20000", "-p", "8")
    m = re.search(r"Error \(u\) : ([0-9.eE+-]+), \(v\) : ([0-9.eE+-]+), \(w\) : ([0-9.eE+-]+)", out)
    assert m, out
    got = [float(m.group(i)) for i in (1, 2, 3)]
    for g, want in zip(got, (2.1859e-04, 5.2482e-06, 5.1619e-04)):
        assert abs(g - want) <= 2e-4 * want, (got, out)
    assert "Stresslet calculation" in out


def test_yukawa_bem_driver():
    """hostcxx/examples/yukawa_bem.cpp (BASELINE config 3 shape, smaller mesh): the C++ YukawaCartesianBEM mirror through
    FMM_plan, checked against Direct::matvec with the host kernel class, and a relaxed device-resident GMRES solve of a
    manufactured first-kind problem."""
    out = run("yukawa_bem", "-recursions", "6", "-p", "8", "-k", "4", "-kappa", "1", "-solver_tol", "1e-6", "-check", "200")
    assert float(re.search(r"matvec vs Direct \(first 200 rows\): ([0-9.eE+-]+)", out).group(1)) < 1e-4
    m = re.search(r"iterations: (\d+), final residual: ([0-9.eE+-]+), relative error of the solution: ([0-9.eE+-]+)", out)
    assert m, out
    assert int(m.group(1)) < 60 and float(m.group(2)) < 1e-6 and float(m.group(3)) < 5e-3

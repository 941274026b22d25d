"""CPU suite: the C-ABI library loads and exports every symbol include/fmmb.h declares, argument
validation works without a GPU, and the host mirror behaves like the reference's option parser.
No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

import fmm_bem_relaxed_b200 as F
from fmm_bem_relaxed_b200 import capi
from conftest import ROOT, has_gpu


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fmmb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmmb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTS) == names
    assert b"sm_100a" in lib.fmmb_version()


def test_struct_layouts_match_header(tmp_path):
    assert ctypes.sizeof(capi.KernelDesc) == 24
    assert ctypes.sizeof(capi.Options) == 40
    assert ctypes.sizeof(capi.Sources) == 32
    assert ctypes.sizeof(capi.PlanInfo) == 12 * 8 + 4 * 4
    # every field of every structure, as a C compiler lays include/fmmb.h out (the header is plain C)
    pairs = {"fmmb_kernel_desc": capi.KernelDesc, "fmmb_options": capi.Options, "fmmb_sources": capi.Sources,
             "fmmb_plan_info": capi.PlanInfo, "fmmb_solver_options": capi.SolverOptions, "fmmb_gmres_info": capi.GmresInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fmmb.h"', 'int main(void) {']
    for cname, st in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in st._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, st in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert got[(cname, fname)] == getattr(st, fname).offset, (cname, fname)


def test_argument_validation_needs_no_gpu():
    lib = capi.load()
    h = ctypes.c_void_p()
    pts = np.random.rand(10, 3)
    src = capi.Sources(10, capi.ptr(pts), None, None)
    bad_kind = capi.KernelDesc(99, 5, 0.0, 0, 0)    # not a fmmb_kernel_kind
    assert lib.fmmb_plan_create(ctypes.byref(bad_kind), ctypes.byref(src), None, ctypes.byref(h)) == -4
    assert b"LAPLACE" in lib.fmmb_last_error()
    bad_p = capi.KernelDesc(0, 17, 0.0, 0, 0)
    assert lib.fmmb_plan_create(ctypes.byref(bad_p), ctypes.byref(src), None, ctypes.byref(h)) == -1
    empty = capi.Sources(0, None, None, None)
    ok_k = capi.KernelDesc(0, 5, 0.0, 0, 0)
    assert lib.fmmb_plan_create(ctypes.byref(ok_k), ctypes.byref(empty), None, ctypes.byref(h)) == -1
    assert lib.fmmb_plan_set_p(None, 3) == -1
    assert lib.fmmb_plan_execute(None, None, None) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    with pytest.raises(F.FmmbError) as e:
        F.FMM_plan(F.LaplaceSpherical(5), np.random.rand(100, 3))
    assert e.value.status == -5
    t = ctypes.c_double()
    assert capi.load().fmmb_measure_fp64_peak(0, ctypes.byref(t), None) == -5


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "fmm_bem_relaxed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "fmm_oracle" not in text and "oracle_lib" not in text, f
                assert "oracle/" not in text, f


def test_options_mirror_reference_parser():
    o = F.FMMOptions()
    assert (o.theta, o.max_per_box(), o.evaluator, o.lazy_evaluation) == (0.5, 64, F.FMMOptions.FMM, True)
    o = F.get_options(["prog", "-theta", "0.4", "-ncrit", "125", "-eval", "TREE", "-printtree"])
    assert (o.theta, o.max_per_box(), o.evaluator, o.printTree) == (0.4, 125, F.FMMOptions.TREECODE, True)
    k = F.LaplaceSpherical()
    assert k.P == 5
    k.set_p(8)
    assert k.P == 8


def test_every_plan_option_is_documented_in_the_header():
    """fmmb_plan_set_option takes its switches by name: every name the implementation compares against
    (csrc/capi.cu) appears, quoted, in the option list of include/fmmb.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "fmm_bem_relaxed_b200", "csrc", "capi.cu")).read()
    hdr = open(os.path.join(root, "include", "fmmb.h")).read()
    names = set(re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', src))
    assert len(names) >= 15
    missing = sorted(n for n in names if '"%s"' % n not in hdr)
    assert not missing, missing

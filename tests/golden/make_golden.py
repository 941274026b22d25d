"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    make -C oracle ref && python tests/golden/make_golden.py
It runs oracle/_ref/ref_laplace (the reference's own FMM_plan / LaplaceSpherical / Octree /
EvalInteractionLazy compiled against oracle/boost_shim) single-threaded (the reference's
multi-threaded M2L loop has a data race) and stores what it dumps.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_laplace")


def run_ref(n, p, ncrit, theta, infile=None, dump=None, direct=0):
    cmd = [REF, "-N", str(n), "-P", str(p), "-ncrit", str(ncrit), "-theta", repr(theta)]
    if infile:
        cmd += ["-in", infile]
    if dump:
        cmd += ["-dump", dump]
    if direct:
        cmd += ["-direct", str(direct)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.check_output(cmd, env=env).decode()
    line = [l for l in out.splitlines() if l.startswith("REF_JSON")][0]
    return json.loads(line[len("REF_JSON "):])


def case(name, n, p, ncrit, theta, points=None, charges=None):
    with tempfile.TemporaryDirectory() as tmp:
        infile = None
        if points is not None:
            infile = os.path.join(tmp, "in.f64")
            np.concatenate([points.ravel(), charges]).tofile(infile)
        pre = os.path.join(tmp, "d")
        meta = run_ref(n, p, ncrit, theta, infile, pre, direct=min(n, 500))
        nc = p * (p + 1) // 2
        inp = np.fromfile(pre + ".input.f64")
        data = dict(
            meta=json.dumps(meta),
            points=inp[:3 * n].reshape(n, 3), charges=inp[3 * n:],
            results=np.fromfile(pre + ".results.f64").reshape(n, 4),
            perm=np.fromfile(pre + ".perm.u32", np.uint32),
            codes=np.fromfile(pre + ".codes.u32", np.uint32),
            boxes=np.fromfile(pre + ".boxes.u32", np.uint32).reshape(-1, 8),
            geom=np.fromfile(pre + ".geom.f64").reshape(-1, 4),
            lr=np.fromfile(pre + ".lr.i32", np.int32).reshape(-1, 2),
            p2p_off=np.fromfile(pre + ".p2p_off.i32", np.int32),
            p2p_idx=np.fromfile(pre + ".p2p_idx.i32", np.int32),
            p2m=np.fromfile(pre + ".p2m.i32", np.int32),
            m2m=np.fromfile(pre + ".m2m.i32", np.int32).reshape(-1, 2),
            l2l=np.fromfile(pre + ".l2l.i32", np.int32).reshape(-1, 2),
            l2p=np.fromfile(pre + ".l2p.i32", np.int32),
            M=np.fromfile(pre + ".M.f64").reshape(-1, nc, 2),
            L=np.fromfile(pre + ".L.f64").reshape(-1, nc, 2),
        )
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, meta["boxes"], "boxes", meta["lr_pairs"], "M2L pairs")


def treecode_case(name, n, p, ncrit, theta, points=None, charges=None):
    """LaplaceSpherical with FMMOptions::TREECODE (-eval TREE): P2M, M2M, M2P per accepted pair, P2P."""
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [REF, "-N", str(n), "-P", str(p), "-ncrit", str(ncrit), "-theta", repr(theta), "-tree", "-direct", "300"]
        if points is not None:
            infile = os.path.join(tmp, "in.f64")
            np.concatenate([points.ravel(), charges]).tofile(infile)
            cmd += ["-in", infile]
        pre = os.path.join(tmp, "d")
        cmd += ["-dump", pre]
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1")).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        inp = np.fromfile(pre + ".input.f64")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), points=inp[:3 * n].reshape(n, 3),
                            charges=inp[3 * n:], results=np.fromfile(pre + ".results.f64").reshape(n, 4))
        print(name, meta["pot"], "err vs direct", meta["err_pot"], meta["err_force"])


def stokes_case(name, stresslet, n, p, ncrit, theta, points=None, charges=None, direct=300, tree=False):
    """StokesSpherical through oracle/_ref/ref_stokeslet (unmodified reference) or ref_stresslet (the
    reference with the two compile patches of SURVEY.md section 8(c)); stores inputs, results, checksums."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_stresslet" if stresslet else "ref_stokeslet")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-N", str(n), "-P", str(p), "-ncrit", str(ncrit), "-theta", repr(theta), "-direct", str(direct)]
        if tree:
            cmd.append("-tree")
        if points is not None:
            infile = os.path.join(tmp, "in.f64")
            np.concatenate([points.ravel(), charges.ravel()]).tofile(infile)
            cmd += ["-in", infile]
        pre = os.path.join(tmp, "d")
        cmd += ["-dump", pre]
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1")).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        cd = 6 if stresslet else 3
        inp = np.fromfile(pre + ".input.f64")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), points=inp[:3 * n].reshape(n, 3),
                            charges=inp[3 * n:].reshape(n, cd), results=np.fromfile(pre + ".results.f64").reshape(n, 3))
        print(name, meta["sum"], "err vs direct", meta["err_vs_direct"])


def yukawa_case(name, n, p, kappa, ncrit, theta, points=None, charges=None, direct=300, tree=False):
    """YukawaCartesian through oracle/_ref/ref_yukawa: the unmodified reference class behind the arity adapter of
    oracle/ref_yukawa.cpp."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_yukawa")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-N", str(n), "-P", str(p), "-kappa", repr(kappa), "-ncrit", str(ncrit), "-theta", repr(theta),
               "-direct", str(direct)] + (["-tree"] if tree else [])
        if points is not None:
            infile = os.path.join(tmp, "in.f64")
            np.concatenate([points.ravel(), charges.ravel()]).tofile(infile)
            cmd += ["-in", infile]
        pre = os.path.join(tmp, "d")
        cmd += ["-dump", pre]
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1")).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        inp = np.fromfile(pre + ".input.f64")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), points=inp[:3 * n].reshape(n, 3),
                            charges=inp[3 * n:], results=np.fromfile(pre + ".results.f64").reshape(n, 4))
        print(name, meta["pot"], "err vs direct", meta["err_pot"], meta["err_force"])


def yukawa_bem_case(name, recursions, p, k, kappa, ncrit, bc):
    """YukawaCartesianBEM through oracle/_ref/ref_yukawa_bem (unmodified class behind the arity adapter): the TREECODE
    evaluator and Direct::matvec.  (The reference's FMM evaluator is broken for this kernel)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_yukawa_bem")
    with tempfile.TemporaryDirectory() as tmp:
        pre = os.path.join(tmp, "d")
        cmd = [exe, "-recursions", str(recursions), "-P", str(p), "-K", str(k), "-kappa", repr(kappa), "-ncrit", str(ncrit),
               "-bc", str(bc), "-rand", "-sparse", "0", "-tree", "-direct", "-dump", pre]
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=tmp).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        meta["kappa"] = kappa
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta),
                            verts=np.fromfile(pre + ".verts.f64").reshape(-1, 3, 3), charges=np.fromfile(pre + ".charges.f64"),
                            results=np.fromfile(pre + ".results.f64"), direct=np.fromfile(pre + ".direct.f64"))
        print(name, "treecode vs direct", meta["err_vs_direct"])


def stokes_bem_case(name, exe, recursions, p, k, kfine, mu, ncrit, bc, tree=False):
    """StokesSphericalBEM through oracle/_ref/ref_stokes_bem_asis (the unmodified reference) or oracle/_ref/ref_stokes_bem
    (the reference with the dangling `auto dist` of kernel/StokesSphericalBEM.hpp:162,262 materialised, oracle/Makefile):
    FMM matvec with the sparse near field (examples/StokesBEM.cpp:126) and Direct::matvec, random Vec<3> charges."""
    with tempfile.TemporaryDirectory() as tmp:
        pre = os.path.join(tmp, "d")
        cmd = [os.path.join(ROOT, "oracle", "_ref", exe), "-recursions", str(recursions), "-P", str(p), "-K", str(k),
               "-kfine", str(kfine), "-mu", repr(mu), "-ncrit", str(ncrit), "-bc", str(bc), "-rand", "-direct", "-dump", pre] + \
              (["-tree"] if tree else [])
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=tmp).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        meta["as_written"] = 0 if exe.endswith("asis") else 1
        meta["treecode"] = int(tree)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta),
                            verts=np.fromfile(pre + ".verts.f64").reshape(-1, 3, 3),
                            bc=np.fromfile(pre + ".bc.f64").astype(np.int32),
                            charges=np.fromfile(pre + ".charges.f64").reshape(-1, 3),
                            results=np.fromfile(pre + ".results.f64").reshape(-1, 3),
                            direct=np.fromfile(pre + ".direct.f64").reshape(-1, 3))
        print(name, "FMM vs direct", meta["err_vs_direct"])


def laplace_bem_case(name, recursions, p, k, ncrit, bc, tree=False):
    """LaplaceSphericalBEM through oracle/_ref/ref_bem (the unmodified reference class): FMM matvec with the sparse near
    field (examples/LaplaceBEM.cpp:81) and Direct::matvec, random charges."""
    with tempfile.TemporaryDirectory() as tmp:
        pre = os.path.join(tmp, "d")
        cmd = [os.path.join(ROOT, "oracle", "_ref", "ref_bem"), "-recursions", str(recursions), "-P", str(p), "-K", str(k),
               "-ncrit", str(ncrit), "-bc", str(bc), "-rand", "-direct", "-dump", pre] + (["-tree"] if tree else [])
        out = subprocess.check_output(cmd, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=tmp).decode()
        meta = json.loads([l for l in out.splitlines() if l.startswith("REF_JSON")][0][len("REF_JSON "):])
        meta["treecode"] = int(tree)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta),
                            verts=np.fromfile(pre + ".verts.f64").reshape(-1, 3, 3), charges=np.fromfile(pre + ".charges.f64"),
                            results=np.fromfile(pre + ".results.f64"), direct=np.fromfile(pre + ".direct.f64"))
        print(name, "FMM vs direct", meta["err_vs_direct"])


def main():
    if "--yukawa-p12" in sys.argv:
        # round 2: orders above 10 (the GPU engine's 969-term build); the unmodified class takes any P
        yukawa_case("yukawa_drand48_n1500_p12", 1500, 12, 0.5, 40, 0.5)
        return
    if "--yukawa-tree" in sys.argv:
        yukawa_case("yukawa_tree_n3000_p5", 3000, 5, 0.125, 32, 0.5, tree=True)
        return
    if "--stokes-tree" in sys.argv:
        # FMMOptions::TREECODE for the Stokes classes: Stokeslet (unmodified reference), stresslet (patched),
        # StokesSphericalBEM as compiled (`StokesBEM -eval TREE`)
        stokes_case("stokeslet_tree_n3000_p5", False, 3000, 5, 32, 0.5, tree=True)
        stokes_case("stresslet_tree_n3000_p6", True, 3000, 6, 32, 0.5, tree=True)
        for bc in (0, 2):
            stokes_bem_case("stokes_bem_tree_asis_2048_p6_bc%d" % bc, "ref_stokes_bem_asis", 5, 6, 4, 19, 1e-3, 40, bc, tree=True)
        return
    if "--laplace-bem-tree" in sys.argv:
        # `LaplaceBEM -eval TREE`: FMMOptions::TREECODE with the sparse near field
        for bc in (0, 1):
            laplace_bem_case("laplace_bem_tree_2048_p6_k4_bc%d" % bc, 5, 6, 4, 40, bc, tree=True)
        return
    if "--laplace-bem" in sys.argv:
        # LaplaceSphericalBEM: both boundary conditions, the 4-point rule of BASELINE config 2 and the 13- and 25-point
        # rules of the reference's table; 2 048 panels with ncrit 40 so that the far field is exercised
        for bc in (0, 1):
            laplace_bem_case("laplace_bem_2048_p6_k4_bc%d" % bc, 5, 6, 4, 40, bc)
            laplace_bem_case("laplace_bem_2048_p6_k13_bc%d" % bc, 5, 6, 13, 40, bc)
        laplace_bem_case("laplace_bem_2048_p8_k25_bc0", 5, 8, 25, 30, 0)
        return
    if "--stokes-bem" in sys.argv:
        # StokesSphericalBEM: VELOCITY (the solve), TRACTION (the right-hand side) and mixed panels, as compiled and
        # as written; 2 048 panels with ncrit 40 so that the far field is exercised (theta 0.5)
        for bc in (0, 1, 2):
            stokes_bem_case("stokes_bem_asis_2048_p6_bc%d" % bc, "ref_stokes_bem_asis", 5, 6, 4, 19, 1e-3, 40, bc)
            stokes_bem_case("stokes_bem_2048_p6_bc%d" % bc, "ref_stokes_bem", 5, 6, 4, 19, 1e-3, 40, bc)
        stokes_bem_case("stokes_bem_2048_p8_k3_kf25", "ref_stokes_bem", 5, 8, 3, 25, 0.7, 30, 2)
        return
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/ref_laplace missing: run `make -C oracle ref` in the build container")
    # 1. the reference's own test input (drand48 points then charges), small
    case("laplace_drand48_n3000_p4", 3000, 4, 32, 0.5)
    # 2. strongly adaptive two-scale cloud, signed charges, non-default theta
    rng = np.random.default_rng(20261018)
    n = 4000
    pts = rng.random((n, 3))
    pts[n // 2:] = 0.3 + 0.04 * rng.random((n - n // 2, 3))
    q = rng.random(n) - 0.4
    case("laplace_two_scale_n4000_p6", n, 6, 12, 0.6, pts, q)
    treecode_case("laplace_treecode_n3000_p4", 3000, 4, 32, 0.5)
    treecode_case("laplace_treecode_two_scale_n4000_p6", n, 6, 12, 0.6, pts, q)
    # 2b. StokesSpherical: Stokeslet (unmodified reference) and stresslet (patched reference)
    stokes_case("stokeslet_drand48_n3000_p5", False, 3000, 5, 32, 0.5)
    stokes_case("stresslet_drand48_n3000_p6", True, 3000, 6, 32, 0.5)
    g = rng.random((n, 3)) - 0.5
    nr = rng.normal(size=(n, 3))
    nr /= np.linalg.norm(nr, axis=1)[:, None]
    stokes_case("stresslet_two_scale_n4000_p7", True, n, 7, 12, 0.6, pts, np.hstack([g, nr]))
    stokes_case("stokeslet_two_scale_n4000_p4", False, n, 4, 12, 0.6, pts, g)
    # 2c. YukawaCartesian (point kernel) through the arity adapter
    yukawa_case("yukawa_drand48_n3000_p5", 3000, 5, 0.125, 32, 0.5)
    yukawa_case("yukawa_two_scale_n4000_p6", n, 6, 2.0, 12, 0.6, pts, q)
    yukawa_bem_case("yukawa_bem_tree_2048_p6_bc0", 5, 6, 4, 1.0, 32, 0)
    yukawa_bem_case("yukawa_bem_tree_2048_p6_bc1", 5, 6, 4, 1.0, 32, 1)
    # 3. checksums only, larger sizes
    sums = {}
    for key, (nn, pp) in {"n10000_p5": (10000, 5), "c1_n100000_p5": (100000, 5)}.items():
        sums[key] = run_ref(nn, pp, 64, 0.5, direct=1000)
    if "--full" in sys.argv:
        sums["cm_n1000000_p8"] = run_ref(1000000, 8, 64, 0.5, direct=1000)
    path = os.path.join(HERE, "checksums.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(sums)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

"""ctypes wrapper of oracle/libfmm_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libfmm_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "ref_laplace")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def load():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "fmm_oracle.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        L.fmmo_create.restype = vp
        L.fmmo_create.argtypes = [ctypes.c_int, vp, ctypes.c_uint, ctypes.c_double]
        L.fmmo_destroy.argtypes = [vp]
        L.fmmo_destroy.restype = None
        for name in ("fmmo_error", "fmmo_nboxes", "fmmo_nlevels"):
            getattr(L, name).argtypes = [vp]
        for name in ("fmmo_lr_count", "fmmo_p2p_count"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = ctypes.c_long
        L.fmmo_list_count.argtypes = [vp, ctypes.c_int]
        L.fmmo_list_count.restype = ctypes.c_long
        L.fmmo_get_level_offsets.argtypes = [vp, vp]
        L.fmmo_get_bounds.argtypes = [vp, vp, vp]
        L.fmmo_get_perm.argtypes = [vp, vp, vp]
        L.fmmo_get_boxes.argtypes = [vp, vp, vp]
        L.fmmo_get_lr.argtypes = [vp, vp]
        L.fmmo_get_p2p.argtypes = [vp, vp, vp]
        L.fmmo_get_list.argtypes = [vp, ctypes.c_int, vp]
        L.fmmo_laplace_execute.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int]
        L.fmmo_get_expansions.argtypes = [vp, vp, vp]
        L.fmmo_laplace_direct.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.fmmo_drand48_inputs.argtypes = [ctypes.c_int, vp, vp]
        L.fmmo_stokes_execute.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int]
        L.fmmo_stokes_direct.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.fmmo_yukawa_execute.argtypes = [vp, ctypes.c_int, ctypes.c_double, vp, vp, ctypes.c_int, ctypes.c_int]
        L.fmmo_yukawa_direct.argtypes = [ctypes.c_int, vp, ctypes.c_double, vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.fmmo_yukawa_bem_execute.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp, vp,
                                              ctypes.c_int, ctypes.c_int]
        L.fmmo_yukawa_bem_direct.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp, vp, ctypes.c_int]
        L.fmmo_unit_sphere.argtypes = [ctypes.c_int, vp]
        L.fmmo_panel_centers.argtypes = [ctypes.c_int, vp, vp]
        L.fmmo_bem_execute.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
        L.fmmo_bem_direct.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int]
        L.fmmo_stokes_bem_execute.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                              vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
        L.fmmo_stokes_bem_direct.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                             vp, vp, vp, vp, ctypes.c_int]
        L.fmmo_stokes_bem_entries.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                              vp, vp, ctypes.c_int, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def drand48_inputs(n):
    """The reference tests' input: glibc drand48 default state, n points then n charges
    (reference tests/scaling.cpp:29-38 as compiled by g++)."""
    pts = np.zeros((n, 3))
    q = np.zeros(n)
    load().fmmo_drand48_inputs(n, _p(pts), _p(q))
    return pts, q


class Oracle:
    def __init__(self, points, ncrit=64, theta=0.5):
        L = load()
        self.L = L
        self.pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
        self.n = self.pts.shape[0]
        self.h = ctypes.c_void_p(L.fmmo_create(self.n, _p(self.pts), ncrit, theta))
        self.error = L.fmmo_error(self.h)
        self.nboxes = L.fmmo_nboxes(self.h)
        self.nlevels = L.fmmo_nlevels(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmmo_destroy(self.h)
            self.h = None

    def tree(self):
        L, h, n, nb = self.L, self.h, self.n, self.nboxes
        t = {"perm": np.zeros(n, np.uint32), "codes": np.zeros(n, np.uint32),
             "boxes": np.zeros((nb, 8), np.uint32), "geom": np.zeros((nb, 4))}
        L.fmmo_get_perm(h, _p(t["perm"]), _p(t["codes"]))
        L.fmmo_get_boxes(h, _p(t["boxes"]), _p(t["geom"]))
        nlr = L.fmmo_lr_count(h)
        t["lr"] = np.zeros((nlr, 2), np.int32)
        L.fmmo_get_lr(h, _p(t["lr"]))
        np2p = L.fmmo_p2p_count(h)
        t["p2p_off"] = np.zeros(nb + 1, np.int32)
        t["p2p_idx"] = np.zeros(np2p, np.int32)
        L.fmmo_get_p2p(h, _p(t["p2p_off"]), _p(t["p2p_idx"]))
        return t

    def call_list(self, which):
        width = {0: 1, 1: 2, 2: 2, 3: 1}[which]
        c = self.L.fmmo_list_count(self.h, which)
        a = np.zeros(c * width, np.int32)
        self.L.fmmo_get_list(self.h, which, _p(a))
        return a.reshape(-1, width) if width == 2 else a

    def execute(self, charges, P, mode=1, threads=None):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        res = np.zeros((self.n, 4))
        threads = threads or os.cpu_count() or 1
        rc = self.L.fmmo_laplace_execute(self.h, P, _p(q), _p(res), mode, threads)
        if rc != 0:
            raise RuntimeError("oracle execute failed: %d" % rc)
        self.P = P
        return res

    def stokes_execute(self, charges, P, stresslet, threads=None, treecode=False):
        """StokesSpherical matvec: charges (n, 3) Stokeslet or (n, 6) stresslet (g, n); results (n, 3)."""
        cd = 6 if stresslet else 3
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1, cd))
        res = np.zeros((self.n, 3))
        rc = self.L.fmmo_stokes_execute(self.h, P, int(bool(stresslet)), _p(q), _p(res), 2 if treecode else 0,
                                        threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle stokes execute failed: %d" % rc)
        return res

    def yukawa_execute(self, charges, P, kappa, threads=None, treecode=False):
        """YukawaCartesian matvec: results (n, 4) = potential and the three force-like components."""
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        res = np.zeros((self.n, 4))
        rc = self.L.fmmo_yukawa_execute(self.h, P, float(kappa), _p(q), _p(res), 2 if treecode else 0,
                                        threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle yukawa execute failed: %d" % rc)
        return res

    def expansions(self):
        nc = self.P * (self.P + 1) // 2
        M = np.zeros((self.nboxes, nc, 2))
        Lx = np.zeros((self.nboxes, nc, 2))
        self.L.fmmo_get_expansions(self.h, _p(M), _p(Lx))
        return M, Lx


def unit_sphere(recursions):
    """Octahedron-subdivision sphere of the reference (Triangulation::UnitSphere): (n, 3, 3) vertices."""
    L = load()
    n = L.fmmo_unit_sphere(recursions, None)
    v = np.zeros((n, 3, 3))
    L.fmmo_unit_sphere(recursions, _p(v))
    return v


def panel_centers(verts):
    verts = np.ascontiguousarray(np.asarray(verts, dtype=np.float64).reshape(-1, 9))
    c = np.zeros((verts.shape[0], 3))
    load().fmmo_panel_centers(verts.shape[0], _p(verts), _p(c))
    return c


class BemOracle(Oracle):
    """Oracle tree on the panel centres + the restated LaplaceSphericalBEM matvec."""

    def __init__(self, verts, bc, ncrit=64, theta=0.5):
        self.verts = np.ascontiguousarray(np.asarray(verts, dtype=np.float64).reshape(-1, 9))
        n = self.verts.shape[0]
        self.bc = np.ascontiguousarray(np.broadcast_to(bc, (n,)), np.int32)
        super().__init__(panel_centers(self.verts), ncrit, theta)

    def execute(self, charges, P, K=4, threads=None, treecode=False):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        res = np.zeros(self.n)
        rc = self.L.fmmo_bem_execute(self.h, P, K, _p(self.verts), _p(self.bc), _p(q), _p(res),
                                     2 if treecode else 0, threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle BEM execute failed: %d" % rc)
        return res

    def direct(self, charges, K=4, threads=None):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        out = np.zeros(self.n)
        self.L.fmmo_bem_direct(self.n, K, _p(self.verts), _p(self.bc), _p(q), _p(out), threads or os.cpu_count() or 1)
        return out


class YukawaBemOracle(BemOracle):
    """Oracle tree on the panel centres + the restated YukawaCartesianBEM matvec (FMM: parity unpinned, see
    oracle/fmm_oracle.cpp; treecode and direct: pinned against the reference class)."""

    def __init__(self, verts, bc, kappa, ncrit=64, theta=0.5):
        super().__init__(verts, bc, ncrit, theta)
        self.kappa = float(kappa)

    def execute(self, charges, P, K=4, treecode=False, threads=None):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        res = np.zeros(self.n)
        rc = self.L.fmmo_yukawa_bem_execute(self.h, P, K, self.kappa, _p(self.verts), _p(self.bc), _p(q), _p(res),
                                            2 if treecode else 0, threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle Yukawa BEM execute failed: %d" % rc)
        return res

    def direct(self, charges, K=4, threads=None):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        out = np.zeros(self.n)
        self.L.fmmo_yukawa_bem_direct(self.n, K, self.kappa, _p(self.verts), _p(self.bc), _p(q), _p(out),
                                      threads or os.cpu_count() or 1)
        return out


class StokesBemOracle(BemOracle):
    """Oracle tree on the panel centres + the restated StokesSphericalBEM matvec.  as_written: False = the entries the
    unmodified reference computes when compiled (K-point rule for every pair), True = the branches of its source text
    (self terms, fine rule); see oracle/fmm_oracle.cpp."""

    def __init__(self, verts, bc, mu=1e-3, K=4, kfine=19, as_written=False, ncrit=64, theta=0.5):
        super().__init__(verts, bc, ncrit, theta)
        self.mu, self.K, self.kfine, self.as_written = float(mu), int(K), int(kfine), int(bool(as_written))

    def execute(self, charges, P, threads=None, treecode=False):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        res = np.zeros((self.n, 3))
        rc = self.L.fmmo_stokes_bem_execute(self.h, P, self.K, self.kfine, self.mu, self.as_written, _p(self.verts),
                                            _p(self.bc), _p(q), _p(res), 2 if treecode else 0,
                                            threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle Stokes BEM execute failed: %d" % rc)
        return res

    def direct(self, charges, threads=None):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        out = np.zeros((self.n, 3))
        rc = self.L.fmmo_stokes_bem_direct(self.n, self.K, self.kfine, self.mu, self.as_written, _p(self.verts),
                                           _p(self.bc), _p(q), _p(out), threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle Stokes BEM direct failed: %d" % rc)
        return out

    def entries(self, ti, si):
        ti = np.ascontiguousarray(ti, np.int32)
        si = np.ascontiguousarray(si, np.int32)
        out = np.zeros((len(ti), 3, 3))
        rc = self.L.fmmo_stokes_bem_entries(self.n, self.K, self.kfine, self.mu, self.as_written, _p(self.verts),
                                            _p(self.bc), len(ti), _p(ti), _p(si), _p(out))
        if rc != 0:
            raise RuntimeError("oracle Stokes BEM entries failed: %d" % rc)
        return out


def direct(spts, q, tpts, threads=None):
    spts = np.ascontiguousarray(np.asarray(spts, dtype=np.float64).reshape(-1, 3))
    tpts = np.ascontiguousarray(np.asarray(tpts, dtype=np.float64).reshape(-1, 3))
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1))
    out = np.zeros((tpts.shape[0], 4))
    load().fmmo_laplace_direct(spts.shape[0], _p(spts), _p(q), tpts.shape[0], _p(tpts), _p(out),
                               threads or os.cpu_count() or 1)
    return out


def stokes_direct(spts, q, tpts, stresslet, threads=None):
    spts = np.ascontiguousarray(np.asarray(spts, dtype=np.float64).reshape(-1, 3))
    tpts = np.ascontiguousarray(np.asarray(tpts, dtype=np.float64).reshape(-1, 3))
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(spts.shape[0], -1))
    out = np.zeros((tpts.shape[0], 3))
    load().fmmo_stokes_direct(spts.shape[0], _p(spts), int(bool(stresslet)), _p(q), tpts.shape[0], _p(tpts), _p(out),
                              threads or os.cpu_count() or 1)
    return out


def yukawa_direct(spts, q, tpts, kappa, threads=None):
    spts = np.ascontiguousarray(np.asarray(spts, dtype=np.float64).reshape(-1, 3))
    tpts = np.ascontiguousarray(np.asarray(tpts, dtype=np.float64).reshape(-1, 3))
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1))
    out = np.zeros((tpts.shape[0], 4))
    load().fmmo_yukawa_direct(spts.shape[0], _p(spts), float(kappa), _p(q), tpts.shape[0], _p(tpts), _p(out),
                              threads or os.cpu_count() or 1)
    return out


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def coverage_counts(tree, n):
    """UnitKernel idea of the reference's tests/correctness.cpp (:58-78) on a set of lists: for every leaf, the number of
    sources it sees through its near-field list plus the far-field lists of itself and all its ancestors.  A correct
    traversal counts every source exactly once, i.e. returns n for every leaf ("Wrong counts: 0")."""
    boxes = tree["boxes"].astype(np.int64)
    nb = len(boxes)
    count = boxes[:, 5] - boxes[:, 4]
    lr = tree["lr"].astype(np.int64)
    far = np.bincount(lr[:, 1], weights=count[lr[:, 0]], minlength=nb).astype(np.int64) if len(lr) else np.zeros(nb, np.int64)
    parent = boxes[:, 1]
    level = boxes[:, 6]
    for l in range(1, int(level.max()) + 1):           # parents precede children: accumulate level by level
        idx = np.nonzero(level == l)[0]
        far[idx] += far[parent[idx]]
    off = tree["p2p_off"].astype(np.int64)
    tgt = np.repeat(np.arange(nb), np.diff(off))
    near = np.bincount(tgt, weights=count[tree["p2p_idx"].astype(np.int64)], minlength=nb).astype(np.int64)
    leaves = np.nonzero(boxes[:, 7])[0]
    assert count[leaves].sum() == n
    return near[leaves] + far[leaves]

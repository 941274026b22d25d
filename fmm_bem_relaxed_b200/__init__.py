"""fmm_bem_relaxed_b200 -- B200-native FMM matvec engine behind the fmm-bem-relaxed plan API.

Python mirror of the reference's boundary for the hot path (the reference itself is C++;
the C++ mirror lives in fmm_bem_relaxed_b200/hostcxx/):

    reference (C++)                                   here
    ------------------------------------------------  ---------------------------------------
    FMMOptions opts; opts.set_mac_theta(.5);          opts = FMMOptions(); opts.set_mac_theta(.5)
    opts.set_max_per_box(64)   (FMMOptions.hpp)       opts.set_max_per_box(64)
    LaplaceSpherical K(5)      (LaplaceSpherical.hpp) K = LaplaceSpherical(5)
    FMM_plan<K> plan(K, points, opts) (FMM_plan.hpp)  plan = FMM_plan(K, points, opts)
    plan.kernel().set_p(p)                            plan.kernel().set_p(p)
    result = plan.execute(charges)                    result = plan.execute(charges)
    Direct::matvec(K, pts, charges, targets, exact)   exact = Direct.matvec(plan, charges, targets)

All computation happens in libfmmb200.so (CUDA, sm_100a) through the C ABI of include/fmmb.h.
"""
import ctypes
import weakref

import numpy as np

from . import capi
from .capi import FmmbError

__all__ = ["FMMOptions", "LaplaceSpherical", "LaplaceSphericalBEM", "StokesSpherical", "YukawaCartesian", "YukawaCartesianBEM", "StokesSphericalBEM", "SolverOptions", "GMRES", "FGMRES", "Panels", "FMM_plan", "Direct", "FmmbError", "capi", "comm_unique_id",
           "partition_ranges", "get_options", "drand48_inputs"]


class FMMOptions:
    """Mirror of reference include/FMMOptions.hpp:9-76 (fields and setters used on the hot path)."""
    FMM, TREECODE = 0, 1

    def __init__(self):
        self.lazy_evaluation = True
        self.local_evaluation = False
        self.sparse_local = False
        self.block_diagonal = False
        self.evaluator = FMMOptions.FMM
        self.theta = 0.5
        self.NCRIT_ = 64
        self.printTree = False
        self.device = -1
        self.cold_plan = False        # extension: FMMB_FLAG_COLD_PLAN (no warm start at construction)

    def set_mac_theta(self, theta):
        self.theta = float(theta)

    def set_max_per_box(self, ncrit):
        self.NCRIT_ = int(ncrit)

    def max_per_box(self):
        return self.NCRIT_


def drand48_inputs(n):
    """The reference drivers' synthetic input (tests/scaling.cpp:29-38, serialrun.cpp): n points in the unit cube, then n
    charges, from glibc drand48() in its default state -- X' = (0x5DEECE66D X + 0xB) mod 2^48 from X0 = 0, value
    X'/2^48; g++ evaluates the three coordinate draws of a point right to left, so the first draw lands in z.
    Vectorised: the first block of the sequence is stepped, every later block is the block before it pushed through
    the block-sized jump X -> A X + C (mod 2^48; 64-bit wrap-around products keep the low 48 bits exact)."""
    a, c, mask = 0x5DEECE66D, 0xB, (1 << 48) - 1
    total = 4 * int(n)
    B = min(total, 1 << 16)
    first = np.empty(B, dtype=np.uint64)
    x = 0
    for i in range(B):
        x = (a * x + c) & mask
        first[i] = x
    A, C = 1, 0                                # B-fold composition of X -> a X + c
    for _ in range(B):
        A, C = (a * A) & mask, (a * C + c) & mask
    out = np.empty(total, dtype=np.uint64)
    out[:B] = first
    m = np.uint64(mask)
    pos = B
    while pos < total:
        k = min(B, total - pos)
        out[pos:pos + k] = (out[pos - B:pos - B + k] * np.uint64(A) + np.uint64(C)) & m
        pos += k
    v = out.astype(np.float64) / 281474976710656.0
    pts = np.ascontiguousarray(v[:3 * n].reshape(n, 3)[:, ::-1])
    return pts, np.ascontiguousarray(v[3 * n:])


def get_options(argv):
    """Mirror of get_options(argc, argv), reference include/FMMOptions.hpp:78-106."""
    opts = FMMOptions()
    i = 1
    while i < len(argv):
        a = argv[i]
        if a == "-theta":
            i += 1
            opts.set_mac_theta(float(argv[i]))
        elif a == "-eval":
            i += 1
            if argv[i] == "FMM":
                opts.evaluator = FMMOptions.FMM
            elif argv[i] == "TREE":
                opts.evaluator = FMMOptions.TREECODE
            else:
                print('[W]: Unknown evaluator type: "%s"' % argv[i])
        elif a == "-lazy_eval":
            opts.lazy_evaluation = True
        elif a == "-ncrit":
            i += 1
            opts.set_max_per_box(int(argv[i]))
        elif a == "-printtree":
            opts.printTree = True
        i += 1
    return opts


class LaplaceSpherical:
    """Mirror of reference kernel/LaplaceSpherical.hpp:14-128: order P, set_p; 1 charge, 4 results."""
    kind = capi.LAPLACE_SPHERICAL
    dimension = 3
    charge_dim = 1
    result_dim = 4

    def __init__(self, p=5):
        self.P = int(p)
        self._plan = None

    def set_p(self, p):
        self.P = int(p)
        plan = self._plan() if self._plan is not None else None     # weak reference: no plan <-> kernel cycle,
        if plan is not None and plan._h is not None:                 # so a plan is destroyed when it is dropped
            capi.check(capi.load().fmmb_plan_set_p(plan._h, self.P))


class LaplaceSphericalBEM(LaplaceSpherical):
    """Mirror of reference kernel/LaplaceSphericalBEM.hpp:14-157: order P and K Gauss points per panel.
    Sources are panels: an (n, 3, 3) array of vertices (p0, p1, p2) plus one boundary-condition flag per
    panel (Panels.POTENTIAL / Panels.NORMAL_DERIV); charges and results are scalars."""
    kind = capi.LAPLACE_SPHERICAL_BEM
    result_dim = 1

    def __init__(self, p=5, k=3):
        super().__init__(p)
        self.K = int(k)


class StokesSpherical(LaplaceSpherical):
    """Mirror of reference kernel/StokesSpherical.hpp:11-45.  The reference selects the flavour at compile
    time (-DSTRESSLET, Makefile target serialrun_stresslet); here it is a constructor flag:
    stresslet=False: charge (f0, f1, f2), the Stokeslet;  stresslet=True: charge (g0, g1, g2, n0, n1, n2).
    Results are velocities (n, 3)."""
    result_dim = 3

    def __init__(self, p=5, stresslet=False):
        super().__init__(p)
        self.stresslet = bool(stresslet)
        self.kind = capi.STOKES_SPHERICAL_STRESSLET if self.stresslet else capi.STOKES_SPHERICAL
        self.charge_dim = 6 if self.stresslet else 3


class YukawaCartesian(LaplaceSpherical):
    """Mirror of reference kernel/YukawaCartesian.hpp:14-159: K(t,s) = exp(-kappa |t-s|) / |t-s|, order p, screening
    parameter kappa (constructor YukawaCartesian(int p, double kappa = 0.125)); 1 charge, 4 results.
    Orders 1..16.  set_p(p) means "the full order-p expansion" (the reference's own p < P path walks a wrong
    coefficient subset)."""
    kind = capi.YUKAWA_CARTESIAN

    def __init__(self, p=4, kappa=0.125):
        super().__init__(p)
        self.Kappa = float(kappa)


class YukawaCartesianBEM(LaplaceSphericalBEM):
    """Mirror of reference kernel/YukawaCartesianBEM.hpp:8-143: YukawaCartesianBEM(int p, double kappa, unsigned k).
    Panel sources like LaplaceSphericalBEM; scalar charges and results; orders 1..16."""
    kind = capi.YUKAWA_CARTESIAN_BEM

    def __init__(self, p=5, kappa=0.125, k=3):
        super().__init__(p, k)
        self.Kappa = float(kappa)


class StokesSphericalBEM(LaplaceSphericalBEM):
    """Mirror of reference kernel/StokesSphericalBEM.hpp:9-158: StokesSphericalBEM(int p, unsigned k, double mu) and
    set_Kfine(k) (examples/StokesBEM.cpp:211-214).  Panel sources with Panels.VELOCITY / Panels.TRACTION flags; charges
    and results are (n, 3).  near_field_as_written: False (default) = near-field entries as the unmodified reference
    computes them when compiled (K-point rule for every pair), True = as its source text means them (self terms, fine
    rule); see fmm_bem_relaxed_b200/hostcxx/stokes_bem_math.hpp."""
    kind = capi.STOKES_SPHERICAL_BEM
    charge_dim = 3
    result_dim = 3

    def __init__(self, p=5, k=3, mu=1e-3, kfine=25, near_field_as_written=False):
        super().__init__(p, k)
        self.Mu = float(mu)
        self.K_fine = int(kfine)
        self.near_field_as_written = bool(near_field_as_written)

    def set_Kfine(self, k):
        self.K_fine = int(k)


class Panels:
    """A set of triangular panels (the std::vector<Panel> a reference driver builds)."""
    POTENTIAL, NORMAL_DERIV = 0, 1
    VELOCITY, TRACTION = 0, 1          # StokesSphericalBEM::Panel::BoundaryType

    def __init__(self, vertices, bc=None):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float64).reshape(-1, 3, 3))
        n = self.vertices.shape[0]
        self.bc = np.zeros(n, np.int32) if bc is None else np.ascontiguousarray(np.broadcast_to(bc, (n,)), np.int32)

    def switch_BC(self):
        """Panel::switch_BC on every panel (examples/LaplaceBEM.cpp:218-232)."""
        self.bc = (1 - self.bc).astype(np.int32)

    @property
    def centers(self):
        v = self.vertices
        return ((v[:, 0] + v[:, 1]) + v[:, 2]) / 3

    def __len__(self):
        return self.vertices.shape[0]


class FMM_plan:
    """Mirror of reference include/FMM_plan.hpp:15-128 for source == target plans."""

    def __init__(self, kernel, sources, opts=None):
        lib = capi.load()
        opts = opts or FMMOptions()
        self.opts_ = opts
        verts = bc = None
        if isinstance(kernel, LaplaceSphericalBEM):
            if not isinstance(sources, Panels):
                sources = Panels(sources)
            pts = np.ascontiguousarray(sources.centers)
            verts, bc = sources.vertices, np.ascontiguousarray(sources.bc)
            if isinstance(kernel, YukawaCartesianBEM):
                self.K = YukawaCartesianBEM(kernel.P, kernel.Kappa, kernel.K)
            elif isinstance(kernel, StokesSphericalBEM):
                self.K = StokesSphericalBEM(kernel.P, kernel.K, kernel.Mu, kernel.K_fine, kernel.near_field_as_written)
            else:
                self.K = LaplaceSphericalBEM(kernel.P, kernel.K)
            if isinstance(kernel, StokesSphericalBEM):     # the viscosity travels in the kappa slot (include/fmmb.h)
                kd = capi.KernelDesc(kernel.kind, kernel.P, kernel.Mu, kernel.K, kernel.K_fine)
            else:
                kd = capi.KernelDesc(kernel.kind, kernel.P, getattr(kernel, "Kappa", 0.0), kernel.K, 0)
        else:
            pts = np.ascontiguousarray(np.asarray(sources, dtype=np.float64).reshape(-1, 3))
            # the plan owns a COPY of the kernel (FMM_plan.hpp:37)
            if isinstance(kernel, StokesSpherical):
                self.K = StokesSpherical(kernel.P, kernel.stresslet)
            elif isinstance(kernel, YukawaCartesian):
                self.K = YukawaCartesian(kernel.P, kernel.Kappa)
            else:
                self.K = LaplaceSpherical(kernel.P)
            kd = capi.KernelDesc(kernel.kind, kernel.P, getattr(kernel, "Kappa", 0.0), 0, 0)
        self._n = pts.shape[0]
        self._rdim = self.K.result_dim
        self._cdim = self.K.charge_dim
        src = capi.Sources(self._n, capi.ptr(pts), capi.ptr(verts), capi.ptr(bc))
        near_only = 2 if opts.block_diagonal else (1 if opts.local_evaluation else 0)
        flags = capi.FLAG_STOKES_BEM_AS_WRITTEN if getattr(kernel, "near_field_as_written", False) else 0
        if getattr(opts, "cold_plan", False):
            flags |= capi.FLAG_COLD_PLAN
        op = capi.Options(opts.theta, opts.NCRIT_, opts.evaluator, opts.device, int(getattr(opts, "m2l_mode", 0)),
                          getattr(opts, "rank", 0), getattr(opts, "nranks", 1), near_only, flags)
        h = ctypes.c_void_p()
        capi.check(lib.fmmb_plan_create(ctypes.byref(kd), ctypes.byref(src), ctypes.byref(op), ctypes.byref(h)))
        self._h = h
        self._lib = lib
        self.K._plan = weakref.ref(self)

    def __del__(self):
        self.close()

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.fmmb_plan_destroy(h)
            self._h = None

    def kernel(self):
        return self.K

    def options(self):
        return self.opts_

    def execute(self, charges):
        """results = A * charges; (n, 4) array: potential, fx, fy, fz in the caller's order."""
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        if q.shape[0] != self._n * self._cdim:
            raise ValueError("charges.size() != sources.size()")
        out = np.empty((self._n, self._rdim) if self._rdim > 1 else (self._n,), dtype=np.float64)
        capi.check(self._lib.fmmb_plan_execute(self._h, capi.ptr(q), capi.ptr(out)))
        return out

    def execute_device(self, charges_ptr, results_ptr):
        """Device-pointer variant (ints / ctypes pointers), asynchronous on the plan stream."""
        capi.check(self._lib.fmmb_plan_execute_device(self._h, ctypes.c_void_p(charges_ptr),
                                                      ctypes.c_void_p(results_ptr)))

    def execute_sharded(self, charges_own_ptr, results_own_ptr):
        """Sharded matvec: this rank's charge / result slices (device pointers, tree order, owned range)."""
        capi.check(self._lib.fmmb_plan_execute_sharded(self._h, ctypes.c_void_p(charges_own_ptr),
                                                       ctypes.c_void_p(results_own_ptr)))

    def execute_sharded_host(self, charges_own, results_own=None):
        """Sharded matvec with HOST buffers: charges of this rank's bodies (tree order, owned range) in, results of
        this rank's bodies out.  Pass pinned numpy views to keep the copies asynchronous to the host."""
        i = self.info()
        own = i.own_body_end - i.own_body_begin
        q = np.ascontiguousarray(np.asarray(charges_own, dtype=np.float64).reshape(-1))
        if q.shape[0] != own * self._cdim:
            raise ValueError("charges_own.size() != owned bodies * charge_dim")
        if results_own is None:
            results_own = np.empty((own, self._rdim) if self._rdim > 1 else (own,), dtype=np.float64)
        capi.check(self._lib.fmmb_plan_execute_sharded_host(self._h, capi.ptr(q), capi.ptr(results_own)))
        return results_own

    def comm_init(self, unique_id):
        """Join the NCCL communicator of a partitioned plan (unique_id: 128 bytes from comm_unique_id)."""
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        capi.check(self._lib.fmmb_plan_comm_init(self._h, ctypes.cast(buf, ctypes.c_void_p)))

    def peer_export(self):
        """128-byte blob (IPC handle of the multipole array) for the peer-memory multipole exchange."""
        buf = (ctypes.c_ubyte * 128)()
        capi.check(self._lib.fmmb_plan_peer_export(self._h, ctypes.cast(buf, ctypes.c_void_p)))
        return bytes(buf)

    def peer_init(self, blobs):
        """blobs: the peer_export() blobs of all ranks, concatenated in rank order."""
        data = bytes(blobs)
        buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
        capi.check(self._lib.fmmb_plan_peer_init(self._h, ctypes.cast(buf, ctypes.c_void_p)))

    def set_option(self, name, value):
        capi.check(self._lib.fmmb_plan_set_option(self._h, name.encode(), int(value)))

    def sync(self):
        capi.check(self._lib.fmmb_plan_sync(self._h))

    def stream(self):
        return self._lib.fmmb_plan_stream(self._h)

    def info(self):
        info = capi.PlanInfo()
        capi.check(self._lib.fmmb_plan_get_info(self._h, ctypes.byref(info)))
        return info

    def phase_times(self):
        ms = np.zeros(capi.T_COUNT)
        capi.check(self._lib.fmmb_plan_phase_times(self._h, capi.ptr(ms), capi.T_COUNT))
        return {"total": ms[0], "upward": ms[1], "m2l": ms[2], "downward": ms[3], "p2p": ms[4],
                "h2d": ms[5], "d2h": ms[6], "launches": int(ms[7]), "m2l_gemm": ms[8]}

    def tree(self):
        """Copies of the device tree and lists (see include/fmmb.h: fmmb_plan_get_tree)."""
        i = self.info()
        n, nb = i.n_bodies, i.n_boxes
        t = {
            "perm": np.zeros(n, np.uint32), "codes": np.zeros(n, np.uint32),
            "boxes": np.zeros((nb, 8), np.uint32), "geom": np.zeros((nb, 4)),
            "lr": np.zeros((i.n_m2l_pairs, 2), np.int32), "p2p_off": np.zeros(nb + 1, np.int32),
            "p2p_idx": np.zeros(i.n_p2p_box_pairs, np.int32),
        }
        capi.check(self._lib.fmmb_plan_get_tree(self._h, capi.ptr(t["perm"]), capi.ptr(t["codes"]),
                                                capi.ptr(t["boxes"]), capi.ptr(t["geom"]), capi.ptr(t["lr"]),
                                                capi.ptr(t["p2p_off"]), capi.ptr(t["p2p_idx"])))
        return t

    def expansions(self):
        i = self.info()
        nc = i.p * (i.p + 1) // 2
        M = np.zeros((i.n_boxes, nc, 2))
        L = np.zeros((i.n_boxes, nc, 2))
        capi.check(self._lib.fmmb_plan_get_expansions(self._h, capi.ptr(M), capi.ptr(L)))
        return M, L


class SolverOptions:
    """Mirror of reference examples/BEM/SolverOptions.hpp:9-38 (defaults of the default constructor)."""
    BOURAS, SIMONCINI = 0, 1

    def __init__(self, residual=1e-5, max_iters=500, restart=500, max_p=16, variable_p=True, relax_type=0, p_min=5):
        self.residual, self.max_iters, self.restart = float(residual), int(max_iters), int(restart)
        self.max_p, self.variable_p, self.relax_type = int(max_p), bool(variable_p), int(relax_type)
        self.p_min = int(p_min)        # SolverOptions::p_min, read by GMRES_Stokes only (:13, GMRES_Stokes.hpp:229)


def GMRES(plan, x, b, opts, diag=None, output=False):
    """GMRES(plan, x, b, solver_options[, M]) of reference examples/BEM/GMRES.hpp:142-252, device resident
    (fmmb_gmres).  x: initial guess, overwritten with the solution.  Returns a dict with the iteration record."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    d = None if diag is None else np.ascontiguousarray(diag, dtype=np.float64)
    # Vec<3> unknowns (StokesSphericalBEM): the order rule of GMRES_Stokes.hpp:229, max(p_min, predict_p - 1)
    stokes = plan._cdim == 3
    so = capi.SolverOptions(opts.residual, opts.max_iters, opts.restart, opts.max_p, int(opts.variable_p),
                            opts.relax_type, int(output), opts.p_min if stokes else 0, 1 if stokes else 0)
    info = capi.GmresInfo()
    cap = 4096
    ps = np.zeros(cap, np.int32)
    rs = np.zeros(cap)
    capi.check(plan._lib.fmmb_gmres(plan._h, capi.ptr(b), capi.ptr(x), capi.ptr(d), ctypes.byref(so), ctypes.byref(info),
                                    capi.ptr(ps), capi.ptr(rs), cap))
    k = min(cap, info.n_records)
    plan.K.P = info.final_p           # like the reference, the kernel is left at the last relaxed order
    return {"x": x, "iterations": info.iterations, "final_residual": info.final_residual,
            "p_schedule": ps[:k].tolist(), "residuals": rs[:k].tolist()}


def FGMRES(plan, x, b, opts, pc_plan=None, output=False):
    """FGMRES(plan, x, b, solver_options[, M]) of reference examples/BEM/GMRES_Stokes.hpp:297-431, device resident
    (fmmb_fgmres): flexible GMRES, order rule max(5, predict_p) (:375).  pc_plan: None (identity) or a near-field-only
    plan over the same panels (FMMOptions.local_evaluation -> Preconditioners::LocalInnerSolver, .block_diagonal ->
    Preconditioners::BlockDiagonal); the inner solves use those classes' options (LocalPC_Stokes.hpp:53-57)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    so = capi.SolverOptions(opts.residual, opts.max_iters, opts.restart, opts.max_p, int(opts.variable_p),
                            opts.relax_type, int(output), 5, 0)
    inner = capi.SolverOptions(1e-1, 1, 50, opts.max_p, 0, 0, 0, 0, 0)
    info = capi.GmresInfo()
    cap = 4096
    ps = np.zeros(cap, np.int32)
    rs = np.zeros(cap)
    capi.check(plan._lib.fmmb_fgmres(plan._h, pc_plan._h if pc_plan is not None else None,
                                     ctypes.byref(inner) if pc_plan is not None else None, capi.ptr(b), capi.ptr(x),
                                     ctypes.byref(so), ctypes.byref(info), capi.ptr(ps), capi.ptr(rs), cap))
    k = min(cap, info.n_records)
    plan.K.P = info.final_p
    return {"x": x, "iterations": info.iterations, "final_residual": info.final_residual,
            "p_schedule": ps[:k].tolist(), "residuals": rs[:k].tolist()}


def comm_unique_id():
    """128-byte NCCL unique id (create on rank 0, ship to the other ranks)."""
    buf = (ctypes.c_ubyte * 128)()
    capi.check(capi.load().fmmb_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return bytes(buf)


def partition_ranges(weights, nranks):
    """Contiguous ranges of nearly equal total weight (host only)."""
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
    cuts = np.zeros(nranks + 1, np.int64)
    capi.check(capi.load().fmmb_partition_ranges(capi.ptr(w), w.shape[0], nranks, capi.ptr(cuts)))
    return cuts


class Direct:
    """Mirror of Direct::matvec (reference include/Direct.hpp:273-288), evaluated on the GPU."""

    @staticmethod
    def matvec(plan, charges, targets):
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64).reshape(-1))
        if isinstance(targets, Panels):                 # BEM kernel classes: K(target panel, source panel)
            nt = len(targets)
            out = np.empty((nt, plan._rdim) if plan._rdim > 1 else (nt,))
            bc = np.ascontiguousarray(targets.bc, np.int32)
            capi.check(plan._lib.fmmb_plan_direct_panels(plan._h, capi.ptr(q), nt, capi.ptr(targets.vertices), capi.ptr(bc),
                                                         capi.ptr(out)))
            return out
        t = np.ascontiguousarray(np.asarray(targets, dtype=np.float64).reshape(-1, 3))
        out = np.empty((t.shape[0], plan._rdim))
        capi.check(plan._lib.fmmb_plan_direct(plan._h, capi.ptr(q), t.shape[0], capi.ptr(t), capi.ptr(out)))
        return out

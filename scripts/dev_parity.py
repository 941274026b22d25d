"""Development check on a GPU box: CUDA plan vs oracle port, array by array."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
import fmm_bem_relaxed_b200 as F

def run(n, P, ncrit=64, theta=0.5, pts=None, q=None, label=""):
    if pts is None:
        pts, q = O.drand48_inputs(n)
    t = time.time(); orc = O.Oracle(pts, ncrit, theta); ot = orc.tree(); t_or = time.time() - t
    opts = F.FMMOptions(); opts.set_mac_theta(theta); opts.set_max_per_box(ncrit)
    t = time.time(); plan = F.FMM_plan(F.LaplaceSpherical(P), pts, opts); t_pl = time.time() - t
    gt = plan.tree()
    i = plan.info()
    print("   m2l classes %d, batched pairs %d of %d" % (i.n_m2l_classes, i.n_m2l_pairs_batched, i.n_m2l_pairs))
    print("== %s N=%d P=%d ncrit=%d: boxes %d levels %d lr %d p2p %d bodypairs %d | plan %.3fs (oracle tree %.3fs)" % (
        label, n, P, ncrit, i.n_boxes, i.n_levels, i.n_m2l_pairs, i.n_p2p_box_pairs, i.n_p2p_body_pairs, t_pl, t_or))
    ok = True
    for k in ("perm", "codes", "boxes", "geom", "lr", "p2p_off", "p2p_idx"):
        same = gt[k].shape == ot[k].shape and np.array_equal(gt[k], ot[k])
        ok &= same
        print("   tree.%-8s bit-exact: %s" % (k, same))
        if not same and gt[k].shape == ot[k].shape:
            bad = np.argwhere(gt[k] != ot[k])
            print("      first mismatches", bad[:5].tolist(), gt[k][tuple(bad[0])], ot[k][tuple(bad[0])])
    t = time.time(); ref = orc.execute(q, P, mode=1); t_o = time.time() - t
    res = plan.execute(q)
    t = time.time(); res = plan.execute(q); t_g = time.time() - t
    M, L = plan.expansions(); oM, oL = orc.expansions()
    print("   M rel %.3e  L rel %.3e" % (O.rel_l2(M, oM), O.rel_l2(L, oL)))
    print("   result rel L2: pot %.3e force %.3e   (oracle %.2fs, gpu e2e %.4fs) phases %s" % (
        O.rel_l2(res[:, 0], ref[:, 0]), O.rel_l2(res[:, 1:], ref[:, 1:]), t_o, t_g,
        {k: round(v, 3) for k, v in plan.phase_times().items()}))
    m = min(n, 500)
    d = F.Direct.matvec(plan, q, pts[:m]); od = O.direct(pts, q, pts[:m])
    print("   direct gpu vs oracle %.3e ; fmm vs direct pot %.3e force %.3e" % (
        O.rel_l2(d, od), O.rel_l2(res[:m, 0], d[:, 0]), O.rel_l2(res[:m, 1:], d[:, 1:])))
    return ok

if __name__ == "__main__":
    run(10000, 5, label="uniform")
    rng = np.random.default_rng(7)
    N = 30000
    pts = rng.random((N, 3)); pts[N // 2:] = 0.3 + 0.05 * rng.random((N - N // 2, 3))
    run(N, 6, ncrit=20, pts=pts, q=rng.random(N) - 0.3, label="two-scale")
    run(100000, 5, label="C1")
    if len(sys.argv) > 1:
        run(1000000, 8, label="CM")

"""Design prototype (numpy): IPM warm-up + primal-dual active-set finisher with KKT certificate.

NOT product code, NOT the oracle.  Establishes the algorithm the CUDA kernel implements:
  phase 1: Mehrotra predictor-corrector IPM on the swing-eliminated, scaled QP until mu < mu_switch
  phase 2: per-foot-step face states (zx, zy in {-1,0,1}; zz in {-1 vertex, 0 free, 1 at fz_max});
           reduced solve; multipliers; simultaneous add/drop; stop when KKT verified.
"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from proto_ipm import gen_state, build_literal, reduce_qp, cons, ipm, exact_active_set, GAZEBO, HARDWARE, MU, FZMAX


SAFE = True


def amax(v, dv):
    msk = dv < 0
    return min(1.0, (-v[msk] / dv[msk]).min()) if msk.any() else 1.0


def ipm_phase(Hs, gs, C, ds, mu_switch, maxit=30):
    n = len(gs)
    m = len(ds)
    x = np.zeros(n)
    x[2::3] = 0.25 * ds[4]
    s = np.maximum(ds - C @ x, 1e-2)
    lam = np.full(m, 1.0) * (np.abs(gs).max() + 1e-3)
    for it in range(maxit):
        rd = Hs @ x + gs + C.T @ lam
        rp = C @ x + s - ds
        mu = s @ lam / m
        if mu < mu_switch and np.abs(rd).max() < 1e-6 and np.abs(rp).max() < 1e-6:
            return x, s, lam, it
        w = lam / s
        L = np.linalg.cholesky(Hs + C.T @ (w[:, None] * C))

        def solve(rc):
            rhs = -rd + C.T @ (rc / s - w * rp)
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds_ = -rp - C @ dx
            dl = -(rc + lam * ds_) / s
            return dx, ds_, dl
        dxa, dsa, dla = solve(s * lam)
        aa = min(amax(s, dsa), amax(lam, dla))
        mu_aff = (s + aa * dsa) @ (lam + aa * dla) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds_, dl = solve(s * lam + dsa * dla - sigma * mu)
        ap, ad = amax(s, ds_), amax(lam, dl)
        a = min(0.995 * ap if ap < 1 else 1.0, 0.995 * ad if ad < 1 else 1.0)
        x += a * dx
        s += a * ds_
        lam += a * dl
    return x, s, lam, maxit


def guess_faces(s, lam):
    K = len(s) // 5
    zx = np.zeros(K, int)
    zy = np.zeros(K, int)
    zz = np.zeros(K, int)
    for k in range(K):
        a = lam[5 * k:5 * k + 5] > s[5 * k:5 * k + 5]
        # rows: 0: -fx-mu fz<=0 (fx=-mu fz), 1: fx-mu fz<=0 (fx=+mu fz), 2/3 same for y, 4: fz<=max
        if (a[0] and a[1]) or (a[2] and a[3]):
            zz[k] = -1
            continue
        zx[k] = -1 if a[0] else (1 if a[1] else 0)
        zy[k] = -1 if a[2] else (1 if a[3] else 0)
        zz[k] = 1 if a[4] else 0
    return zx, zy, zz


def pdas(Hs, gs, fzmax, zx, zy, zz, maxround=12, tol=1e-11, verbose=False):
    n = len(gs)
    K = n // 3
    for rnd in range(maxround):
        Z = np.zeros((n, n))
        c = np.zeros(n)
        fixed = np.zeros(n, bool)
        for k in range(K):
            ix, iy, iz = 3 * k, 3 * k + 1, 3 * k + 2
            if zz[k] == -1:
                fixed[[ix, iy, iz]] = True
                continue
            if zz[k] == 0:
                Z[iz, iz] = 1
                if zx[k]:
                    Z[ix, iz] = zx[k] * MU
                    fixed[ix] = True
                else:
                    Z[ix, ix] = 1
                if zy[k]:
                    Z[iy, iz] = zy[k] * MU
                    fixed[iy] = True
                else:
                    Z[iy, iy] = 1
            else:
                fixed[iz] = True
                c[iz] = fzmax
                if zx[k]:
                    c[ix] = zx[k] * MU * fzmax
                    fixed[ix] = True
                else:
                    Z[ix, ix] = 1
                if zy[k]:
                    c[iy] = zy[k] * MU * fzmax
                    fixed[iy] = True
                else:
                    Z[iy, iy] = 1
        M = Z.T @ Hs @ Z + np.diag(fixed.astype(float))
        rhs = -Z.T @ (gs + Hs @ c)
        L = np.linalg.cholesky(M)
        y = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
        u = Z @ y + c
        r = -(Hs @ u + gs)
        changed = 0
        # primal violation anywhere?
        pv = False
        for k in range(K):
            fx, fy, fz = u[3 * k:3 * k + 3]
            if zz[k] == 0 and (fz > fzmax + tol or fz < -tol):
                pv = True
            if zz[k] != -1 and ((zx[k] == 0 and abs(fx) > MU * fz + tol) or (zy[k] == 0 and abs(fy) > MU * fz + tol)):
                pv = True
        allow_drop = (not pv) or (not SAFE)
        for k in range(K):
            fx, fy, fz = u[3 * k:3 * k + 3]
            rx, ry, rz = r[3 * k:3 * k + 3]
            if zz[k] == -1:
                if allow_drop and -rz / MU < abs(rx) + abs(ry) - tol:
                    # leave the vertex: free fz, keep the faces the residual points to
                    zz[k] = 0
                    zx[k] = (1 if rx > 0 else -1) if abs(rx) > tol else 0
                    zy[k] = (1 if ry > 0 else -1) if abs(ry) > tol else 0
                    # face multipliers lam_x=|rx|, lam_y=|ry| ; keep only if consistent (they are >=0 by construction)
                    changed += 1
                continue
            lx = zx[k] * rx if zx[k] else 0.0      # multiplier of the active x face
            ly = zy[k] * ry if zy[k] else 0.0
            l5 = rz + MU * (lx + ly)
            newzx, newzy, newzz = zx[k], zy[k], zz[k]
            # dual: drop faces with negative multipliers (only in rounds with no primal violation anywhere)
            if allow_drop:
                if zx[k] and lx < -tol:
                    newzx = 0
                if zy[k] and ly < -tol:
                    newzy = 0
                if zz[k] == 1 and l5 < -tol:
                    newzz = 0
            # primal: add violated faces
            if zz[k] == 0:
                if fz > fzmax + tol:
                    newzz = 1
                elif fz < -tol:
                    newzz = -1
            if newzz != -1:
                if zx[k] == 0 and abs(fx) > MU * fz + tol:
                    newzx = 1 if fx > 0 else -1
                if zy[k] == 0 and abs(fy) > MU * fz + tol:
                    newzy = 1 if fy > 0 else -1
            if (newzx, newzy, newzz) != (zx[k], zy[k], zz[k]):
                changed += 1
                zx[k], zy[k], zz[k] = newzx, newzy, newzz
                if newzz == -1:
                    zx[k] = zy[k] = 0
        if verbose:
            print("  pdas round", rnd, "changed", changed)
        if changed == 0:
            return u, rnd + 1, True
    return u, maxround, False


def solve(H, g, fscale=100.0, mu_switch=1e-7, verbose=False):
    n = len(g)
    C, d = cons(n)
    Hs = H * fscale * fscale
    gs = g * fscale
    cs = np.abs(Hs).max()
    Hs = Hs / cs
    gs = gs / cs
    ds = d / fscale
    x, s, lam, it = ipm_phase(Hs, gs, C, ds, mu_switch)
    zx, zy, zz = guess_faces(s, lam)
    u, rounds, ok = pdas(Hs, gs, FZMAX / fscale, zx, zy, zz, verbose=verbose)
    return u * fscale, it, rounds, ok


if __name__ == "__main__":
    ntest = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    mu_switch = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-7
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rng = np.random.default_rng(11)
    its, rnds, errs, fails = [], [], [], 0
    for t in range(ntest):
        st = gen_state(rng, wide=(t % 2 == 1))
        par = GAZEBO if t % 4 < 2 else HARDWARE
        H, g = build_literal(st, par, N=N)
        Hr, gr, idx = reduce_qp(H, g, st['contacts'], N=N)
        C, d = cons(len(gr))
        u, it, rounds, ok = solve(Hr, gr, mu_switch=mu_switch)
        # certificate check in original units
        r = -(Hr @ u + gr)
        viol = (C @ u - d).max()
        ui, li, nit, hist = ipm(Hr, gr, C, d)
        err = np.abs(u - ui).max()
        its.append(it)
        rnds.append(rounds)
        errs.append(err)
        if not ok:
            fails += 1
        if not ok or err > 1e-2:
            print(t, "S", st['contacts'].sum(), "ipm its", it, "pdas rounds", rounds, "ok", ok, "viol %.1e" % viol, "diff vs ipm-only %.1e" % err)
    its = np.array(its)
    rnds = np.array(rnds)
    print("ipm its p50 %.0f max %d | pdas rounds p50 %.0f p90 %.0f max %d | total p50 %.0f p90 %.0f max %d | fails %d/%d" % (
        np.median(its), its.max(), np.median(rnds), np.percentile(rnds, 90), rnds.max(),
        np.median(its + rnds), np.percentile(its + rnds, 90), (its + rnds).max(), fails, ntest))
    print("diff vs ipm-only: p50 %.1e p99 %.1e max %.1e" % (np.median(errs), np.percentile(errs, 99), max(errs)))

"""Design prototype (numpy): wrench-space reduction for >= 3 stance feet.

With S stance feet the step-s block of B_qp factors through the 6-dim net wrench:  H = V' Hw V,
V = I_N (x) M0 (6 x 3S), Hw = T0 (x) Q0 + T1 (x) Q1' (6N x 6N).  K = D + V' Hw V (D block diagonal 3x3) is
solved with   x = D^-1 (b - V' y),   y = Ls^-T (I + Ls' Hw Ls)^-1 Ls' Hw V D^-1 b,   S = V D^-1 V' = Ls Ls'
so the dense factorisation is 6N x 6N whatever S is.  Checks the accuracy of IPM + finisher built on it.
NOT product code.
"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from proto_ipm import gen_state, build_literal, reduce_qp, cons, ipm, GAZEBO, HARDWARE, MU, FZMAX, skew
from proto_pdas import guess_faces, amax


def wrench_factors(st, par, N=10, dt=0.0025):
    """M0 (6 x 3S) and Hw (6N x 6N), unscaled"""
    yaw = st['euler'][2]
    c, s = np.cos(yaw), np.sin(yaw)
    E = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
    Iw = st['R'] @ par['inertia'] @ st['R'].T
    legs = [i for i in range(4) if st['contacts'][i]]
    M0 = np.zeros((6, 3 * len(legs)))
    for k, i in enumerate(legs):
        M0[0:3, 3 * k:3 * k + 3] = dt * np.linalg.inv(Iw) @ skew(st['foot'][:, i])
        M0[3:6, 3 * k:3 * k + 3] = dt * np.eye(3) / par['mass']
    P = np.zeros((6, 6)); P[0:3, 0:3] = E; P[3:6, 3:6] = np.eye(3)
    Q0 = np.diag(2 * par['q'][6:12]); Q1 = np.diag(2 * par['q'][0:6])
    Q1p = dt * dt * P.T @ Q1 @ P
    T0 = np.array([[N - max(a, b) for b in range(N)] for a in range(N)], float)
    T1 = np.array([[sum((i - a) * (i - b) for i in range(max(a, b), N)) for b in range(N)] for a in range(N)], float)
    Hw = np.kron(T0, Q0) + np.kron(T1, Q1p)
    return M0, Hw


def psd_chol(S):
    """Cholesky that tolerates a singular (rank-deficient) PSD matrix: a vanishing pivot zeroes its column"""
    n = S.shape[0]
    L = np.zeros_like(S)
    A = S.copy()
    scale = max(np.abs(np.diag(S)).max(), 1e-300)
    for j in range(n):
        d = A[j, j]
        if d <= 1e-14 * scale:
            continue
        L[j:, j] = A[j:, j] / np.sqrt(d)
        A[j:, j:] -= np.outer(L[j:, j], L[j:, j])
    return L


REFINE = 1


class WrenchK:
    """solve (D + V' Hw V) x = b with D block-diagonal (list of 3x3), B_k = M0_f Z_k (6x3) per foot-step.
    K^-1 = D^-1 - D^-1 V' [Hw - Hw Ls (I + Ls' Hw Ls)^-1 Ls' Hw] V D^-1,  S = V D^-1 V' = Ls Ls' (Ls may be singular)"""
    def __init__(self, Dblocks, Bblocks, Hw, N, S, matvec):
        self.N, self.S, self.Hw, self.matvec = N, S, Hw, matvec
        self.Dinv = [np.linalg.inv(d) for d in Dblocks]
        self.B = Bblocks
        Lbig = np.zeros((6 * N, 6 * N))
        for s in range(N):
            Ss = sum(Bblocks[s * S + f] @ self.Dinv[s * S + f] @ Bblocks[s * S + f].T for f in range(S))
            Lbig[6 * s:6 * s + 6, 6 * s:6 * s + 6] = psd_chol(Ss)
        self.Lbig = Lbig
        self.Kr = np.eye(6 * N) + Lbig.T @ Hw @ Lbig
        self.Lk = np.linalg.cholesky(self.Kr)

    def solve0(self, b):
        N, S = self.N, self.S
        t = np.concatenate([self.Dinv[k] @ b[3 * k:3 * k + 3] for k in range(N * S)])
        v = np.concatenate([sum(self.B[s * S + f] @ t[3 * (s * S + f):3 * (s * S + f) + 3] for f in range(S)) for s in range(N)])
        hv = self.Hw @ v
        z = self.Lbig.T @ hv
        z = np.linalg.solve(self.Lk.T, np.linalg.solve(self.Lk, z))
        y = hv - self.Hw @ (self.Lbig @ z)
        x = t.copy()
        for s in range(N):
            for f in range(S):
                k = s * S + f
                x[3 * k:3 * k + 3] -= self.Dinv[k] @ (self.B[k].T @ y[6 * s:6 * s + 6])
        return x

    def solve(self, b):
        x = self.solve0(b)
        for _ in range(REFINE):
            r = b - self.matvec(x)
            x = x + self.solve0(r)
        return x


def solve_wrench(st, par, N=10, fscale=100.0, mu_switch=1e-9, dense_check=None):
    S = int(st['contacts'].sum())
    M0, Hw = wrench_factors(st, par, N)
    H, g = build_literal(st, par, N=N)
    Hr, gr, idx = reduce_qp(H, g, st['contacts'], N=N)
    n = len(gr); K = n // 3; m = 5 * K
    C, d = cons(n)
    cs = np.abs(Hr).max() * fscale ** 2
    hs = fscale ** 2 / cs
    Hs = Hr * hs; gs = gr * fscale / cs; ds = d / fscale
    Hws = Hw * hs
    R2 = np.tile(2 * par['r'][[3 * i + a for i in range(4) if st['contacts'][i] for a in range(3)]], N) * hs
    Mf = [M0[:, 3 * f:3 * f + 3] for f in range(S)]
    # consistency of the factorisation itself
    V = np.kron(np.eye(N), M0)
    assert np.abs(V.T @ Hws @ V + np.diag(R2) - Hs).max() < 1e-12 * np.abs(Hs).max()
    x = np.zeros(n); x[2::3] = 0.25 * ds[4]
    s = np.maximum(ds - C @ x, 1e-2); lam = np.full(m, np.abs(gs).max() + 1e-3)
    it = 0
    mu_target = mu_switch
    for attempt in range(3):
        ok = False
        while it < 40:
            rd = Hs @ x + gs + C.T @ lam; rp = C @ x + s - ds; mu = s @ lam / m
            if mu < mu_target and max(np.abs(rd).max(), np.abs(rp).max()) < 1e-6:
                ok = True; break
            w = lam / s
            Db = []
            for k in range(K):
                Ck = C[5 * k:5 * k + 5, 3 * k:3 * k + 3]
                Db.append(np.diag(R2[3 * k:3 * k + 3]) + Ck.T @ (w[5 * k:5 * k + 5, None] * Ck))
            Kd = Hs + C.T @ (w[:, None] * C)
            WK = WrenchK(Db, [Mf[k % S] for k in range(K)], Hws, N, S, lambda v: Kd @ v)

            def solve(rc):
                rhs = -rd + C.T @ (rc / s - w * rp)
                dx = WK.solve0(rhs); ds_ = -rp - C @ dx; dl = -(rc + lam * ds_) / s
                return dx, ds_, dl
            dxa, dsa, dla = solve(s * lam)
            aa = min(amax(s, dsa), amax(lam, dla)); mu_aff = (s + aa * dsa) @ (lam + aa * dla) / m
            dx, ds_, dl = solve(s * lam + dsa * dla - (mu_aff / mu) ** 3 * mu)
            ap, ad = amax(s, ds_), amax(lam, dl)
            a = min(0.995 * ap if ap < 1 else 1.0, 0.995 * ad if ad < 1 else 1.0)
            x += a * dx; s += a * ds_; lam += a * dl; it += 1
        zx, zy, zz = guess_faces(s, lam)
        dmax = FZMAX / fscale
        tol = 1e-11
        for rnd in range(4):
            Db, Bb, cvec = [], [], np.zeros(n)
            Zs = []
            for k in range(K):
                xf = float(zx[k] == 0 and zz[k] != -1); yf = float(zy[k] == 0 and zz[k] != -1); zf = float(zz[k] == 0)
                Z = np.array([[xf, 0, zx[k] * MU * zf], [0, yf, zy[k] * MU * zf], [0, 0, zf]])
                Zs.append(Z)
                Db.append(Z.T @ np.diag(R2[3 * k:3 * k + 3]) @ Z + np.diag([1 - xf, 1 - yf, 1 - zf]))
                Bb.append(Mf[k % S] @ Z)
                cz = dmax if zz[k] == 1 else 0.0
                cvec[3 * k:3 * k + 3] = [zx[k] * MU * cz, zy[k] * MU * cz, cz]
            t = Hs @ cvec + gs
            rhs = np.concatenate([-(Zs[k].T @ t[3 * k:3 * k + 3]) for k in range(K)])
            Zb = np.zeros((n, n))
            for k in range(K):
                Zb[3 * k:3 * k + 3, 3 * k:3 * k + 3] = Zs[k]
            Md = Zb.T @ Hs @ Zb + np.diag([1.0 - Zs[k][a, a] if a < 2 else 1.0 - Zs[k][2, 2] for k in range(K) for a in range(3)])
            WK = WrenchK(Db, Bb, Hws, N, S, lambda v: Md @ v)
            y = WK.solve(rhs)
            u = np.concatenate([Zs[k] @ y[3 * k:3 * k + 3] for k in range(K)]) + cvec
            r = -(Hs @ u + gs)
            pv = False
            for k in range(K):
                fx, fy, fz = u[3 * k:3 * k + 3]
                if zz[k] == 0 and (fz > dmax + tol or fz < -tol): pv = True
                if zz[k] != -1 and ((zx[k] == 0 and abs(fx) > MU * fz + tol) or (zy[k] == 0 and abs(fy) > MU * fz + tol)): pv = True
            changed = 0
            for k in range(K):
                fx, fy, fz = u[3 * k:3 * k + 3]; rx, ry, rz = r[3 * k:3 * k + 3]
                if zz[k] == -1:
                    if not pv and -rz / MU < abs(rx) + abs(ry) - tol:
                        zz[k] = 0; zx[k] = (1 if rx > 0 else -1) if abs(rx) > tol else 0; zy[k] = (1 if ry > 0 else -1) if abs(ry) > tol else 0; changed += 1
                    continue
                lx = zx[k] * rx if zx[k] else 0.0; ly = zy[k] * ry if zy[k] else 0.0; l5 = rz + MU * (lx + ly)
                nzx, nzy, nzz = zx[k], zy[k], zz[k]
                if not pv:
                    if zx[k] and lx < -tol: nzx = 0
                    if zy[k] and ly < -tol: nzy = 0
                    if zz[k] == 1 and l5 < -tol: nzz = 0
                if zz[k] == 0:
                    if fz > dmax + tol: nzz = 1
                    elif fz < -tol: nzz = -1
                if nzz != -1:
                    if zx[k] == 0 and abs(fx) > MU * fz + tol: nzx = 1 if fx > 0 else -1
                    if zy[k] == 0 and abs(fy) > MU * fz + tol: nzy = 1 if fy > 0 else -1
                else:
                    nzx = nzy = 0
                if (nzx, nzy, nzz) != (zx[k], zy[k], zz[k]): changed += 1; zx[k], zy[k], zz[k] = nzx, nzy, nzz
            if changed == 0:
                return u * fscale, it, True
        if not ok: break
        mu_target *= 1e-2
    return x * fscale, it, False


if __name__ == "__main__":
    from proto_ipm import exact_active_set
    rng = np.random.default_rng(5)
    errs = []; fails = 0
    ntest = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    if len(sys.argv) > 2:
        REFINE = int(sys.argv[2])
    for t in range(ntest):
        st = gen_state(rng, wide=(t % 2 == 1))
        st['contacts'] = np.array([1, 1, 1, 1]) if t % 3 else np.array([1, 1, 1, 0])[rng.permutation(4)]
        par = GAZEBO if t % 4 < 2 else HARDWARE
        u, it, ok = solve_wrench(st, par)
        H, g = build_literal(st, par); Hr, gr, idx = reduce_qp(H, g, st['contacts']); C, d = cons(len(gr))
        # reference: dense fp64 IPM + finisher from proto_pdas (validated against the long-double oracle on the GPU path)
        sys.path.insert(0, '/tmp')
        from proto_pdas import solve as solve_dense
        ud, itd, rd_, okd = solve_dense(Hr, gr, mu_switch=1e-9)
        e = np.abs(u - ud).max()
        errs.append(e); fails += (not ok)
        print(t, "S", int(st['contacts'].sum()), "its", it, "ok", ok, okd, "diff vs dense path %.2e" % e)
    print("max diff %.2e  fails %d/%d" % (max(errs), fails, ntest))

"""Design prototype (numpy) for the in-warp IPM.  NOT product code, NOT the oracle.

Used once, here on the CPU box, to pick the IPM variant / scaling / stopping rule
before the CUDA kernel was written.  Problem construction follows
ConvexMpc.cpp:110-245 and A1RobotControl.cpp:452-488 of the reference (restated
independently of oracle/ so that the two can be compared).
"""
import numpy as np

MU = 0.3
FZMAX = 180.0


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def rot_zyx(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


GAZEBO = dict(mass=12.0, inertia=np.diag([0.0158533, 0.0377999, 0.0456542]),
              q=np.array([20, 10, 1, 0, 0, 420, .05, .05, .05, 30, 30, 10, 0.]),
              r=np.full(12, 1e-7))
HARDWARE = dict(mass=13.5, inertia=np.diag([0.0178533, 0.0377999, 0.0456542]),
                q=np.array([150, 150, 50, 0, 0, 80, .2, .2, .2, .3, .3, .3, 0.]),
                r=np.array([1e-2, 1e-2, 1e-3] * 4))


def gen_state(rng, wide=False):
    yaw = rng.uniform(-np.pi, np.pi)
    roll, pitch = rng.normal(0, 0.02, 2)
    R = rot_zyx(roll, pitch, yaw)
    pos = np.array([rng.normal(), rng.normal(), rng.uniform(0.25, 0.32) if wide else rng.uniform(0.296, 0.304)])
    omega = rng.normal(0, 0.1, 3)
    vcmd = np.array([rng.uniform(-.6, .6), rng.uniform(-.3, .3), 0.0])
    yawrate_d = rng.uniform(-.8, .8)
    v = R @ vcmd + rng.normal(0, 0.03, 3)
    default = np.array([[.17, .17, -.17, -.17], [.15, -.15, .15, -.15], [-.35] * 4])
    foot = default.copy()
    foot[0] += rng.uniform(-.1, .1, 4)
    foot[1] += rng.uniform(-.1, .1, 4)
    foot[2] += rng.normal(0, .01, 4)
    foot = R @ foot
    u = rng.uniform()
    contacts = [1, 0, 0, 1] if u < .45 else ([0, 1, 1, 0] if u < .9 else [1, 1, 1, 1])
    return dict(euler=np.array([roll, pitch, yaw]), pos=pos, omega=omega, vel=v, R=R, foot=foot,
                euler_d=np.zeros(3), pos_d=np.array([0, 0, .3]), vel_d=vcmd,
                omega_d=np.array([0, 0, yawrate_d]), contacts=np.array(contacts))


def build_literal(st, par, N=10, dt=0.0025):
    """dense rollout exactly as the reference formulates it"""
    yaw = st['euler'][2]
    c, s = np.cos(yaw), np.sin(yaw)
    Ac = np.zeros((13, 13))
    Ac[0:3, 6:9] = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
    Ac[3:6, 9:12] = np.eye(3)
    Ac[11, 12] = 1
    Iw = st['R'] @ par['inertia'] @ st['R'].T
    Bc = np.zeros((13, 12))
    for i in range(4):
        Bc[6:9, 3 * i:3 * i + 3] = np.linalg.inv(Iw) @ skew(st['foot'][:, i])
        Bc[9:12, 3 * i:3 * i + 3] = np.eye(3) / par['mass']
    Ad = np.eye(13) + Ac * dt
    Bd = Bc * dt
    Aqp = np.zeros((13 * N, 13))
    Bqp = np.zeros((13 * N, 12 * N))
    for i in range(N):
        Aqp[13 * i:13 * i + 13] = Ad if i == 0 else Aqp[13 * (i - 1):13 * i] @ Ad
        for j in range(i + 1):
            Bqp[13 * i:13 * i + 13, 12 * j:12 * j + 12] = Bd if i == j else Aqp[13 * (i - j - 1):13 * (i - j)] @ Bd
    Q = 2 * np.tile(par['q'], N)
    Rw = 2 * np.tile(par['r'], N)
    H = Bqp.T @ (Q[:, None] * Bqp) + np.diag(Rw)
    x0 = np.concatenate([st['euler'], st['pos'], st['omega'], st['vel'], [-9.8]])
    vdw = st['R'] @ st['vel_d']
    xd = np.zeros(13 * N)
    for i in range(N):
        xd[13 * i:13 * i + 13] = [st['euler_d'][0], st['euler_d'][1], st['euler'][2] + st['omega_d'][2] * dt * (i + 1),
                                  st['pos'][0] + vdw[0] * dt * (i + 1), st['pos'][1] + vdw[1] * dt * (i + 1),
                                  st['pos_d'][2], *st['omega_d'], vdw[0], vdw[1], 0, -9.8]
    g = Bqp.T @ (Q * (Aqp @ x0 - xd))
    return H, g


def reduce_qp(H, g, contacts, N=10):
    idx = [12 * k + 3 * i + a for k in range(N) for i in range(4) if contacts[i] for a in range(3)]
    idx = np.array(idx)
    return H[np.ix_(idx, idx)], g[idx], idx


def cons(n):
    """C u <= d, 5 rows per foot-step"""
    K = n // 3
    C = np.zeros((5 * K, n))
    d = np.zeros(5 * K)
    for k in range(K):
        C[5 * k + 0, 3 * k:3 * k + 3] = [-1, 0, -MU]
        C[5 * k + 1, 3 * k:3 * k + 3] = [1, 0, -MU]
        C[5 * k + 2, 3 * k:3 * k + 3] = [0, -1, -MU]
        C[5 * k + 3, 3 * k:3 * k + 3] = [0, 1, -MU]
        C[5 * k + 4, 3 * k:3 * k + 3] = [0, 0, 1]
        d[5 * k + 4] = FZMAX
    return C, d


def ipm(H, g, C, d, tol=1e-13, maxit=40, fscale=100.0, verbose=False):
    """Mehrotra predictor-corrector on min 1/2 u'Hu+g'u, Cu<=d, scaled."""
    n = len(g)
    m = len(d)
    # scaling: u = fscale * x ; cost / cs
    Hs = H * fscale * fscale
    gs = g * fscale
    cs = np.abs(Hs).max()
    Hs = Hs / cs
    gs = gs / cs
    ds = d / fscale
    x = np.zeros(n)
    for k in range(n // 3):
        x[3 * k + 2] = 0.5 * FZMAX / fscale * 0.5
    s = ds - C @ x
    s = np.maximum(s, 1e-2)
    lam = np.ones(m) * 1e-2 / s * 1.0
    lam = np.full(m, 1.0) * (np.abs(gs).max() + 1e-3)
    hist = []
    for it in range(maxit):
        rd = Hs @ x + gs + C.T @ lam
        rp = C @ x + s - ds
        mu = s @ lam / m
        hist.append((np.abs(rd).max(), np.abs(rp).max(), mu))
        if verbose:
            print(it, hist[-1])
        if mu < tol and np.abs(rd).max() < tol * 10 and np.abs(rp).max() < 1e-12:
            break
        w = lam / s
        Kmat = Hs + C.T @ (w[:, None] * C)
        try:
            L = np.linalg.cholesky(Kmat)
        except np.linalg.LinAlgError:
            break

        def solve(rc):
            rhs = -rd + C.T @ (rc / s - w * rp)
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds_ = -rp - C @ dx
            dl = -(rc + lam * ds_) / s
            return dx, ds_, dl

        dxa, dsa, dla = solve(s * lam)
        aa_p = min(1.0, (-s[dsa < 0] / dsa[dsa < 0]).min() if (dsa < 0).any() else 1.0)
        aa_d = min(1.0, (-lam[dla < 0] / dla[dla < 0]).min() if (dla < 0).any() else 1.0)
        aa = min(aa_p, aa_d)
        mu_aff = (s + aa * dsa) @ (lam + aa * dla) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds_, dl = solve(s * lam + dsa * dla - sigma * mu)
        ap = min(1.0, 0.995 * (-s[ds_ < 0] / ds_[ds_ < 0]).min() if (ds_ < 0).any() else 1.0)
        ad = min(1.0, 0.995 * (-lam[dl < 0] / dl[dl < 0]).min() if (dl < 0).any() else 1.0)
        a = min(ap, ad)
        x = x + a * dx
        s = s + a * ds_
        lam = lam + a * dl
    return x * fscale, lam * cs / fscale, it + 1, hist


def exact_active_set(H, g, C, d, u0, lam0, maxround=50):
    """primal-dual active-set polish with KKT certificate (fp64 here; long double in the oracle)."""
    n = len(g)
    act = lam0 > (d - C @ u0)
    for rnd in range(maxround):
        A = C[act]
        b = d[act]
        na = A.shape[0]
        KKT = np.block([[H, A.T], [A, np.zeros((na, na))]])
        rhs = np.concatenate([-g, b])
        try:
            sol = np.linalg.solve(KKT, rhs)
        except np.linalg.LinAlgError:
            sol = np.linalg.lstsq(KKT, rhs, rcond=None)[0]
        u = sol[:n]
        la = sol[n:]
        viol = C @ u - d
        viol[act] = -1
        bad_p = viol > 1e-9
        idx_act = np.where(act)[0]
        bad_d = la < -1e-12
        if not bad_p.any() and not bad_d.any():
            lam = np.zeros(len(d))
            lam[idx_act] = la
            return u, lam, rnd
        if bad_d.any():
            act[idx_act[np.argmin(la)]] = False
        elif bad_p.any():
            act[np.argmax(viol)] = True
    raise RuntimeError("active set did not converge")


if __name__ == "__main__":
    import sys
    rng = np.random.default_rng(1)
    par = GAZEBO
    errs = []
    its = []
    for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 50):
        st = gen_state(rng)
        H, g = build_literal(st, par)
        Hr, gr, idx = reduce_qp(H, g, st['contacts'])
        C, d = cons(len(gr))
        u, lam, nit, hist = ipm(Hr, gr, C, d)
        ue, le, rnd = exact_active_set(Hr, gr, C, d, u, lam)
        errs.append(np.abs(u - ue)[:len(u) // 10].max())
        its.append(nit)
        print(t, st['contacts'], nit, rnd, "err_f0 %.2e" % errs[-1], "err_all %.2e" % np.abs(u - ue).max(), "cond %.1e" % np.linalg.cond(Hr))
    print("max err", max(errs), "iters p50", np.median(its), "max", max(its))

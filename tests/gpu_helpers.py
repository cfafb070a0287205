import numpy as np


def discrete_model(cfg, st, b):
    """A_d, B_d (13x13, 13x12), x0 (13), x_d (13N) exactly as compute_grf assembles them
    (A1RobotControl.cpp:452-514, ConvexMpc.cpp:110-156) -- numpy, used to feed the general qp_mats entry points"""
    N, dt = cfg.horizon, cfg.dt
    e = st["x0"][0:3, b]; p = st["x0"][3:6, b]; w = st["x0"][6:9, b]; v = st["x0"][9:12, b]
    R = st["rot"][:, b].reshape(3, 3)
    foot = st["foot"][:, b].reshape(4, 3)
    ref = st["ref"][:, b]
    c, s = np.cos(e[2]), np.sin(e[2])
    Ac = np.zeros((13, 13))
    Ac[0:3, 6:9] = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
    Ac[3:6, 9:12] = np.eye(3)
    Ac[11, 12] = 1
    I = np.array(list(cfg.inertia)).reshape(3, 3)
    Iw = R @ I @ R.T
    Bc = np.zeros((13, 12))
    for i in range(4):
        r = foot[i]
        S = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        Bc[6:9, 3 * i:3 * i + 3] = np.linalg.inv(Iw) @ S
        Bc[9:12, 3 * i:3 * i + 3] = np.eye(3) / cfg.mass
    Ad = np.eye(13) + Ac * dt
    Bd = Bc * dt
    x0 = np.concatenate([e, p, w, v, [-9.8]])
    vdw = R @ ref[5:8]
    xd = np.zeros(13 * N)
    for i in range(N):
        xd[13 * i:13 * i + 13] = [ref[0], ref[1], e[2] + ref[4] * dt * (i + 1), p[0] + vdw[0] * dt * (i + 1), p[1] + vdw[1] * dt * (i + 1),
                                  ref[8], ref[2], ref[3], ref[4], vdw[0], vdw[1], 0, -9.8]
    return Ad, Bd, x0, xd

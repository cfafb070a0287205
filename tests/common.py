import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")) as fh:
        return json.load(fh)


def golden_groups():
    """golden cases grouped by (horizon, weights) -> (cfg kwargs, state dict, f_body [12,B])"""
    g = load_golden()
    groups = {}
    for c in g["cases"]:
        groups.setdefault((c["horizon"], c["weights"]), []).append(c)
    out = []
    for (hz, w), cs in sorted(groups.items()):
        st = dict(x0=np.array([c["x0"] for c in cs]).T.copy(), rot=np.array([c["rot"] for c in cs]).T.copy(),
                  foot=np.array([c["foot"] for c in cs]).T.copy(), ref=np.array([c["ref"] for c in cs]).T.copy(),
                  contact=np.array([c["contact"] for c in cs], dtype=np.uint32))
        f = np.array([c["f_body"] for c in cs]).T.copy()
        out.append((hz, w, g["weights"][w], st, f))
    return out


def obatch(O, st, sl=None):
    if sl is None:
        return O.Batch(st["x0"], st["rot"], st["foot"], st["ref"], st["contact"])
    return O.Batch(st["x0"][:, sl], st["rot"][:, sl], st["foot"][:, sl], st["ref"][:, sl], st["contact"][sl])


def check_feasible(f_world_u, mu, fzmax, contact, tol=1e-7):
    """u_full [12N,B] world-frame: friction pyramid and fz bounds (ConvexMpc.cpp:46-58, 223-245)"""
    n12, B = f_world_u.shape
    u = f_world_u.reshape(n12 // 12, 4, 3, B)
    fx, fy, fz = u[:, :, 0], u[:, :, 1], u[:, :, 2]
    c = np.array([[(int(m) >> leg) & 1 for m in contact] for leg in range(4)])[None]  # [1,4,B]
    assert (np.abs(fx) <= mu * fz + tol).all() and (np.abs(fy) <= mu * fz + tol).all()
    assert (fz >= -tol).all() and (fz <= fzmax * c + tol).all()

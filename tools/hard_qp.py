"""dev tool: the one QP of the 1.44 M robustness sweep that the DMMA kernels left at IPM_ONLY (err 7e-5 N), alone and in company"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
from oracle import oracle_py as O
d = dict(np.load(os.path.join(ROOT, "tools", "data", "hard_qp_63168.npz")))
fo, info = O.compute_grf_batch(O.make_config(horizon=10), O.Batch(d["x0"], d["rot"], d["foot"], d["ref"], d["contact"]), O.MODE_EXACT, nthreads=4)
for tol in (0.0, 1e-10, 1e-8):
    eng = a1mpc.Engine(a1mpc.default_config(horizon=10, tol=tol))
    for name, sl in (("alone", slice(0, 1)), ("with 16 neighbours", slice(0, 17))):
        st = {k: (v[sl].copy() if k == "contact" else v[:, sl].copy()) for k, v in d.items()}
        f, status, iters = eng.solve(st)
        print("lib %s tol %.0e %-18s status %s iters %s err %.2e" % (os.environ.get("A1MPC_LIB", "default")[-12:], tol, name, status[:3], iters[:3], np.abs(f - fo[:, sl]).max(axis=0)[0]), flush=True)
    # 4096 copies with tiny perturbations of the state: how wide is the hard region?
    rng = np.random.default_rng(3)
    st = {k: (np.repeat(v[:1], 4096) if k == "contact" else np.repeat(v[:, :1], 4096, axis=1).copy()) for k, v in d.items()}
    st["x0"] += 1e-9 * rng.standard_normal(st["x0"].shape)
    f, status, iters = eng.solve(st)
    print("   4096 perturbed (1e-9) copies: status hist", np.bincount(status, minlength=5), "rounds max", (iters // 100).max(), "ipm max", (iters % 100).max(), flush=True)
    eng.close()

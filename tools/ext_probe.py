"""dev tool: config-4 path -- status histogram, timing, oracle spot check"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
from oracle import oracle_py as O
for N, B in ((10, 16384), (10, 131072), (20, 16384)):
    eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
    st = a1mpc.gen_states(B, 4, 77); sched, normals = a1mpc.gen_schedule(B, N, 4, 77)
    f, status, iters = eng.solve_ext(st, sched, normals)
    t = time.time(); f, status, iters = eng.solve_ext(st, sched, normals); dt = time.time() - t
    print("N=%d B=%d ext: wall %.3fs (%.0f QPs/s incl H2D/D2H) status %s ipm max %d rounds max %d" % (N, B, dt, B / dt, np.bincount(status, minlength=5), (iters % 100).max(), (iters // 100).max()), flush=True)
    rng = np.random.default_rng(0); chk = np.concatenate([np.nonzero(status != 0)[0][:10], rng.choice(B, 200, replace=False)])
    sub = {k: (v[chk].copy() if k == "contact" else v[:, chk].copy()) for k, v in st.items()}
    fo, info = O.compute_grf_batch_ext(O.make_config(horizon=N), O.Batch(sub["x0"], sub["rot"], sub["foot"], sub["ref"], sub["contact"]), sched[:, chk].copy(), normals[:, chk].copy(), O.MODE_EXACT, nthreads=32)
    print("   spot check: max err %.2e, oracle verified %d/%d" % (np.abs(f[:, chk] - fo).max(), int(info[:, 1].sum()), len(chk)), flush=True)
    eng.close()

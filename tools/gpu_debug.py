"""first-contact GPU check: build parity, solve parity vs oracle on small batches (dev tool)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
from oracle import oracle_py as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfgid = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = a1mpc.default_config(horizon=N)
eng = a1mpc.Engine(cfg)
ocfg = O.make_config(horizon=N)
st = a1mpc.gen_states(B, cfgid, 0)
ob = O.Batch(st["x0"], st["rot"], st["foot"], st["ref"], st["contact"])
print("contact classes", np.bincount([bin(c).count("1") for c in st["contact"]], minlength=5))
# ---- P0 build parity
nb = min(B, 16)
sub = {k: (v[:nb] if k == "contact" else v[:, :nb]) for k, v in st.items()}
H, g, lb, ub = eng.build_qp(sub)
for b in range(min(nb, 4)):
    Ho, go, Ao, lbo, ubo = O.build_qp(ocfg, ob, b)
    print("build b=%d  relH %.2e relg %.2e  lb %s ub %s" % (b, abs(H[b] - Ho).max() / abs(Ho).max(), abs(g[b] - go).max() / abs(go).max(),
          np.array_equal(lb[b], lbo), np.array_equal(ub[b], ubo)))
# ---- solve parity
t = time.time(); f, status, iters, u = eng.solve(st, want_u=True); t1 = time.time() - t
print("gpu solve wall %.3fs  status hist %s" % (t1, np.bincount(status, minlength=5)))
print("iters ipm p50 %d max %d ; rounds p50 %d max %d" % (np.median(iters % 100), (iters % 100).max(), np.median(iters // 100), (iters // 100).max()))
t = time.time(); fo, info, uo = O.compute_grf_batch(ocfg, ob, O.MODE_EXACT, nthreads=8, want_u=True); t2 = time.time() - t
print("oracle exact %.2fs verified %d/%d  kkt_stat max %.1e prim %.1e dual %.1e" % (t2, int(info[:, 1].sum()), B, info[:, 2].max(), info[:, 3].max(), info[:, 4].max()))
err = abs(f - fo).max(axis=0)
erru = abs(u.T - uo).max(axis=1)
print("f_body err: p50 %.2e p99 %.2e max %.2e   u_full err max %.2e" % (np.median(err), np.percentile(err, 99), err.max(), erru.max()))
bad = np.argsort(-err)[:5]
for b in bad:
    print("  b=%d contact=%s status=%d iters=%d err=%.3e" % (b, bin(st["contact"][b]), status[b], iters[b], err[b]))
    if err[b] > 1e-3:
        print("   gpu", f[:, b]); print("   ora", fo[:, b])

import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
from oracle import oracle_py as O
B = int(sys.argv[1]); stream = int(sys.argv[2]); cid = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N = int(sys.argv[4]) if len(sys.argv) > 4 else 10
eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
st = a1mpc.gen_states(B, cid, stream)
f, status, iters = eng.solve(st)
bad = np.nonzero(status != 0)[0]
print("bad", len(bad), "of", B, "status hist", np.bincount(status, minlength=5))
print("iters hist (ipm)", np.bincount(iters % 100)[:45])
print("rounds hist", np.bincount(iters // 100)[:20])
for b in bad[:10]:
    sub = {k: (v[b:b + 1].copy() if k == "contact" else v[:, b:b + 1].copy()) for k, v in st.items()}
    fo, info = O.compute_grf_batch(O.make_config(horizon=N), O.Batch(sub["x0"], sub["rot"], sub["foot"], sub["ref"], sub["contact"]), O.MODE_EXACT)
    print("b=%d contact=%s status=%d iters=%d  err vs oracle %.3e  oracle verified %d ipm %d rounds %d" % (b, bin(st["contact"][b]), status[b], iters[b], np.abs(f[:, b] - fo[:, 0]).max(), info[0, 1], info[0, 0], info[0, 5]))
    np.save("gpurun_out/bad%d_%d.npy" % (N, b), np.concatenate([sub["x0"][:, 0], sub["rot"][:, 0], sub["foot"][:, 0], sub["ref"][:, 0], [float(sub["contact"][0])]]))

"""dev tool for ncu: a few solves of one forced stance pattern"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
B = int(sys.argv[1]); pat = int(sys.argv[2], 0); N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
st = a1mpc.gen_states(B, 2, 3)
if pat:
    st["contact"][:] = pat
d = a1mpc.DeviceBatch(eng, B); d.upload(st)
for _ in range(4):
    eng.solve_ptrs(B, d.inp, d.out)
eng.sync()

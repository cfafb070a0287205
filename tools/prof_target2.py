"""dev tool for ncu: the benchmark mix (45/45/10 % stance patterns) at the two batch sizes the bench line and config 5 use,
so that one capture holds every solve kernel of the default path (`ncu -k regex:solve_kernel`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 32768]
eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
for B in sizes:
    st = a1mpc.gen_states(B, 2, 3)
    d = a1mpc.DeviceBatch(eng, B); d.upload(st)
    for _ in range(2):
        eng.solve_ptrs(B, d.inp, d.out)
    eng.sync()

"""dev tool: quick timing of the main classes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_probe import run
import a1mpc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
for pattern, name in ((0b1001, "trot(2)"), (0b1111, "four(4)"), (None, "mix")):
    for B in (1, 1024, 16384):
        ms, fact, ok = run(eng, B, pattern, reps=5 if B > 4096 else 20, N=N)
        print("N=%d %-9s B=%6d  %9.3f ms/batch  %10.0f QPs/s  fact/QP %.2f  optimal %.3f" % (N, name, B, ms, B / ms * 1e3, fact, ok), flush=True)

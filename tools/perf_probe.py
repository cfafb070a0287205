"""dev tool: device time of one solve_batch for a forced stance pattern and batch size"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc

def run(eng, B, pattern, reps=20, N=10):
    st = a1mpc.gen_states(B, 2, 3)
    if pattern is not None:
        st["contact"][:] = pattern
    d = a1mpc.DeviceBatch(eng, B)
    d.upload(st)
    for _ in range(3):
        eng.solve_ptrs(B, d.inp, d.out)
    eng.sync()
    e0, e1 = eng.event(), eng.event()
    eng.record(e0)
    for _ in range(reps):
        eng.solve_ptrs(B, d.inp, d.out)
    eng.record(e1)
    ms = eng.elapsed_ms(e0, e1) / reps
    f, status = d.download()
    it = np.zeros(B, dtype=np.int32)
    a1mpc._check(a1mpc.lib().a1mpc_memcpy_d2h(eng.h, it.ctypes.data, d.iters, it.nbytes)); eng.sync()
    d.free()
    return ms, (it % 100 + it // 100).mean(), (status == 0).mean()

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    eng = a1mpc.Engine(a1mpc.default_config(horizon=N))
    for pattern, name in ((0b1001, "trot(2)"), (0b1111, "four(4)"), (0b0001, "one(1)"), (0b0111, "three(3)"), (None, "mix")):
        if N == 20 and pattern == 0b1111:
            continue
        for B in (1, 148, 1024, 8192, 32768):
            if B > 8192 and pattern in (0b0001, 0b0111):
                continue
            ms, fact, ok = run(eng, B, pattern, reps=5 if B > 4096 else 20, N=N)
            print("N=%d %-9s B=%6d  %9.3f ms/batch  %10.0f QPs/s  %8.2f us/QP-slot  fact/QP %.2f  optimal %.3f" % (N, name, B, ms, B / ms * 1e3, ms * 1e3 / B, fact, ok), flush=True)

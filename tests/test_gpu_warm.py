"""SURVEY 8f.3 -- device-resident warm start across control ticks (a1mpc_solve_batch_warm), through the C ABI on the GPU.
The reference keeps one warm-started OsqpEigen::Solver alive (A1RobotControl.h:67, A1RobotControl.cpp:522-538); here the
state kept on the device is every robot's verified active face.  Same checks as tests/test_emu.py runs on the CPU emulator."""
import ctypes as C

import numpy as np
import pytest

from common import obatch

pytestmark = pytest.mark.gpu

TOL_F = 1e-4      # N, north_star


@pytest.fixture(scope="module")
def a1(built):
    import a1mpc
    return a1mpc


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


def _next_tick(st, rng, noise):
    st2 = {k: v.copy() for k, v in st.items()}
    st2["x0"][3:6] += 0.0025 * st["x0"][9:12]
    st2["x0"][0:3] += 0.0025 * st["x0"][6:9]
    st2["x0"] += noise * rng.standard_normal(st2["x0"].shape) * np.array([.02, .02, .02, .01, .01, .005, .1, .1, .1, .05, .05, .05])[:, None]
    return st2


def test_warm_start_across_ticks(a1, O):
    B = 2048
    eng = a1.Engine(a1.default_config(horizon=10))
    st = a1.gen_states(B, 2, 5)
    for i, p in enumerate([1, 2, 4, 8, 7, 11, 13, 14, 15, 0]):
        st["contact"][i] = p
    nz = st["contact"] != 0
    warm = eng.warm_alloc(B)
    f1, s1, it1 = eng.solve_warm(st, warm)
    fc, sc, itc = eng.solve(st)
    assert (s1[nz] == a1.STATUS_OPTIMAL).all() and ((it1 % 100)[nz] > 0).all()
    assert np.abs(f1 - fc).max() < 1e-8                         # no guess yet: the cold path (same algorithm, separately compiled kernel)
    rng = np.random.default_rng(0)
    st2 = _next_tick(st, rng, 0.03)
    st2["contact"][20:24] = [3, 5, 15, 6]
    changed = st2["contact"] != st["contact"]
    f2, s2, it2 = eng.solve_warm(st2, warm)
    ocfg = O.make_config(horizon=10)
    fo2, info = O.compute_grf_batch(ocfg, obatch(O, st2), O.MODE_EXACT, nthreads=O.hardware_threads())
    assert (s2[nz] == a1.STATUS_OPTIMAL).all() and np.abs(f2 - fo2).max() <= TOL_F
    hit = ((it2 % 100) == 0) & nz
    assert not hit[changed].any()
    assert hit[nz & ~changed].mean() > 0.8
    assert (it2 % 100 + it2 // 100)[nz & ~changed].mean() < 0.5 * (it1 % 100 + it1 // 100)[nz].mean()
    # a third tick keeps working from the faces stored by the second; a reset forgets them
    st3 = _next_tick(st2, rng, 0.03)
    f3, s3, it3 = eng.solve_warm(st3, warm)
    fo3, _ = O.compute_grf_batch(ocfg, obatch(O, st3), O.MODE_EXACT, nthreads=O.hardware_threads())
    assert (s3[nz] == a1.STATUS_OPTIMAL).all() and np.abs(f3 - fo3).max() <= TOL_F
    assert (((it3 % 100) == 0) & nz).mean() > 0.7
    assert a1.lib().a1mpc_warm_reset(eng.h, warm, B) == 0
    f4, s4, it4 = eng.solve_warm(st3, warm)
    assert ((it4 % 100)[nz] > 0).all() and np.abs(f4 - fo3).max() <= TOL_F
    eng.close()


def test_warm_start_argument_errors(a1):
    eng20 = a1.Engine(a1.default_config(horizon=20))
    st = a1.gen_states(8, 2, 1)
    w = eng20.dalloc(4096)
    with pytest.raises(a1.A1MpcError, match="horizon 10"):
        eng20.solve_warm(st, w)
    eng20.close()
    eng = a1.Engine(a1.default_config(horizon=10))
    host = np.zeros(8 * 44, dtype=np.uint32)
    with pytest.raises(a1.A1MpcError, match="device memory"):
        eng.solve_warm(st, host.ctypes.data_as(C.c_void_p))
    eng.close()

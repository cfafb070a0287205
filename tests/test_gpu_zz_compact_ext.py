"""BASELINE config 4 on the compacted class (a1mpc_sched.cuh): schedules with two stance feet in every horizon step solved as a
60-variable problem by the direct tensor-core kernel instead of the four-leg wrench-space form.  Opt-in this round
(A1MPC_EXT_COMPACT=1 when the handle is created); results must equal the oracle's and the general extended kernel's."""
import os

import numpy as np
import pytest

from common import obatch

pytestmark = pytest.mark.gpu

TOL_F = 1e-4


def test_compact_schedule_class_matches_oracle_and_general_kernel(built):
    import a1mpc as a1
    from oracle import oracle_py as O
    B = 1536
    st = a1.gen_states(B, 4, 77)
    sched, normals = a1.gen_schedule(B, 10, 4, 77)
    sched[:, 5] = 0b1111              # a robot with four feet down all the time: general kernel, inside the same call
    sched[3, 6] = 0b0111              # one step with three feet: general kernel
    gen = a1.Engine(a1.default_config(horizon=10))
    fg, sg, ig, ug = gen.solve_ext(st, sched, normals, want_u=True)
    gen.close()
    os.environ["A1MPC_EXT_COMPACT"] = "1"
    try:
        eng = a1.Engine(a1.default_config(horizon=10))
    finally:
        del os.environ["A1MPC_EXT_COMPACT"]
    f, status, iters, u = eng.solve_ext(st, sched, normals, want_u=True)
    eng.close()
    fo, info, uo = O.compute_grf_batch_ext(O.make_config(), obatch(O, st), sched, normals, O.MODE_EXACT, nthreads=O.hardware_threads(), want_u=True)
    assert (status == a1.STATUS_OPTIMAL).all(), np.bincount(status)
    assert np.abs(f - fo).max() <= TOL_F and np.abs(u.T - uo).max() <= TOL_F
    assert np.abs(f - fg).max() <= TOL_F
    # the compacted class needs fewer factorizations than the pinned four-leg form on the same problems
    two = np.array([all(bin(int(sched[s, b]) & 15).count("1") == 2 for s in range(10)) for b in range(B)])
    assert two.sum() > B // 2

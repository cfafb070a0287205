"""-m gpu: BASELINE configs 3 and 4 at their stated sizes, EVERY QP against the oracle (collected late: the largest batches).

config 3: trot, N = 20, batch 8192, precision 32 -- fp32 arrays at the boundary, fp64 arithmetic + in-kernel KKT certificate inside
          (include/a1mpc.h, "precision").  Contract: the returned forces are the optimum of the QP posed by the fp32-ROUNDED inputs,
          rounded to fp32:  |f - f*(rounded inputs)|_inf <= 1e-4 N + 1 fp32 ulp of 180 N (1.53e-5 N).
config 4: randomised contact schedules + terrain normals, batch 16384, fp64 (an extension beyond the reference; the oracle is the
          literal restatement generalised the same way), compacted two-feet class on by default:  <= 1e-4 N.
"""
import os

import numpy as np
import pytest

import a1mpc
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
NT = max(1, min(64, os.cpu_count() or 1))
ULP32_180 = float(np.spacing(np.float32(180.0)))


def _round32(st):
    return {k: (v if k == "contact" else v.astype(np.float32).astype(np.float64)) for k, v in st.items()}


def test_config3_n20_b8192_precision32_every_qp(built):
    B, N = 8192, 20
    eng = a1mpc.Engine(a1mpc.default_config(horizon=N, precision=32), device=0)
    st = a1mpc.gen_states(B, 2, 31)
    f, status, iters = eng.solve(st)
    eng.close()
    assert f.dtype == np.float32 and (status == 0).all(), np.bincount(status)
    st32 = _round32(st)
    fo, info = O.compute_grf_batch(O.make_config(horizon=N), O.Batch(st32["x0"], st32["rot"], st32["foot"], st32["ref"], st32["contact"]), O.MODE_EXACT, nthreads=NT)
    assert (info[:, 1] == 1).all()
    err = float(np.abs(f.astype(np.float64) - fo).max())
    assert err <= 1e-4 + ULP32_180, err


def test_precision32_n10_matches_fp64_engine_on_rounded_inputs_and_reports_input_sensitivity(built, gpu_engine):
    """same engine arithmetic either way: precision 32 on inputs x == precision 64 on float32(x), up to the output rounding;
    reported (not gated): how far the fp32 rounding of the INPUTS moves the optimum"""
    B = 2048
    st = a1mpc.gen_states(B, 2, 32)
    e32 = a1mpc.Engine(a1mpc.default_config(precision=32), device=0)
    f32, s32, _ = e32.solve(st)
    e32.close()
    f64r, s64r, _ = gpu_engine.solve(_round32(st))
    f64, s64, _ = gpu_engine.solve(st)
    assert (s32 == 0).all() and (s64r == 0).all() and (s64 == 0).all()
    assert np.abs(f32.astype(np.float64) - f64r).max() <= ULP32_180
    sens = float(np.abs(f64r - f64).max())
    print("input rounding to fp32 moves the optimal forces by up to %.3e N (N=10, narrow noise)" % sens)
    assert sens < 0.5       # sanity only: a well-posed QP


def test_config4_b16384_schedules_and_normals_every_qp(built):
    B, N = 16384, 10
    eng = a1mpc.Engine(a1mpc.default_config(horizon=N), device=0)
    st = a1mpc.gen_states(B, 4, 41)
    sched, normals = a1mpc.gen_schedule(B, N, 4, 41)
    f, status, iters = eng.solve_ext(st, sched, normals)
    eng.close()
    assert (status == 0).all(), np.bincount(status)
    fo, info = O.compute_grf_batch_ext(O.make_config(horizon=N), O.Batch(st["x0"], st["rot"], st["foot"], st["ref"], st["contact"]), sched, normals, O.MODE_EXACT, nthreads=NT)
    assert (info[:, 1] == 1).all()
    assert np.abs(f - fo).max() <= 1e-4

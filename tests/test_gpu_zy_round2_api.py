"""-m gpu: round-2 API additions through the C ABI -- the fused collect on a single rank, caller-owned memory that holds garbage,
invalid terrain normals, mixed host / device pointers, the rollout members."""
import ctypes as C

import numpy as np
import pytest

import a1mpc
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def test_fused_collect_single_rank_matches_f_body(built):
    """nranks = 1: the solve kernels' epilogue stores every QP's 12 forces as one QP-major record into the gathered buffer; after
    the wait it must equal f_body (batch-major) bit for bit, for fp64 and for precision 32"""
    for prec in (64, 32):
        eng = a1mpc.Engine(a1mpc.default_config(precision=prec), device=0)
        B = 2048
        st = a1mpc.gen_states(B, 2, 61)
        st["contact"][::97] = 0          # a few robots without any stance foot: the pack kernel writes their zeros
        hd = eng.peer_gather_create(1, 0, B)
        assert len(hd) == 64
        eng.peer_gather_connect([hd])
        d = a1mpc.DeviceBatch(eng, B)
        d.upload(st)
        for _ in range(2):
            eng.solve_ptrs(B, d.inp, d.out)
            eng.peer_gather_wait()
        eng.sync()
        assert eng.peer_gather_status() == 0
        f, status = d.download()
        g = np.zeros((B, 12), dtype=eng.ftype)
        a1mpc._check(a1mpc.lib().a1mpc_memcpy_d2h(eng.h, g.ctypes.data, eng.peer_gather_buffer(), g.nbytes))
        eng.sync()
        assert np.isin(status, (0, 4)).all() and (status == 4).sum() == len(st["contact"][::97])
        assert np.array_equal(g.T, f)
        # a different batch size falls back to the plain call (no peer stores, no signal)
        st2 = {k: (v[:100] if k == "contact" else v[:, :100]) for k, v in st.items()}
        f2, s2, _ = eng.solve(st2)
        assert np.isin(s2, (0, 4)).all()
        eng.peer_gather_destroy()
        d.free()
        eng.close()


def test_warm_slot_garbage_is_no_guess(built, gpu_engine):
    """a warm-start buffer that was never reset and happens to carry a valid header but impossible face codes (2-bit fields == 3)
    must be treated as "no guess": results are the cold path's"""
    B = 256
    st = a1mpc.gen_states(B, 2, 62)
    warm = gpu_engine.warm_alloc(B)
    words = 4 + 4 * 10
    junk = np.full((B, words), 0xFFFFFFFF, dtype=np.uint32)
    junk[:, 0] = 1; junk[:, 1] = st["contact"]; junk[:, 2] = 10; junk[:, 3] = 0
    a1mpc._check(a1mpc.lib().a1mpc_memcpy_h2d(gpu_engine.h, warm, junk.ctypes.data, junk.nbytes))
    gpu_engine.sync()
    f, status, iters = gpu_engine.solve_warm(st, warm, shift=0)
    fc, sc, ic = gpu_engine.solve(st)
    assert (status == 0).all() and np.abs(f - fc).max() <= 1e-9
    assert np.array_equal(iters, ic)          # the cold path, iteration for iteration


def test_terrain_normal_pointing_into_the_ground_is_flagged(built, gpu_engine):
    B, N = 64, 10
    st = a1mpc.gen_states(B, 4, 63)
    sched, normals = a1mpc.gen_schedule(B, N, 4, 63)
    normals[2, 5] = -abs(normals[2, 5])       # FL normal of robot 5 points down
    normals[5, 9] = 0.0; normals[3, 9] = 1.0; normals[4, 9] = 0.0   # FR normal of robot 9 horizontal (nz = 0)
    f, status, iters = gpu_engine.solve_ext(st, sched, normals)
    assert status[5] == 3 and status[9] == 3 and np.abs(f[:, [5, 9]]).max() == 0.0
    ok = np.ones(B, dtype=bool); ok[[5, 9]] = False
    assert (status[ok] == 0).all()


def test_mixed_host_and_device_pointers_are_rejected(built, gpu_engine):
    B = 32
    st = a1mpc.gen_states(B, 2, 64)
    d = a1mpc.DeviceBatch(gpu_engine, B)
    d.upload(st)
    f = np.zeros((12, B)); status = np.zeros(B, dtype=np.int32)
    # device inputs, host outputs
    out = a1mpc.Outputs(f.ctypes.data, status.ctypes.data, None, None, B)
    assert a1mpc.lib().a1mpc_solve_batch(gpu_engine.h, B, C.byref(d.inp), C.byref(out)) == -1
    # host x0, device rot (the case the round-1 advisor flagged: only some pointers were classified)
    inp = a1mpc.Inputs(st["x0"].ctypes.data, d.rot, st["foot"].ctypes.data, st["ref"].ctypes.data, st["contact"].ctypes.data, B)
    assert a1mpc.lib().a1mpc_solve_batch(gpu_engine.h, B, C.byref(inp), C.byref(out)) == -1
    assert b"all-host or all-device" in a1mpc.lib().a1mpc_last_error()
    # the handle is still healthy
    f2, s2, _ = gpu_engine.solve(st)
    assert (s2 == 0).all()
    d.free()


def test_rollout_members_match_the_oracle(built, gpu_engine):
    """a1mpc_qp_rollout_batch: the public ConvexMpc members A_qp, B_qp (ConvexMpc.h:77-78) next to H and g"""
    B = 6
    st = a1mpc.gen_states(B, 2, 65)
    ocfg = O.make_config()
    ob = O.Batch(st["x0"], st["rot"], st["foot"], st["ref"], st["contact"])
    ros = [O.rollout(ocfg, ob, b) for b in range(B)]
    Aq, Bq, H, g = gpu_engine.qp_rollout(np.stack([r["A_d"] for r in ros]), np.stack([r["B_d_list"] for r in ros]),
                                         np.stack([r["mpc_states"] for r in ros]), np.stack([r["mpc_states_d"] for r in ros]))
    for b in range(B):
        assert np.abs(Aq[b] - ros[b]["A_qp"]).max() <= 1e-14 and np.abs(Bq[b] - ros[b]["B_qp"]).max() <= 1e-16
        assert np.array_equal(Bq[b] == 0, ros[b]["B_qp"] == 0)
        Ho, go = O.qp_mats(ocfg, ros[b]["A_d"], ros[b]["B_d_list"], ros[b]["mpc_states"], ros[b]["mpc_states_d"])
        assert np.abs(H[b] - Ho).max() <= 1e-13 * np.abs(Ho).max() and np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()

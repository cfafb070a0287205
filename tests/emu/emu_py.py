"""ctypes binding of the CPU block emulator (tests/emu/liba1mpc_emu.so).  TEST INFRASTRUCTURE: the device code of
a1mpc_device.cuh compiled by g++ against a lane-accurate emulation of the warp primitives (cuda_emu.h)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200"))
import a1mpc  # noqa: E402  (struct definitions only; the emulator never touches liba1mpc.so)

_LIBS = {}


def lib(variant=""):
    """variant "" = the product's default switches; "nohyst" = -DA1MPC_FIN_HYST=0"""
    if variant not in _LIBS:
        name = "liba1mpc_emu%s.so" % ("_" + variant if variant else "")
        subprocess.check_call(["make", "-C", _HERE, "-s", name])
        _LIBS[variant] = C.CDLL(os.path.join(_HERE, name))
    return _LIBS[variant]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def solve(cfg, st, sched=None, normals=None, order=0, nthreads=8, want_u=False, variant="", warm=None, shift=1):
    """st: dict x0[12,B] rot[9,B] foot[12,B] ref[9,B] contact[B]; returns f_body[12,B], status[B], iters[B] (, u_full), stats"""
    B = st["contact"].shape[0]
    arrs = [np.ascontiguousarray(st[k], dtype=np.float64) for k in ("x0", "rot", "foot", "ref")]
    contact = np.ascontiguousarray(st["contact"], dtype=np.uint32)
    inp = a1mpc.Inputs(_p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(contact), B)
    f = np.zeros((12, B)); status = np.full(B, -7, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    u = np.zeros((12 * cfg.horizon, B)) if want_u else None
    out = a1mpc.Outputs(_p(f), _p(status), _p(iters), _p(u), B)
    sc = np.ascontiguousarray(sched, dtype=np.uint32) if sched is not None else None
    nm = np.ascontiguousarray(normals, dtype=np.float64) if normals is not None else None
    stats = (C.c_ulong * 2)()
    if warm is not None:
        assert warm.dtype == np.uint32 and warm.shape == (B, 4 + 4 * cfg.horizon) and warm.flags["C_CONTIGUOUS"]
    rc = lib(variant).emu_solve_batch(C.byref(cfg), B, C.byref(inp), _p(sc), _p(nm), C.byref(out), order, nthreads, stats, _p(warm), shift)
    assert rc == 0
    res = (f, status, iters) + ((u,) if want_u else ())
    return res + ({"collectives": int(stats[0]), "mma": int(stats[1])},)


def solve_dense(cfg, H, g, contact, order=0):
    """H [B,n,n], g [B,n], contact [B] -> u [B,n], status [B]  (a1mpc_solve_dense_batch on the emulator, N = 10)"""
    H = np.ascontiguousarray(H, dtype=np.float64); g = np.ascontiguousarray(g, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint32)
    B, n = g.shape
    u = np.zeros((B, n)); status = np.full(B, -7, dtype=np.int32)
    assert lib().emu_solve_dense(C.byref(cfg), B, _p(H), _p(g), _p(contact), _p(u), _p(status), order) == 0
    return u, status


def grf_qp(root_acc, rot_z, rot, foot, contact, order=0):
    """a1mpc_grf_qp_batch on the emulator: root_acc [B,6], rot_z/rot [B,9], foot [B,12], contact [B] -> f_body [B,12], status"""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (root_acc, rot_z, rot, foot)]
    contact = np.ascontiguousarray(contact, dtype=np.uint32)
    B = contact.shape[0]
    f = np.zeros((B, 12)); status = np.full(B, -7, dtype=np.int32)
    assert lib().emu_grf_qp(B, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(contact), _p(f), _p(status), order) == 0
    return f, status


def leg_kinematics(joint_pos, joint_vel, rot, rho_opt, rho_fix):
    """a1mpc_leg_kinematics_batch on the emulator: [12,B], [12,B], [9,B] -> foot_pos_rel, jac [36,B], foot_vel_rel, foot_pos_abs, foot_vel_abs"""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (joint_pos, joint_vel, rot, rho_opt, rho_fix)]
    B = a[0].shape[1]
    outs = [np.zeros((12, B)), np.zeros((36, B)), np.zeros((12, B)), np.zeros((12, B)), np.zeros((12, B))]
    assert lib().emu_leg_kinematics(B, *[_p(v) for v in a], *[_p(o) for o in outs]) == 0
    return outs


def ekf_init(foot_pos_rel, rot):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (foot_pos_rel, rot)]
    B = a[0].shape[1]
    state = np.zeros((B, 342))
    assert lib().emu_ekf_init(B, _p(state), _p(a[0]), _p(a[1])) == 0
    return state


def ekf_update(state, dt, assume_flat_ground, movement_mode, imu_acc, imu_ang_vel, rot, foot_pos_rel, foot_vel_rel, foot_force, order=0):
    """in place on state [B,342]; returns root_pos [3,B], root_lin_vel [3,B], estimated_contacts [B], status [B]"""
    B = state.shape[0]
    assert state.flags["C_CONTIGUOUS"] and state.dtype == np.float64
    mm = np.ascontiguousarray(movement_mode, dtype=np.uint32)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (imu_acc, imu_ang_vel, rot, foot_pos_rel, foot_vel_rel, foot_force)]
    pos = np.zeros((3, B)); vel = np.zeros((3, B)); ec = np.zeros(B, dtype=np.uint32); status = np.full(B, -7, dtype=np.int32)
    assert lib().emu_ekf_update(B, _p(state), C.c_double(dt), int(assume_flat_ground), _p(mm), *[_p(v) for v in a], _p(pos), _p(vel), _p(ec), _p(status), order) == 0
    return pos, vel, ec, status


def solve_sched2(cfg, st, sched, normals=None, order=0, nthreads=8, want_u=False):
    """the compacted two-stance-feet-per-step kernel (a1mpc_sched.cuh); same contract as solve(..., sched=...)"""
    B = st["contact"].shape[0]
    arrs = [np.ascontiguousarray(st[k], dtype=np.float64) for k in ("x0", "rot", "foot", "ref")]
    contact = np.ascontiguousarray(st["contact"], dtype=np.uint32)
    inp = a1mpc.Inputs(_p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(contact), B)
    f = np.zeros((12, B)); status = np.full(B, -7, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    u = np.zeros((12 * cfg.horizon, B)) if want_u else None
    out = a1mpc.Outputs(_p(f), _p(status), _p(iters), _p(u), B)
    sc = np.ascontiguousarray(sched, dtype=np.uint32)
    nm = np.ascontiguousarray(normals, dtype=np.float64) if normals is not None else None
    rc = lib().emu_solve_sched2(C.byref(cfg), B, C.byref(inp), _p(sc), _p(nm), C.byref(out), order, nthreads)
    assert rc == 0, rc
    return (f, status, iters, u) if want_u else (f, status, iters)

"""Generates tests/golden/convexmpc_v1.npz from the REFERENCE'S OWN code.

Source of every number in the file: oracle/_ref/libref_mpc.so = the reference's ConvexMpc.cpp, A1RobotControl.cpp, A1BasicEKF.cpp and
utils/Utils.cpp compiled UNMODIFIED from /root/reference against the header stand-ins in oracle/ref_shim/ (`make -C oracle ref`;
Eigen / OsqpEigen / ROS are not installable offline).  The file travels to the GPU box, which has no /root/reference.
The one thing the reference does not compute itself is the OSQP iteration (third party, absent): where a solution is stored
(`f_body`, `grf_f`) the QP that the reference handed to OsqpEigen::Solver was solved by the oracle's OSQP-algorithm restatement
run to eps 1e-11, and the file says so (`solver`); such values are only ever compared at >= 1e-5 N.

Contents (all float64, row-major):
  mpc_*   41 states: the test/test_mpc.cpp fixture, 24 narrow-noise + 8 wide-noise generator states (gazebo weights), 8 with the
          hardware weights.  Inputs in the a1mpc_inputs layout; from A1RobotControl::compute_grf (MPC branch): the QP exactly as handed
          to OsqpEigen -- g, lb, ub for all, the Hessian in full (upper triangle, packed) for the first 9 and as a sketch for
          all (diag(H) and H @ V for the fixed probe matrix `probe_V`, 8 columns) -- plus the state write-backs mpc_states,
          mpc_states_d, root_lin_vel_d_world and the returned body-frame forces.
  Ac      the constant pyramid matrix linear_constraints (200 x 120)
  roll_*  4 states through ConvexMpc member by member: A_qp (130 x 13), B_qp (130 x 120)
  step_*  2 per-step-B cases in the order of test/test_mpc.cpp:106-125 (feet move every step): A_d, B_d_list, x0, x_d -> H, g
  grf_*   8 states through compute_grf's single-step QP branch (BASELINE config 1): P (12 x 12), q, A (20 x 12), l, u, f_body
  tau_*   16 compute_joint_torques cases;  plan_* 16 update_plan ticks;  ekf_* one 30-tick A1BasicEKF run (x, P per tick)
Run:  python tests/golden/make_ref_golden.py        (CPU only, ~1 min; needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import a1mpc
from oracle import oracle_py as O
from oracle import ref_py as R
from common import estimation_scenario

WEIGHTS = {
    "gazebo": dict(mass=12.0, inertia=(0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                   q=(20, 10, 1, 0, 0, 420, .05, .05, .05, 30, 30, 10, 0), r=(1e-7,) * 12),          # config/gazebo_a1_mpc.yaml
    "hardware": dict(mass=13.5, inertia=(0.0178533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                     q=(150, 150, 50, 0, 0, 80, .2, .2, .2, .3, .3, .3, 0), r=(1e-2, 1e-2, 1e-3) * 4),  # config/hardware_a1_mpc.yaml
    "test_mpc": dict(mass=15.0, inertia=(0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                     q=(1, 1, 1, 0, 0, 50, 0, 0, 1, 1, 1, 1, 0), r=(1e-6,) * 12),                    # test/test_mpc.cpp:15-58
}
WNAMES = ["gazebo", "hardware", "test_mpc"]
NFULL = 9
N = 10
IU = np.triu_indices(12 * N)


def probe_matrix():
    return np.random.default_rng(20260923).standard_normal((12 * N, 8))


def mpc_cases():
    pats = [0b1001, 0b0110, 0b1111, 0b0001, 0b0111, 0b1110, 0b0011, 0b1000, 0b1011, 0b0101, 0b1101, 0b0010, 0b0100, 0b1100, 0b1010]
    cases = []
    _, fb = O.test_mpc_fixture()
    cases.append(("test_mpc", fb.x0[:, 0], fb.rot[:, 0], fb.foot[:, 0], fb.ref[:, 0], int(fb.contact[0])))
    for wname, cid, n, stream in (("gazebo", 2, 24, 1234), ("gazebo", 4, 8, 1235), ("hardware", 2, 8, 1236)):
        st = a1mpc.gen_states(n, cid, stream=stream)
        for b in range(n):
            c = int(st["contact"][b]) if b % 2 else pats[(b // 2) % len(pats)]   # generator patterns and every stance class
            cases.append((wname, st["x0"][:, b], st["rot"][:, b], st["foot"][:, b], st["ref"][:, b], c))
    return cases


def main():
    assert R.available(), "oracle/_ref/libref_mpc.so missing: run `make -C oracle ref` where /root/reference is mounted"
    out = {}
    V = probe_matrix()
    out["probe_V"] = V
    cases = mpc_cases()
    B = len(cases)
    keys = ["x0", "rot", "foot", "ref", "g", "lb", "ub", "Hdiag", "HV", "mpc_states", "mpc_states_d", "vd_world", "f_body"]
    acc = {k: [] for k in keys}
    Hfull, widx, contact = [], [], []
    for i, (wname, x0, rot, foot, ref, c) in enumerate(cases):
        cfg = O.make_config(**WEIGHTS[wname])
        r = R.compute_grf(cfg, x0, rot, foot, ref, c, control_type=1, solver="tight")
        P, q, A, l, u = r["qp"]
        assert P.shape == (120, 120) and np.array_equal(P, P.T)
        if i == 0:
            out["Ac"] = A.copy()
        else:
            assert np.array_equal(A, out["Ac"])
        if i < NFULL:
            Hfull.append(P[IU])
        for k, v in zip(keys, [x0, rot, foot, ref, q, l, u, np.diag(P).copy(), P @ V, r["mpc_states"], r["mpc_states_d"], r["root_lin_vel_d_world"], r["f_body"]]):
            acc[k].append(np.array(v, dtype=np.float64))
        widx.append(WNAMES.index(wname)); contact.append(c)
    for k in keys:
        out["mpc_" + k] = np.stack(acc[k])
    out["mpc_Hfull_triu"] = np.stack(Hfull)
    out["mpc_weights"] = np.array(widx, dtype=np.int32)
    out["mpc_contact"] = np.array(contact, dtype=np.uint32)
    for j, wname in enumerate(WNAMES):
        w = WEIGHTS[wname]
        out["w%d" % j] = np.array([w["mass"]] + list(w["inertia"]) + list(w["q"]) + list(w["r"]), dtype=np.float64)

    # --- rollout members ---
    roll = dict(idx=[], A_qp=[], B_qp=[], A_d=[], B_d_list=[])
    for i in (0, 1, 2, 35):
        wname, x0, rot, foot, ref, c = cases[i]
        w = WEIGHTS[wname]
        o = R.convexmpc(w["q"], w["r"], x0[0:3], w["mass"], w["inertia"], rot, foot, 0.0025, out["mpc_mpc_states"][i], out["mpc_mpc_states_d"][i], c)
        # the member-by-member drive reproduces what compute_grf handed to the solver, bit for bit
        assert np.array_equal(np.triu(o["H"])[IU], out["mpc_Hfull_triu"][i]) if i < NFULL else True
        assert np.array_equal(o["g"], out["mpc_g"][i]) and np.array_equal(o["lb"], out["mpc_lb"][i])
        roll["idx"].append(i)
        for k in ("A_qp", "B_qp", "A_d", "B_d_list"):
            roll[k].append(o[k])
    out["roll_idx"] = np.array(roll["idx"], dtype=np.int32)
    for k in ("A_qp", "B_qp", "A_d", "B_d_list"):
        out["roll_" + k] = np.stack(roll[k])

    # --- per-step B (test_mpc.cpp order) ---
    step = dict(A_d=[], B_d_list=[], x0=[], x_d=[], H_triu=[], g=[], weights=[])
    for i, shift in ((1, (0.4 * 0.0025, -0.2 * 0.0025, 0.0)), (3, (-0.5 * 0.0025, 0.3 * 0.0025, 0.05 * 0.0025))):
        wname, x0, rot, foot, ref, c = cases[i]
        w = WEIGHTS[wname]
        o = R.convexmpc(w["q"], w["r"], x0[0:3], w["mass"], w["inertia"], rot, foot, 0.0025, out["mpc_mpc_states"][i], out["mpc_mpc_states_d"][i], 0b1111, foot_shift=shift)
        for k, v in zip(("A_d", "B_d_list", "x0", "x_d", "H_triu", "g"), (o["A_d"], o["B_d_list"], out["mpc_mpc_states"][i], out["mpc_mpc_states_d"][i], o["H"][IU], o["g"])):
            step[k].append(v)
        step["weights"].append(WNAMES.index(wname))
    for k in step:
        out["step_" + k] = np.stack(step[k]) if k != "weights" else np.array(step[k], dtype=np.int32)

    # --- single-step GRF QP (config 1): standing default state first ---
    rng = np.random.default_rng(7)
    grf = {k: [] for k in ("x0", "rot", "rot_z", "foot", "ref12", "gains", "contact", "P", "q", "A", "l", "u", "f_body")}
    gains = np.array([1000.0, 1000, 1000, 200, 70, 120, 650, 35, 1, 4.5, 4.5, 30])     # A1CtrlStates.h:120-123
    cfg = O.make_config(**WEIGHTS["gazebo"])
    for i in range(8):
        if i == 0:
            x0 = np.zeros(12); x0[5] = 0.3
            rot = np.eye(3).reshape(-1); rot_z = rot.copy()
            foot = np.array([.17, .15, -.3, .17, -.15, -.3, -.17, .15, -.3, -.17, -.15, -.3])
            refv = np.zeros(9); refv[8] = 0.3
            c = 0b1111; pdxy = (0.0, 0.0); yaw_d = 0.0
        else:
            wname, x0, rot, foot, refv, c = cases[i]
            yaw = x0[2]
            rot_z = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]).reshape(-1)
            pdxy = tuple(x0[3:5] + rng.normal(0, 0.02, 2)); yaw_d = yaw + rng.normal(0, 0.05)
            if i == 7:
                c = 0b1111
        r = R.compute_grf(cfg, x0, rot, foot, refv, c, control_type=0, solver="tight", rot_z=rot_z, root_pos_d_xy=pdxy, yaw_d=yaw_d, gains=gains)
        P, q, A, l, u = r["qp"]
        ref12 = np.array([refv[0], refv[1], yaw_d, pdxy[0], pdxy[1], refv[8], refv[5], refv[6], refv[7], refv[2], refv[3], refv[4]])
        for k, v in zip(grf.keys(), (x0, rot, rot_z, foot, ref12, gains, c, P, q, A, l, u, r["f_body"])):
            grf[k].append(np.array(v))
    for k in grf:
        out["grf_" + k] = np.stack(grf[k]).astype(np.uint32 if k == "contact" else np.float64)

    # --- compute_joint_torques ---
    tau = {k: [] for k in ("f_grf", "f_kin", "jac", "contact", "km", "grav", "tau_prev", "tau")}
    grav = np.array([0.80, 0, 0, -0.80, 0, 0, 0.80, 0, 0, -0.80, 0, 0])
    for i in range(16):
        f_grf = rng.normal(0, 40, 12); f_kin = rng.normal(0, 20, 12); jac = rng.normal(0, 0.2, 36); c = int(rng.integers(0, 16))
        km = np.array([0.1, 0.1, 0.04]) if i % 2 else np.array([0.1, 0.1, 0.1])
        tp = rng.normal(0, 1, 12)
        if i == 15:
            jac[9:18] = 0.0     # singular swing-leg Jacobian -> NaN torques keep the previous values (A1RobotControl.cpp:313-317)
            c &= ~2
        t = R.joint_torques(f_grf, f_kin, jac, c, km, grav, tp)
        for k, v in zip(tau.keys(), (f_grf, f_kin, jac, c, km, grav, tp, t)):
            tau[k].append(np.array(v))
    for k in tau:
        out["tau_" + k] = np.stack(tau[k]).astype(np.uint32 if k == "contact" else np.float64)

    # --- update_plan ---
    plan = {k: [] for k in ("mode", "gc_in", "gcs", "lin_vel", "lin_vel_d", "rot_z", "rot", "root_pos", "gc_out", "plan", "trel", "tabs", "tworld")}
    dfp = np.array([0.17, 0.15, -0.35, 0.17, -0.15, -0.35, -0.17, 0.15, -0.35, -0.17, -0.15, -0.35])
    out["plan_params"] = np.array([240.0, 120.0, 0.0025])
    out["plan_default_foot_pos"] = dfp
    for i in range(16):
        wname, x0, rot, foot, refv, c = cases[1 + i]
        yaw = x0[2]
        rot_z = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]).reshape(-1)
        gc = rng.uniform(0, 240, 4); gcs = np.array([2.0, 2, 2, 2]) if i % 3 else rng.uniform(1, 3, 4)
        mode = 0 if i == 5 else 1
        lv = x0[9:12] * (3.0 if i % 4 == 0 else 1.0); lvd = refv[5:8]
        g2, pc, trel, tabs, tw = R.update_plan(240.0, 120.0, 0.0025, dfp, mode, gc, gcs, lv, lvd, rot_z, rot, x0[3:6])
        for k, v in zip(plan.keys(), (mode, gc, gcs, lv, lvd, rot_z, rot, x0[3:6], g2, pc, trel, tabs, tw)):
            plan[k].append(np.array(v))
    for k in plan:
        out["plan_" + k] = np.stack(plan[k]).astype(np.int32 if k in ("mode", "plan") else np.float64)

    # --- A1BasicEKF, 30 ticks, walking, one robot ---
    rngk, rho_opt, rho_fix, q, dq, rot = estimation_scenario(1, seed=11)
    T = 30
    ekf = R.Ekf(True)
    fpr0 = rngk.normal(0, 0.1, 12) + np.array([.18, .13, -.3, .18, -.13, -.3, -.18, .13, -.3, -.18, -.13, -.3])
    x, P = ekf.init(fpr0, rot[:, 0])
    e = dict(fpr0=fpr0, rot=rot[:, 0], x_init=x, P_init=P, acc=[], gyro=[], fpr=[], fvr=[], force=[], mode=[], x=[], P=[], pos=[], vel=[], ec=[])
    for t in range(T):
        acc = rngk.normal(0, 0.5, 3) + np.array([0, 0, 9.81]); gyro = rngk.normal(0, 0.2, 3)
        fpr = fpr0 + rngk.normal(0, 0.01, 12); fvr = rngk.normal(0, 0.3, 12); force = rngk.uniform(-20, 160, 4)
        mode = 0 if t < 3 else 1
        x, P, pos, vel, ec = ekf.update(0.0025, mode, acc, gyro, rot[:, 0], fpr, fvr, force)
        for k, v in zip(("acc", "gyro", "fpr", "fvr", "force", "mode", "x", "P", "pos", "vel", "ec"), (acc, gyro, fpr, fvr, force, mode, x, P, pos, vel, ec)):
            e[k].append(np.array(v))
    for k, v in e.items():
        out["ekf_" + k] = np.stack(v) if isinstance(v, list) else np.array(v)

    out["meta"] = np.array(["reference @ /root/reference (79c91302), compiled unmodified via oracle/ref_shim; solver for stored forces: "
                            "oracle OSQP-algorithm restatement, eps 1e-11 (OSQP absent offline); horizon %d" % N])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convexmpc_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.0f KB" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

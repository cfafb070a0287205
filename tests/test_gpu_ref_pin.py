"""-m gpu: the CUDA path, through the C ABI, against vectors produced by the REFERENCE'S OWN compiled code.

tests/golden/convexmpc_v1.npz comes from oracle/_ref/libref_mpc.so (the reference's ConvexMpc.cpp / A1RobotControl.cpp / A1BasicEKF.cpp / Utils.cpp
compiled unmodified against oracle/ref_shim; generator tests/golden/make_ref_golden.py).  Nothing here reads /root/reference or oracle/_ref.
Tolerances: P0 build 1e-13 relative; P1 forces 1e-4 N; the KKT certificate is evaluated on the REFERENCE-built H, g, Ac, lb, ub.
"""
import numpy as np
import pytest

import a1mpc
from common import load_ref_golden, ref_cfg_kwargs
from test_ref_pin import _root_acc

pytestmark = pytest.mark.gpu
N = 10
IU = np.triu_indices(12 * N)


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _states(G, sel):
    return dict(x0=G["mpc_x0"][sel].T.copy(), rot=G["mpc_rot"][sel].T.copy(), foot=G["mpc_foot"][sel].T.copy(), ref=G["mpc_ref"][sel].T.copy(),
                contact=G["mpc_contact"][sel].copy())


def _engines(G):
    for w in range(3):
        sel = np.nonzero(G["mpc_weights"] == w)[0]
        eng = a1mpc.Engine(a1mpc.default_config(**ref_cfg_kwargs(G, w)), device=0)
        yield w, sel, eng
        eng.close()


def test_gpu_build_matches_reference_built_qp(built):
    """a1mpc_build_qp_batch (ConvexMpc members hessian / gradient / lb / ub as compute_grf leaves them) vs the reference build"""
    G = load_ref_golden()
    V = G["probe_V"]
    worst = dict(H=0.0, HV=0.0, g=0.0)
    for w, sel, eng in _engines(G):
        H, g, lb, ub = eng.build_qp(_states(G, sel))
        for k, i in enumerate(sel):
            assert np.array_equal(lb[k], G["mpc_lb"][i]) and np.array_equal(ub[k], G["mpc_ub"][i])
            assert np.abs(H[k] - H[k].T).max() <= 1e-15 * np.abs(H[k]).max()   # symmetric to rounding (the two triangles are summed in different orders)
            if i < G["mpc_Hfull_triu"].shape[0]:
                worst["H"] = max(worst["H"], _relerr(H[k][IU], G["mpc_Hfull_triu"][i]))
            worst["HV"] = max(worst["HV"], _relerr(H[k] @ V, G["mpc_HV"][i]), _relerr(np.diag(H[k]), G["mpc_Hdiag"][i]))
            worst["g"] = max(worst["g"], _relerr(g[k], G["mpc_g"][i]))
    assert worst["H"] <= 1e-13 and worst["HV"] <= 1e-13 and worst["g"] <= 1e-12, worst


def test_gpu_general_rollout_matches_reference_built_qp(built):
    """a1mpc_qp_mats_batch (per-step B_d, ConvexMpc::calculate_qp_mats driven as test/test_mpc.cpp:106-125 drives it) vs the reference build"""
    G = load_ref_golden()
    for k in range(G["step_A_d"].shape[0]):
        eng = a1mpc.Engine(a1mpc.default_config(**ref_cfg_kwargs(G, int(G["step_weights"][k]))), device=0)
        H, g = eng.qp_mats(G["step_A_d"][k][None], G["step_B_d_list"][k][None], G["step_x0"][k][None], G["step_x_d"][k][None])
        eng.close()
        assert _relerr(H[0][IU], G["step_H_triu"][k]) <= 1e-13 and _relerr(g[0], G["step_g"][k]) <= 1e-12


def test_gpu_solutions_carry_a_kkt_certificate_on_the_reference_built_qp(built):
    """the whole-horizon GPU solution u must satisfy the KKT conditions of  min 1/2 u'Hu + g'u, lb <= Ac u <= ub  with H, g, Ac, lb, ub
    as the REFERENCE built them (9 states with the full Hessian on file): primal feasibility, and a sign-correct multiplier vector that
    closes stationarity -- found by non-negative least squares on the active rows.  Also forces vs the reference's compute_grf (<= 1e-4 N)."""
    from scipy.optimize import nnls
    G = load_ref_golden()
    nfull = G["mpc_Hfull_triu"].shape[0]
    Ac = G["Ac"]
    worst = dict(f=0.0, stat=0.0, prim=0.0)
    for w, sel, eng in _engines(G):
        f, status, iters, u = eng.solve(_states(G, sel), want_u=True)
        assert (status == 0).all(), status
        worst["f"] = max(worst["f"], float(np.abs(f.T - G["mpc_f_body"][sel]).max()))
        for k, i in enumerate(sel):
            if i >= nfull:
                continue
            H = np.zeros((12 * N, 12 * N)); H[IU] = G["mpc_Hfull_triu"][i]; H = H + H.T - np.diag(np.diag(H))
            g, lb, ub = G["mpc_g"][i], G["mpc_lb"][i], G["mpc_ub"][i]
            x = u[:, k]
            z = Ac @ x
            scale = max(1.0, np.abs(x).max())
            worst["prim"] = max(worst["prim"], float(np.maximum(lb - z, 0).max()), float(np.maximum(z - ub, 0).max()))
            # stationarity: H x + g + Ac' y = 0 with y_i >= 0 on rows at the upper bound, <= 0 at the lower bound, 0 elsewhere
            tol = 1e-7 * scale
            up = np.abs(z - ub) <= tol
            lo = np.abs(z - lb) <= tol
            cols = np.concatenate([Ac[up].T, -Ac[lo].T], axis=1)
            r = -(H @ x + g)
            sc = np.abs(r).max() + 1e-300
            y, res = nnls(cols / sc * 1.0, r / sc, maxiter=20000)
            worst["stat"] = max(worst["stat"], float(res))     # relative to |Hx+g|_inf
    assert worst["f"] <= 1e-4 and worst["prim"] <= 1e-7 and worst["stat"] <= 1e-6, worst


def test_gpu_grf_qp_matches_reference_compute_grf(built, gpu_engine):
    """BASELINE config 1 on the GPU (a1mpc_grf_qp_batch) vs the reference's compute_grf QP branch"""
    G = load_ref_golden()
    mass = float(G["w0"][0])
    acc = np.stack([_root_acc(G["grf_x0"][k].copy(), G["grf_rot"][k], G["grf_ref12"][k], G["grf_gains"][k], mass) for k in range(G["grf_P"].shape[0])])
    f, status = gpu_engine.grf_qp(acc, G["grf_rot_z"], G["grf_rot"], G["grf_foot"], G["grf_contact"])
    assert (status == 0).all() and np.abs(f - G["grf_f_body"]).max() <= 1e-4


def test_gpu_joint_torques_update_plan_match_reference(built, gpu_engine):
    """SURVEY 8f.1 / 8f.2 kernels vs A1RobotControl::compute_joint_torques / update_plan of the reference build"""
    G = load_ref_golden()
    for km in np.unique(G["tau_km"], axis=0):
        sel = np.nonzero((G["tau_km"] == km).all(axis=1))[0]
        tau = gpu_engine.joint_torques(G["tau_f_grf"][sel].T, G["tau_f_kin"][sel].T, G["tau_jac"][sel].T, G["tau_contact"][sel], km, G["tau_grav"][0], G["tau_tau_prev"][sel].T)
        assert np.abs(tau.T - G["tau_tau"][sel]).max() <= 1e-11 * max(1.0, np.abs(G["tau_tau"][sel]).max())
    gp = a1mpc.default_gait_params(horizon=N)
    gc, plan, sched, trel, tabs, tw = gpu_engine.update_plan(gp, G["plan_gc_in"].T, G["plan_gcs"].T, G["plan_mode"].astype(np.uint32), G["plan_lin_vel"].T, G["plan_lin_vel_d"].T,
                                                            G["plan_rot_z"].T, G["plan_rot"].T, G["plan_root_pos"].T)
    assert np.abs(gc.T - G["plan_gc_out"]).max() <= 1e-12 and np.array_equal(plan, G["plan_plan"].astype(np.uint32))
    assert np.abs(trel.T - G["plan_trel"]).max() <= 1e-14 and np.abs(tabs.T - G["plan_tabs"]).max() <= 1e-14 and np.abs(tw.T - G["plan_tworld"]).max() <= 1e-14


def test_gpu_ekf_matches_reference(built, gpu_engine):
    """the batched Kalman filter vs the reference's A1BasicEKF over 30 ticks (state device-resident on one side, inside the
    reference object on the other)"""
    G = load_ref_golden()
    ekf = gpu_engine.ekf_alloc(1)
    gpu_engine.ekf_init(ekf, G["ekf_fpr0"][:, None], G["ekf_rot"][:, None])
    x, P = gpu_engine.ekf_state(ekf, 1)
    assert np.abs(x[0] - G["ekf_x_init"]).max() <= 1e-15 and np.array_equal(P[0], G["ekf_P_init"])
    for t in range(G["ekf_x"].shape[0]):
        pos, vel, ec, status = gpu_engine.ekf_update(ekf, 0.0025, 1, np.array([G["ekf_mode"][t]], dtype=np.uint32), G["ekf_acc"][t][:, None], G["ekf_gyro"][t][:, None],
                                                     G["ekf_rot"][:, None], G["ekf_fpr"][t][:, None], G["ekf_fvr"][t][:, None], G["ekf_force"][t][:, None])
        assert status[0] == 0 and int(ec[0]) == int(G["ekf_ec"][t])
        assert np.abs(pos[:, 0] - G["ekf_pos"][t]).max() <= 1e-10 and np.abs(vel[:, 0] - G["ekf_vel"][t]).max() <= 1e-10, t
    x, P = gpu_engine.ekf_state(ekf, 1)
    assert np.abs(x[0] - G["ekf_x"][-1]).max() <= 1e-10 and np.abs(P[0] - G["ekf_P"][-1]).max() <= 1e-10

"""SURVEY 8f.4 -- upstream producers of the path's inputs, through the C ABI on the GPU: batched leg kinematics
(A1Kinematics::fk / jac, legKinematics/A1Kinematics.cpp) and the batched Kalman filter (A1BasicEKF.cpp).
The kinematics oracle is pinned to the reference's own source (tests/golden/kinematics_v1.json); the same comparisons run on
the CPU emulator in tests/test_emu.py."""
import numpy as np
import pytest

from common import check_kinematics, ekf_walk, estimation_scenario, load_kin_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def a1(built):
    import a1mpc
    return a1mpc


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="module")
def eng(a1):
    e = a1.Engine(a1.default_config(horizon=10))
    yield e
    e.close()


def test_leg_kinematics_golden_vectors_of_the_reference(a1, eng):
    g = load_kin_golden()["cases"]
    # rho_opt / rho_fix are batch-uniform parameters and differ from case to case: one call per golden case
    worst = 0.0
    for c in g:
        leg = c["leg"]
        q1 = np.zeros((12, 1)); q1[3 * leg:3 * leg + 3, 0] = c["q"]
        ro = np.zeros(12); ro[3 * leg:3 * leg + 3] = c["rho_opt"]
        rf = np.tile(np.array(c["rho_fix"]), 4)
        fpr, jac, fvr, fpa, fva = eng.leg_kinematics(q1, np.zeros((12, 1)), np.eye(3).reshape(9, 1), ro, rf)
        worst = max(worst, np.abs(fpr[3 * leg:3 * leg + 3, 0] - c["p"]).max(), np.abs(jac[9 * leg:9 * leg + 9, 0] - c["J"]).max())
    assert worst <= 1e-12, worst


def test_leg_kinematics_batch_chains_into_the_solver(a1, O, eng):
    B = 3000
    rng, rho_opt, rho_fix, q, dq, rot = estimation_scenario(B, 5)
    outs = eng.leg_kinematics(q, dq, rot, rho_opt.reshape(12), rho_fix.reshape(20))
    idx = rng.choice(B, 200, replace=False)
    check_kinematics(O, [o[:, idx] for o in outs], q[:, idx], dq[:, idx], rot[:, idx], rho_opt, rho_fix)
    # foot_pos_abs is the `foot` input of the hot path: solve with it
    st = a1.gen_states(B, 2, 9)
    st["rot"] = rot; st["foot"] = outs[3]
    f, status, iters = eng.solve(st)
    assert (status == a1.STATUS_OPTIMAL).all()


def test_ekf_batch_against_oracle(a1, O, eng):
    B = 300
    box = {}

    def init(fpr, rot_):
        box["ekf"] = eng.ekf_alloc(B)
        eng.ekf_init(box["ekf"], fpr, rot_)
        return lambda: eng.ekf_state(box["ekf"], B)

    def update(dt, flat, mode, acc, gyro, rot_, fpr, fvr, force, tick):
        return eng.ekf_update(box["ekf"], dt, flat, mode, acc, gyro, rot_, fpr, fvr, force)
    worst = ekf_walk(O, B, 25, eng.leg_kinematics, init, update, seed=11)
    assert worst < 1e-10, worst

// oracle/ref_mpc_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the REFERENCE'S OWN control sources, which
// `make -C oracle ref` compiles UNMODIFIED from where they lie under /root/reference/src/a1_cpp/src (ConvexMpc.cpp,
// A1RobotControl.cpp, A1BasicEKF.cpp, utils/Utils.cpp, legKinematics/A1Kinematics.cpp) against the header stand-ins in
// oracle/ref_shim/ (Eigen, OsqpEigen and ROS are not installed here) into oracle/_ref/libref_mpc.so.
//
// What this pins: everything the reference computes itself on the hot path -- x0 / x_d packing, A_c, B_c, the Euler
// discretisation, the A_qp / B_qp rollout, the dense Hessian, the gradient, the pyramid matrix, the bounds, the problem exactly
// as it is handed to OsqpEigen (upper triangle of H), the R^T u force extraction, the state write-backs -- plus
// compute_joint_torques, update_plan, the single-step GRF QP set-up and the Kalman filter.  What it cannot pin: OSQP's own
// iterations (third party, absent); the QP handed to OsqpEigen::Solver::solve() is forwarded to a solver installed with
// ref_set_qp_solver (tests install the oracle's OSQP-algorithm restatement run to eps 1e-11, or its exact solver).
//
// Nothing here is product code; nothing under a1-qp-mpc-controller_b200/ may link it.  tools/make_ref_golden.py turns its outputs
// into tests/golden/convexmpc_v1.json so that the GPU box (which has no /root/reference) can check against them.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include "OsqpEigen/OsqpEigen.h"
#include <ros/ros.h>
#include "utils/Utils.h"
#include "A1CtrlStates.h"
#include "ConvexMpc.h"
#include "A1RobotControl.h"
// the filter keeps its state (x, P) private; the tests compare it with the oracle's, so open the class up for this one header
#define private public
#include "A1BasicEKF.h"
#undef private

namespace {

struct LastQp {
  int n = 0, m = 0, calls = 0;
  std::vector<double> P, q, A, l, u;  // P, A column-major as handed over
};
LastQp g_last;
OsqpEigen::SolveHook g_user_solver = nullptr;

int capture_hook(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u, int warm,
                 double* x, double* y) {
  g_last.n = n; g_last.m = m; g_last.calls++;
  g_last.P.assign(P, P + (size_t)n * n);
  g_last.q.assign(q, q + n);
  g_last.A.assign(A, A + (size_t)m * n);
  g_last.l.assign(l, l + m);
  g_last.u.assign(u, u + m);
  if (!g_user_solver) return 1;
  return g_user_solver(n, m, P, q, A, l, u, warm, x, y);
}

struct CoutMute {  // the reference prints from constructors and from compute_grf
  std::streambuf* old;
  std::ostringstream sink;
  CoutMute() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutMute() { std::cout.rdbuf(old); }
};

Eigen::Matrix3d mat3_rowmajor(const double* a) {
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m(i, j) = a[3 * i + j];
  return m;
}
Eigen::Vector3d vec3(const double* a) { return Eigen::Vector3d(a[0], a[1], a[2]); }
Eigen::Matrix<double, 3, NUM_LEG> legs(const double* a) {  // leg-major [3*leg + axis] -> 3 x 4, column = leg
  Eigen::Matrix<double, 3, NUM_LEG> m;
  for (int l = 0; l < NUM_LEG; ++l)
    for (int k = 0; k < 3; ++k) m(k, l) = a[3 * l + k];
  return m;
}
template <class M> void out_rowmajor(const M& m, double* o) {
  if (!o) return;
  for (Eigen::Index i = 0; i < m.rows(); ++i)
    for (Eigen::Index j = 0; j < m.cols(); ++j) o[i * m.cols() + j] = m(i, j);
}
template <class M> void out_legs(const M& m, double* o) {  // 3 x 4 -> leg-major
  if (!o) return;
  for (int l = 0; l < NUM_LEG; ++l)
    for (int k = 0; k < 3; ++k) o[3 * l + k] = m(k, l);
}

}  // namespace

extern "C" {

int ref_plan_horizon(void) { return PLAN_HORIZON; }

// fn(n, m, P, q, A, l, u, warm, x, y) -> 0 on success; P (n x n) and A (m x n) are column-major
void ref_set_qp_solver(OsqpEigen::SolveHook fn) {
  g_user_solver = fn;
  OsqpEigen::solve_hook() = capture_hook;
}
int ref_last_qp_dims(int* n, int* m) { *n = g_last.n; *m = g_last.m; return g_last.calls; }
// the problem of the latest OsqpEigen::Solver::solve(), P and A ROW-major
int ref_last_qp(double* P, double* q, double* A, double* l, double* u) {
  const int n = g_last.n, m = g_last.m;
  if (!n) return 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) P[(size_t)i * n + j] = g_last.P[(size_t)j * n + i];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = g_last.A[(size_t)j * m + i];
  std::memcpy(q, g_last.q.data(), sizeof(double) * n);
  std::memcpy(l, g_last.l.data(), sizeof(double) * m);
  std::memcpy(u, g_last.u.data(), sizeof(double) * m);
  return 0;
}

// ConvexMpc driven member by member in the order A1RobotControl::compute_grf drives it (A1RobotControl.cpp:447, 492-517), or, with
// foot_shift != NULL, in the order of test/test_mpc.cpp:59-125 (the feet move by -foot_shift every step, so B_d differs per step).
// Outputs (any may be NULL), all ROW-major: A_qp [13N x 13], B_qp [13N x 12N], H [12N x 12N], g [12N], Ac [20N x 12N], lb, ub [20N],
// B_d_list [13N x 12], A_d [13 x 13].
int ref_convexmpc(const double* q13, const double* r12, const double* euler3, double mass, const double* inertia9,
                  const double* rot9, const double* foot12, const double* foot_shift3, double dt, const double* mpc_states13,
                  const double* mpc_states_d, uint32_t contacts, double* A_qp, double* B_qp, double* H, double* g, double* Ac,
                  double* lb, double* ub, double* B_d_list, double* A_d) {
  CoutMute mute;
  Eigen::VectorXd qw(13), rw(12);
  for (int i = 0; i < 13; ++i) qw(i) = q13[i];
  for (int i = 0; i < 12; ++i) rw(i) = r12[i];
  A1CtrlStates state;
  for (int i = 0; i < MPC_STATE_DIM; ++i) state.mpc_states(i) = mpc_states13[i];
  for (int i = 0; i < MPC_STATE_DIM * PLAN_HORIZON; ++i) state.mpc_states_d(i) = mpc_states_d[i];
  for (int i = 0; i < NUM_LEG; ++i) state.contacts[i] = (contacts >> i) & 1u;
  ConvexMpc mpc_solver = ConvexMpc(qw, rw);
  mpc_solver.reset();
  mpc_solver.calculate_A_mat_c(vec3(euler3));
  Eigen::Matrix<double, 3, NUM_LEG> foot = legs(foot12);
  const Eigen::Matrix3d I_b = mat3_rowmajor(inertia9), R = mat3_rowmajor(rot9);
  for (int i = 0; i < PLAN_HORIZON; i++) {
    mpc_solver.calculate_B_mat_c(mass, I_b, R, foot);
    if (foot_shift3)
      for (int l = 0; l < NUM_LEG; ++l) foot.block<3, 1>(0, l) = foot.block<3, 1>(0, l) - vec3(foot_shift3);
    mpc_solver.state_space_discretization(dt);
    mpc_solver.B_mat_d_list.block<13, 12>(i * 13, 0) = mpc_solver.B_mat_d;
  }
  mpc_solver.calculate_qp_mats(state);
  out_rowmajor(mpc_solver.A_qp, A_qp);
  out_rowmajor(mpc_solver.B_qp, B_qp);
  if (H) out_rowmajor(mpc_solver.hessian.toDense(), H);
  if (g) for (int i = 0; i < NUM_DOF * PLAN_HORIZON; ++i) g[i] = mpc_solver.gradient(i);
  if (Ac) out_rowmajor(mpc_solver.linear_constraints.toDense(), Ac);
  if (lb) for (int i = 0; i < MPC_CONSTRAINT_DIM * PLAN_HORIZON; ++i) lb[i] = mpc_solver.lb(i);
  if (ub) for (int i = 0; i < MPC_CONSTRAINT_DIM * PLAN_HORIZON; ++i) ub[i] = mpc_solver.ub(i);
  out_rowmajor(mpc_solver.B_mat_d_list, B_d_list);
  out_rowmajor(mpc_solver.A_mat_d, A_d);
  return 0;
}

// A1RobotControl::compute_grf itself (A1RobotControl.cpp:321-565), `ticks` calls on ONE controller object (so that the
// persistent, warm-started OsqpEigen::Solver of the MPC branch is exercised the way the control loop does).
//   control_type 0: single-step QP branch (:377-445), 1: MPC branch (:446-562)
//   state13x     root_euler(3) root_pos(3) root_ang_vel(3) root_lin_vel(3)            [the a1mpc_inputs x0 layout]
//   ref12        root_euler_d(3) root_pos_d(3) root_lin_vel_d(3, body) root_ang_vel_d(3)
//   gains12      kp_linear kd_linear kp_angular kd_angular (QP branch; NULL keeps the A1CtrlStates defaults)
// Outputs: f_body12 leg-major (the returned 3x4), mpc_states[13], mpc_states_d[13N], root_lin_vel_d_world[3], root_euler_d[3]
// as compute_grf leaves them in `state`.
int ref_compute_grf(int control_type, int use_terrain_adapt, double dt, int ticks, const double* q13, const double* r12, double mass,
                    const double* inertia9, const double* state12x, const double* rot9, const double* rot_z9, const double* foot_abs12,
                    uint32_t contacts, const double* ref12, const double* gains12, double* f_body12, double* mpc_states13,
                    double* mpc_states_d, double* root_lin_vel_d_world3, double* root_euler_d3) {
  CoutMute mute;
  A1RobotControl ctrl;
  A1CtrlStates state;
  state.stance_leg_control_type = control_type;
  state.use_terrain_adapt = use_terrain_adapt;
  state.robot_mass = mass;
  state.a1_trunk_inertia = mat3_rowmajor(inertia9);
  for (int i = 0; i < 13; ++i) state.q_weights(i) = q13[i];
  for (int i = 0; i < 12; ++i) state.r_weights(i) = r12[i];
  state.root_euler = vec3(state12x);
  state.root_pos = vec3(state12x + 3);
  state.root_ang_vel = vec3(state12x + 6);
  state.root_lin_vel = vec3(state12x + 9);
  state.root_rot_mat = mat3_rowmajor(rot9);
  state.root_rot_mat_z = mat3_rowmajor(rot_z9 ? rot_z9 : rot9);
  state.foot_pos_abs = legs(foot_abs12);
  for (int i = 0; i < NUM_LEG; ++i) state.contacts[i] = (contacts >> i) & 1u;
  state.root_euler_d = vec3(ref12);
  state.root_pos_d = vec3(ref12 + 3);
  state.root_lin_vel_d = vec3(ref12 + 6);
  state.root_ang_vel_d = vec3(ref12 + 9);
  if (gains12) {
    state.kp_linear = vec3(gains12);
    state.kd_linear = vec3(gains12 + 3);
    state.kp_angular = vec3(gains12 + 6);
    state.kd_angular = vec3(gains12 + 9);
  }
  state.terrain_pitch_angle = 0;
  Eigen::Matrix<double, 3, NUM_LEG> f;
  f.setZero();
  for (int t = 0; t < ticks; ++t) f = ctrl.compute_grf(state, dt);
  out_legs(f, f_body12);
  if (mpc_states13) for (int i = 0; i < 13; ++i) mpc_states13[i] = state.mpc_states(i);
  if (mpc_states_d) for (int i = 0; i < 13 * PLAN_HORIZON; ++i) mpc_states_d[i] = state.mpc_states_d(i);
  if (root_lin_vel_d_world3) for (int i = 0; i < 3; ++i) root_lin_vel_d_world3[i] = state.root_lin_vel_d_world(i);
  if (root_euler_d3) for (int i = 0; i < 3; ++i) root_euler_d3[i] = state.root_euler_d(i);
  return 0;
}

// A1RobotControl::compute_joint_torques (A1RobotControl.cpp:289-319).  jac: four 3x3 ROW-major blocks (the diagonal blocks of
// j_foot); forces leg-major; tau_prev = state.joint_torques before the call (NaN results keep it).  The first nine calls of a
// fresh controller return zero torques (:293-296), so the controller is ticked past them first.
int ref_joint_torques(const double* f_grf12, const double* f_kin12, const double* jac36, uint32_t contacts, const double* km_foot3,
                      const double* torques_gravity12, const double* tau_prev12, double* tau12) {
  CoutMute mute;
  A1RobotControl ctrl;
  A1CtrlStates state;
  state.j_foot.setZero();
  for (int l = 0; l < NUM_LEG; ++l)
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < 3; ++k) state.j_foot(3 * l + a, 3 * l + k) = jac36[9 * l + 3 * a + k];
  state.foot_forces_grf = legs(f_grf12);
  state.foot_forces_kin = legs(f_kin12);
  for (int i = 0; i < NUM_LEG; ++i) state.contacts[i] = (contacts >> i) & 1u;
  state.km_foot = vec3(km_foot3);
  for (int i = 0; i < 12; ++i) state.torques_gravity(i) = torques_gravity12[i];
  for (int t = 0; t < 9; ++t) ctrl.compute_joint_torques(state);
  for (int i = 0; i < 12; ++i) state.joint_torques(i) = tau_prev12 ? tau_prev12[i] : 0.0;
  ctrl.compute_joint_torques(state);
  for (int i = 0; i < 12; ++i) tau12[i] = state.joint_torques(i);
  return 0;
}

// A1RobotControl::update_plan (A1RobotControl.cpp:148-202), one tick.  gait_counter4 is updated in place.
int ref_update_plan(double counter_per_gait, double counter_per_swing, double control_dt, const double* default_foot_pos12,
                    int movement_mode, double* gait_counter4, const double* gait_counter_speed4, const double* lin_vel3,
                    const double* lin_vel_d3, const double* rot_z9, const double* rot9, const double* root_pos3, double dt,
                    uint32_t* plan_contacts, double* target_rel12, double* target_abs12, double* target_world12) {
  CoutMute mute;
  A1RobotControl ctrl;
  A1CtrlStates state;
  state.counter_per_gait = counter_per_gait;
  state.counter_per_swing = counter_per_swing;
  state.control_dt = control_dt;
  state.default_foot_pos = legs(default_foot_pos12);
  state.movement_mode = movement_mode;
  for (int i = 0; i < 4; ++i) { state.gait_counter(i) = gait_counter4[i]; state.gait_counter_speed(i) = gait_counter_speed4[i]; }
  state.root_lin_vel = vec3(lin_vel3);
  state.root_lin_vel_d = vec3(lin_vel_d3);
  state.root_rot_mat_z = mat3_rowmajor(rot_z9);
  state.root_rot_mat = mat3_rowmajor(rot9);
  state.root_pos = vec3(root_pos3);
  ctrl.update_plan(state, dt);
  for (int i = 0; i < 4; ++i) gait_counter4[i] = state.gait_counter(i);
  uint32_t pc = 0;
  for (int i = 0; i < 4; ++i) pc |= (state.plan_contacts[i] ? 1u : 0u) << i;
  *plan_contacts = pc;
  out_legs(state.foot_pos_target_rel, target_rel12);
  out_legs(state.foot_pos_target_abs, target_abs12);
  out_legs(state.foot_pos_target_world, target_world12);
  return 0;
}

// A1BasicEKF (A1BasicEKF.cpp:5-164) behind an opaque handle: the filter state lives in the reference object
void* ref_ekf_new(int assume_flat_ground) { return new A1BasicEKF(assume_flat_ground != 0); }
void ref_ekf_free(void* h) { delete static_cast<A1BasicEKF*>(h); }
int ref_ekf_init(void* h, const double* foot_pos_rel12, const double* rot9, double* x18, double* P324) {
  A1BasicEKF* ekf = static_cast<A1BasicEKF*>(h);
  A1CtrlStates state;
  state.foot_pos_rel = legs(foot_pos_rel12);
  state.root_rot_mat = mat3_rowmajor(rot9);
  ekf->init_state(state);
  if (x18) for (int i = 0; i < STATE_SIZE; ++i) x18[i] = ekf->x(i);
  out_rowmajor(ekf->P, P324);
  return 0;
}
int ref_ekf_update(void* h, double dt, int movement_mode, const double* imu_acc3, const double* imu_ang_vel3, const double* rot9,
                   const double* foot_pos_rel12, const double* foot_vel_rel12, const double* foot_force4, double* x18, double* P324,
                   double* root_pos3, double* root_lin_vel3, uint32_t* estimated_contacts) {
  A1BasicEKF* ekf = static_cast<A1BasicEKF*>(h);
  A1CtrlStates state;
  state.movement_mode = movement_mode;
  state.imu_acc = vec3(imu_acc3);
  state.imu_ang_vel = vec3(imu_ang_vel3);
  state.root_rot_mat = mat3_rowmajor(rot9);
  state.foot_pos_rel = legs(foot_pos_rel12);
  state.foot_vel_rel = legs(foot_vel_rel12);
  for (int i = 0; i < 4; ++i) state.foot_force(i) = foot_force4[i];
  ekf->update_estimation(state, dt);
  if (x18) for (int i = 0; i < STATE_SIZE; ++i) x18[i] = ekf->x(i);
  out_rowmajor(ekf->P, P324);
  for (int i = 0; i < 3; ++i) { root_pos3[i] = state.root_pos(i); root_lin_vel3[i] = state.root_lin_vel(i); }
  uint32_t ec = 0;
  for (int i = 0; i < 4; ++i) ec |= (state.estimated_contacts[i] ? 1u : 0u) << i;
  *estimated_contacts = ec;
  return 0;
}

// Utils::skew (utils/Utils.cpp:35-41), row-major out
void ref_skew(const double* v3, double* S9) { out_rowmajor(Utils::skew(vec3(v3)), S9); }

}  // extern "C"

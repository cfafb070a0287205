// A1MpcBatch.h -- C++ host shims over the C ABI (include/a1mpc.h) that mirror the reference's class
// surface, so that a1_cpp can swap them in:
//
//   ConvexMpcBatch                 <->  ConvexMpc            (src/a1_cpp/src/ConvexMpc.h:22-94)
//   A1RobotControlBatch::compute_grf <-> A1RobotControl::compute_grf (src/a1_cpp/src/A1RobotControl.h:44)
//   A1CtrlStatesLite               <->  the A1CtrlStates fields the MPC branch reads (A1CtrlStates.h:347-413)
//
// Same method names, argument meaning and "void, no error return" behaviour as the reference (errors
// throw std::runtime_error instead of being ignored).  No Eigen dependency: matrices are plain row-major
// arrays; with Eigen available, Eigen::Map<Eigen::Matrix<double,R,C,Eigen::RowMajor>> views them in place.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/a1mpc.h"

namespace a1mpc_host {

constexpr int NUM_LEG = 4;          // A1Params.h:31
constexpr int MPC_STATE_DIM = 13;   // A1Params.h:27
constexpr int NUM_DOF = 12;         // A1Params.h:34
constexpr int MPC_CONSTRAINT_DIM = 20;  // A1Params.h:28

// The slice of A1CtrlStates that compute_grf's MPC branch reads (A1RobotControl.cpp:452-503).
struct A1CtrlStatesLite {
  double root_euler[3] = {0, 0, 0}, root_pos[3] = {0, 0, 0}, root_ang_vel[3] = {0, 0, 0}, root_lin_vel[3] = {0, 0, 0};
  double root_rot_mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  double foot_pos_abs[12] = {0};                         // 3 x NUM_LEG, row-major like Eigen's operator<< listing
  double root_euler_d[3] = {0, 0, 0}, root_pos_d[3] = {0, 0, 0}, root_lin_vel_d[3] = {0, 0, 0}, root_ang_vel_d[3] = {0, 0, 0};
  bool contacts[NUM_LEG] = {false, false, false, false};
  // written by compute_grf, as the reference writes them into A1CtrlStates (A1RobotControl.cpp:452-488):
  double mpc_states[MPC_STATE_DIM] = {0};              // [euler, pos, ang_vel, lin_vel, -9.8]
  std::vector<double> mpc_states_d;                    // 13 * horizon desired states
  double root_lin_vel_d_world[3] = {0, 0, 0};          // root_rot_mat * root_lin_vel_d
};

struct Handle {
  a1mpc_handle* h = nullptr;
  a1mpc_config cfg;
  Handle(const a1mpc_config& c, int device);
  ~Handle();
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
};

// Batched ConvexMpc: robot b of the batch plays the role of one reference ConvexMpc object.
class ConvexMpcBatch {
 public:
  // q_weights_ (13), r_weights_ (12): ConvexMpc::ConvexMpc(Eigen::VectorXd&, Eigen::VectorXd&) (ConvexMpc.cpp:7)
  ConvexMpcBatch(int batch, const double* q_weights_, const double* r_weights_, int horizon = 10, int device = 0);
  void reset();                                                                   // ConvexMpc.cpp:70
  void calculate_A_mat_c(int b, const double root_euler[3]);                      // ConvexMpc.cpp:110
  void calculate_B_mat_c(int b, double robot_mass, const double a1_trunk_inertia[9], const double root_rot_mat[9],
                         const double foot_pos[12] /*3 x NUM_LEG row-major*/);   // ConvexMpc.cpp:132
  void state_space_discretization(int b, double dt);                              // ConvexMpc.cpp:145
  void store_B_mat_d(int b, int i);               // mpc_solver.B_mat_d_list.block<13,12>(i*13,0) = B_mat_d (A1RobotControl.cpp:513)
  // mpc_states (13) / mpc_states_d (13N) / contacts per robot, then the GPU builds hessian/gradient/bounds for the batch
  void set_states(int b, const double* mpc_states, const double* mpc_states_d, const bool contacts[NUM_LEG]);
  void calculate_qp_mats();                                                       // ConvexMpc.cpp:158 (all robots at once)
  // OsqpEigen::Solver::solve() + getSolution() replacement (A1RobotControl.cpp:522-555): solution(b) has 12N entries
  void solve();

  int batch() const { return B_; }
  int horizon() const { return N_; }
  // public result members, QP-major (ConvexMpc.h:87-93): hessian [B][12N][12N], gradient [B][12N], lb/ub [B][20N]
  std::vector<double> hessian, gradient, lb, ub, solution;
  // the other public members of the reference class (its `private:` is commented out, ConvexMpc.h:37):
  //   A_qp [B][13N][13], B_qp [B][13N][12N] (ConvexMpc.h:77-78; filled by calculate_qp_mats on the GPU),
  //   linear_constraints [20N][12N] (the constant pyramid matrix, ConvexMpc.cpp:46-58; dense row-major here, shared by the batch),
  //   q_weights_mpc [13N], r_weights_mpc [12N] (tiled weights), Q = 2 q_weights_mpc, R = 2 r_weights_mpc (diagonals, ConvexMpc.cpp:20,41)
  std::vector<double> A_qp, B_qp, linear_constraints, q_weights_mpc, r_weights_mpc, Q, R;
  std::vector<int32_t> status;
  // per-robot working matrices, row-major (ConvexMpc.h:68-76)
  std::vector<double> A_mat_c, B_mat_c, A_mat_d, B_mat_d, B_mat_d_list;
  double mu = 0.3, fz_min = 0.0, fz_max = 180.0;   // ConvexMpc.cpp:8, 223-224

 private:
  int B_, N_;
  std::vector<double> x0_, xd_;
  std::vector<uint32_t> contact_;
  Handle handle_;
};

// Batched A1RobotControl::compute_grf, MPC branch: one call = B robots.
class A1RobotControlBatch {
 public:
  // robot_mass, a1_trunk_inertia (row-major 3x3), q_weights (13), r_weights (12): the A1CtrlStates fields that are
  // shared by the whole batch (A1CtrlStates.h:40-60, 365-366)
  A1RobotControlBatch(double robot_mass, const double a1_trunk_inertia[9], const double* q_weights, const double* r_weights,
                      int horizon = 10, int device = 0);
  // foot_forces_grf: 3 x NUM_LEG per robot, row-major ([b][xyz][leg]), body frame -- what compute_grf returns
  // (A1RobotControl.cpp:563).  dt is mpc_dt when use_sim_time == "true" (A1RobotControl.cpp:465-467), else pass 0.0025.
  void compute_grf(const std::vector<A1CtrlStatesLite>& states, double dt, std::vector<std::array<double, 12>>& foot_forces_grf,
                   std::vector<int32_t>* status = nullptr);
  // the same, and like the reference it leaves mpc_states, mpc_states_d and root_lin_vel_d_world in every state (A1RobotControl.cpp:452-488)
  void compute_grf(std::vector<A1CtrlStatesLite>& states, double dt, std::vector<std::array<double, 12>>& foot_forces_grf,
                   std::vector<int32_t>* status = nullptr);
  // The reference's solver object persists and warm-starts every tick (A1RobotControl.h:67, A1RobotControl.cpp:522-538).
  // true: robot b of consecutive compute_grf calls is the same robot, and its previous active faces are tried first
  // (a1mpc_solve_batch_warm; horizon 10).  Off by default: batches of unrelated states gain nothing from it.
  void set_warm_start(bool on) { warm_start_ = on; }

 private:
  a1mpc_config cfg_;
  int device_;
  double dt_ = -1.0;
  Handle* handle_ = nullptr;
  bool warm_start_ = false;
  void* warm_ = nullptr;      // device-resident warm-start state, owned by this object
  size_t warm_B_ = 0;
  std::vector<double> x0_, rot_, foot_, ref_, f_;
  std::vector<uint32_t> contact_;
  std::vector<int32_t> status_;

 public:
  ~A1RobotControlBatch();
};

}  // namespace a1mpc_host

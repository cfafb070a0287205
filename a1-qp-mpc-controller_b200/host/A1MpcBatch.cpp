// A1MpcBatch.cpp -- see A1MpcBatch.h.  Host-side packing only; every QP build/solve runs on the GPU through
// the C ABI.  The small per-robot model matrices (A_c, B_c, Euler discretisation; "negligible" stages in the
// reference) are filled on the host exactly as the caller of ConvexMpc drives them.
#include "A1MpcBatch.h"

#include <cmath>
#include <cstring>

namespace a1mpc_host {

static void check(int rc, const char* what) {
  if (rc != A1MPC_OK) throw std::runtime_error(std::string(what) + ": " + a1mpc_last_error());
}

Handle::Handle(const a1mpc_config& c, int device) : cfg(c) { check(a1mpc_create(&h, &cfg, device), "a1mpc_create"); }
Handle::~Handle() { a1mpc_destroy(h); }

static a1mpc_config make_cfg(const double* q, const double* r, int horizon) {
  a1mpc_config c;
  a1mpc_default_config(&c);
  c.horizon = horizon;
  for (int i = 0; i < 13; ++i) c.q[i] = q[i];
  for (int i = 0; i < 12; ++i) c.r[i] = r[i];
  return c;
}

ConvexMpcBatch::ConvexMpcBatch(int batch, const double* q, const double* r, int horizon, int device)
    : B_(batch), N_(horizon), handle_(make_cfg(q, r, horizon), device) {
  // what the reference's constructor builds (ConvexMpc.cpp:7-68): tiled weights, Q = 2q, R = 2r, the pyramid matrix
  const int N = N_;
  q_weights_mpc.resize(13 * N); r_weights_mpc.resize(12 * N); Q.resize(13 * N); R.resize(12 * N);
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < 13; ++k) { q_weights_mpc[13 * i + k] = q[k]; Q[13 * i + k] = 2 * q[k]; }
    for (int k = 0; k < 12; ++k) { r_weights_mpc[12 * i + k] = r[k]; R[12 * i + k] = 2 * r[k]; }
  }
  const size_t nc = 12 * (size_t)N;
  linear_constraints.assign(20 * (size_t)N * nc, 0.0);
  for (int i = 0; i < NUM_LEG * N; ++i) {
    auto at = [&](int row, int col) -> double& { return linear_constraints[(size_t)row * nc + col]; };
    at(0 + 5 * i, 0 + 3 * i) = 1; at(1 + 5 * i, 0 + 3 * i) = 1; at(2 + 5 * i, 1 + 3 * i) = 1; at(3 + 5 * i, 1 + 3 * i) = 1; at(4 + 5 * i, 2 + 3 * i) = 1;
    at(0 + 5 * i, 2 + 3 * i) = mu; at(1 + 5 * i, 2 + 3 * i) = -mu; at(2 + 5 * i, 2 + 3 * i) = mu; at(3 + 5 * i, 2 + 3 * i) = -mu;
  }
  reset();
}

void ConvexMpcBatch::reset() {
  const size_t B = B_, N = N_;
  A_mat_c.assign(B * 169, 0.0); B_mat_c.assign(B * 156, 0.0); A_mat_d.assign(B * 169, 0.0); B_mat_d.assign(B * 156, 0.0);
  B_mat_d_list.assign(B * 13 * N * 12, 0.0);
  hessian.assign(B * 12 * N * 12 * N, 0.0); gradient.assign(B * 12 * N, 0.0);
  A_qp.assign(B * 13 * N * 13, 0.0); B_qp.assign(B * 13 * N * 12 * N, 0.0);
  lb.assign(B * 20 * N, 0.0); ub.assign(B * 20 * N, 0.0); solution.assign(B * 12 * N, 0.0);
  status.assign(B, 0);
  x0_.assign(B * 13, 0.0); xd_.assign(B * 13 * N, 0.0); contact_.assign(B, 0u);
}

void ConvexMpcBatch::calculate_A_mat_c(int b, const double root_euler[3]) {
  double* A = &A_mat_c[(size_t)b * 169];
  const double cy = std::cos(root_euler[2]), sy = std::sin(root_euler[2]);
  A[0 * 13 + 6] = cy;  A[0 * 13 + 7] = sy; A[0 * 13 + 8] = 0;
  A[1 * 13 + 6] = -sy; A[1 * 13 + 7] = cy; A[1 * 13 + 8] = 0;
  A[2 * 13 + 6] = 0;   A[2 * 13 + 7] = 0;  A[2 * 13 + 8] = 1;
  for (int k = 0; k < 3; ++k) A[(3 + k) * 13 + 9 + k] = 1;
  A[11 * 13 + 12] = 1;
}

static void inv3(const double* m, double* o) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  const double id = 1.0 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

void ConvexMpcBatch::calculate_B_mat_c(int b, double mass, const double I[9], const double R[9], const double foot[12]) {
  double* Bc = &B_mat_c[(size_t)b * 156];
  double t[9], Iw[9], Iwi[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = R[3 * i] * I[j] + R[3 * i + 1] * I[3 + j] + R[3 * i + 2] * I[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Iw[3 * i + j] = t[3 * i] * R[3 * j] + t[3 * i + 1] * R[3 * j + 1] + t[3 * i + 2] * R[3 * j + 2];
  inv3(Iw, Iwi);
  for (int leg = 0; leg < NUM_LEG; ++leg) {
    const double v[3] = {foot[0 * 4 + leg], foot[1 * 4 + leg], foot[2 * 4 + leg]};
    const double S[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};  // Utils::skew
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        Bc[(6 + a) * 12 + 3 * leg + c] = Iwi[3 * a] * S[c] + Iwi[3 * a + 1] * S[3 + c] + Iwi[3 * a + 2] * S[6 + c];
        Bc[(9 + a) * 12 + 3 * leg + c] = (a == c) ? 1.0 / mass : 0.0;
      }
  }
}

void ConvexMpcBatch::state_space_discretization(int b, double dt) {
  const double* Ac = &A_mat_c[(size_t)b * 169];
  const double* Bc = &B_mat_c[(size_t)b * 156];
  double* Ad = &A_mat_d[(size_t)b * 169];
  double* Bd = &B_mat_d[(size_t)b * 156];
  for (int i = 0; i < 13; ++i)
    for (int j = 0; j < 13; ++j) Ad[13 * i + j] = (i == j ? 1.0 : 0.0) + Ac[13 * i + j] * dt;
  for (int k = 0; k < 156; ++k) Bd[k] = Bc[k] * dt;
}

void ConvexMpcBatch::store_B_mat_d(int b, int i) {
  std::memcpy(&B_mat_d_list[((size_t)b * 13 * N_ + 13 * i) * 12], &B_mat_d[(size_t)b * 156], sizeof(double) * 156);
}

void ConvexMpcBatch::set_states(int b, const double* mpc_states, const double* mpc_states_d, const bool contacts[NUM_LEG]) {
  std::memcpy(&x0_[(size_t)b * 13], mpc_states, sizeof(double) * 13);
  std::memcpy(&xd_[(size_t)b * 13 * N_], mpc_states_d, sizeof(double) * 13 * N_);
  uint32_t m = 0;
  for (int i = 0; i < NUM_LEG; ++i) m |= contacts[i] ? (1u << i) : 0u;
  contact_[b] = m;
}

void ConvexMpcBatch::calculate_qp_mats() {
  check(a1mpc_qp_rollout_batch(handle_.h, B_, A_mat_d.data(), B_mat_d_list.data(), x0_.data(), xd_.data(), A_qp.data(), B_qp.data(),
                               hessian.data(), gradient.data()),
        "a1mpc_qp_rollout_batch");
  // bounds, ConvexMpc.cpp:223-245 (host: 40N scalars per robot)
  const double INFTY = 1e30;  // OsqpEigen::INFTY
  for (int b = 0; b < B_; ++b)
    for (int k = 0; k < 4 * N_; ++k) {
      const double c = ((contact_[b] >> (k % 4)) & 1u) ? 1.0 : 0.0;
      double* l = &lb[((size_t)b * 4 * N_ + k) * 5];
      double* u = &ub[((size_t)b * 4 * N_ + k) * 5];
      l[0] = 0; u[0] = INFTY; l[1] = -INFTY; u[1] = 0; l[2] = 0; u[2] = INFTY; l[3] = -INFTY; u[3] = 0;
      l[4] = fz_min * c; u[4] = fz_max * c;
    }
}

void ConvexMpcBatch::solve() {
  check(a1mpc_solve_dense_batch(handle_.h, B_, hessian.data(), gradient.data(), contact_.data(), solution.data(), status.data()),
        "a1mpc_solve_dense_batch");
}

A1RobotControlBatch::A1RobotControlBatch(double mass, const double I[9], const double* q, const double* r, int horizon, int device)
    : device_(device) {
  cfg_ = make_cfg(q, r, horizon);
  cfg_.mass = mass;
  for (int i = 0; i < 9; ++i) cfg_.inertia[i] = I[i];
}

A1RobotControlBatch::~A1RobotControlBatch() {
  if (warm_ && handle_) a1mpc_device_free(handle_->h, warm_);
  delete handle_;
}

void A1RobotControlBatch::compute_grf(const std::vector<A1CtrlStatesLite>& st, double dt, std::vector<std::array<double, 12>>& out,
                                      std::vector<int32_t>* status) {
  const size_t B = st.size();
  if (B == 0) { out.clear(); return; }
  if (!handle_ || dt != dt_) {  // mpc_dt is a handle-level constant (A1RobotControl.cpp:462-467)
    if (warm_ && handle_) a1mpc_device_free(handle_->h, warm_);
    warm_ = nullptr; warm_B_ = 0;
    delete handle_;
    handle_ = nullptr;
    cfg_.dt = dt;
    handle_ = new Handle(cfg_, device_);
    dt_ = dt;
  }
  x0_.resize(12 * B); rot_.resize(9 * B); foot_.resize(12 * B); ref_.resize(9 * B); f_.resize(12 * B);
  contact_.resize(B); status_.resize(B);
  for (size_t b = 0; b < B; ++b) {  // batch-major SoA packing (include/a1mpc.h)
    const A1CtrlStatesLite& s = st[b];
    for (int k = 0; k < 3; ++k) {
      x0_[(0 + k) * B + b] = s.root_euler[k]; x0_[(3 + k) * B + b] = s.root_pos[k];
      x0_[(6 + k) * B + b] = s.root_ang_vel[k]; x0_[(9 + k) * B + b] = s.root_lin_vel[k];
    }
    for (int k = 0; k < 9; ++k) rot_[k * B + b] = s.root_rot_mat[k];
    for (int leg = 0; leg < 4; ++leg)
      for (int a = 0; a < 3; ++a) foot_[(3 * leg + a) * B + b] = s.foot_pos_abs[a * 4 + leg];
    ref_[0 * B + b] = s.root_euler_d[0]; ref_[1 * B + b] = s.root_euler_d[1];
    for (int k = 0; k < 3; ++k) { ref_[(2 + k) * B + b] = s.root_ang_vel_d[k]; ref_[(5 + k) * B + b] = s.root_lin_vel_d[k]; }
    ref_[8 * B + b] = s.root_pos_d[2];
    uint32_t m = 0;
    for (int i = 0; i < 4; ++i) m |= s.contacts[i] ? (1u << i) : 0u;
    contact_[b] = m;
  }
  a1mpc_inputs in{x0_.data(), rot_.data(), foot_.data(), ref_.data(), contact_.data(), B};
  a1mpc_outputs o{f_.data(), status_.data(), nullptr, nullptr, B};
  if (warm_start_) {
    if (warm_B_ != B) {   // a different batch: the slots mean different robots, start cold
      if (warm_) a1mpc_device_free(handle_->h, warm_);
      warm_ = nullptr;
      check(a1mpc_device_alloc(handle_->h, a1mpc_warm_bytes(handle_->h, (int)B), &warm_), "a1mpc_device_alloc");
      check(a1mpc_warm_reset(handle_->h, warm_, (int)B), "a1mpc_warm_reset");
      warm_B_ = B;
    }
    check(a1mpc_solve_batch_warm(handle_->h, (int)B, &in, &o, warm_, 0), "a1mpc_solve_batch_warm");
  } else {
    check(a1mpc_solve_batch(handle_->h, (int)B, &in, &o), "a1mpc_solve_batch");
  }
  out.resize(B);
  for (size_t b = 0; b < B; ++b)
    for (int leg = 0; leg < 4; ++leg)
      for (int a = 0; a < 3; ++a) out[b][a * 4 + leg] = f_[(3 * leg + a) * B + b];  // 3 x NUM_LEG, row-major
  if (status) *status = status_;
}

void A1RobotControlBatch::compute_grf(std::vector<A1CtrlStatesLite>& st, double dt, std::vector<std::array<double, 12>>& out,
                                      std::vector<int32_t>* status) {
  compute_grf(static_cast<const std::vector<A1CtrlStatesLite>&>(st), dt, out, status);
  // state write-backs of the reference (A1RobotControl.cpp:452-488); the same scalars the pack kernel forms on the device
  const int N = cfg_.horizon;
  for (A1CtrlStatesLite& s : st) {
    for (int k = 0; k < 3; ++k) {
      s.mpc_states[k] = s.root_euler[k]; s.mpc_states[3 + k] = s.root_pos[k];
      s.mpc_states[6 + k] = s.root_ang_vel[k]; s.mpc_states[9 + k] = s.root_lin_vel[k];
    }
    s.mpc_states[12] = -9.8;
    for (int a = 0; a < 3; ++a)
      s.root_lin_vel_d_world[a] = s.root_rot_mat[3 * a] * s.root_lin_vel_d[0] + s.root_rot_mat[3 * a + 1] * s.root_lin_vel_d[1] + s.root_rot_mat[3 * a + 2] * s.root_lin_vel_d[2];
    s.mpc_states_d.resize(13 * (size_t)N);
    for (int i = 0; i < N; ++i) {
      double* d = &s.mpc_states_d[13 * (size_t)i];
      d[0] = s.root_euler_d[0]; d[1] = s.root_euler_d[1]; d[2] = s.root_euler[2] + s.root_ang_vel_d[2] * dt * (i + 1);
      d[3] = s.root_pos[0] + s.root_lin_vel_d_world[0] * dt * (i + 1); d[4] = s.root_pos[1] + s.root_lin_vel_d_world[1] * dt * (i + 1);
      d[5] = s.root_pos_d[2];
      d[6] = s.root_ang_vel_d[0]; d[7] = s.root_ang_vel_d[1]; d[8] = s.root_ang_vel_d[2];
      d[9] = s.root_lin_vel_d_world[0]; d[10] = s.root_lin_vel_d_world[1]; d[11] = 0; d[12] = -9.8;
    }
  }
}

}  // namespace a1mpc_host

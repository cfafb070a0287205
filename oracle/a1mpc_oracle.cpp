// oracle/a1mpc_oracle.cpp -- TEST INFRASTRUCTURE, not product code.
//
// CPU restatement of the reference hot path (ShuoYangRobotics/A1-QP-MPC-Controller), used only
// as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing under a1-qp-mpc-controller_b200/ may include, link or call this file.
//
// PARITY PINNED to the reference's own compiled code (everything but the third-party OSQP iteration):
// the reference's ConvexMpc.cpp, A1RobotControl.cpp, A1BasicEKF.cpp, utils/Utils.cpp and legKinematics/A1Kinematics.cpp compile
// UNMODIFIED, from where they lie under /root/reference, against the header stand-ins of oracle/ref_shim/ (a small API-compatible
// subset of Eigen, plus inert OsqpEigen / ROS stubs; Eigen, OSQP and ROS are not installable offline) -- `make -C oracle ref` ->
// oracle/_ref/libref_mpc.so, libref_kin.so and ref_test_mpc (the reference's test/test_mpc.cpp, main() and all).  Against that
// build this file agrees to <= 2e-15 relative on H, g, A_qp, B_qp, x_d, bit-exactly on lb, ub, the pyramid matrix and x0, on the
// single-step QP set-up, compute_joint_torques, update_plan, and to 1e-11 over 30 ticks of A1BasicEKF (tests/test_ref_pin.py);
// tests/golden/convexmpc_v1.npz carries vectors generated from that build to the GPU box (tests/golden/make_ref_golden.py).
// What cannot be pinned: OSQP's ADMM iterates (github.com/oxfordcontrol/osqp, unpinned master ~v0.6.2, docker/Dockerfile:77-83,
// behind osqp-eigen 0.6.3, docker/Dockerfile:91-98 -- neither vendored).  The reference holds no golden vector for the solve
// (test/test_mpc.cpp:157-161 prints and returns 0).  The QP is strictly convex (2r > 0), so "what OSQP returns when run to
// convergence" is the unique optimum; it is pinned by (1) two independent solvers below (OSQP-algorithm restatement at eps 1e-11
// vs. long-double exact solver) and (2) a KKT certificate evaluated on the LITERAL 12N-variable problem -- since round 2 also on
// the REFERENCE-built H, g, Ac, lb, ub (tests/test_gpu_ref_pin.py) -- which proves optimality independently of the method.
//
// Dependency-free C++17.  Literal where the reference is literal: dense rollout, dense
// B_qp^T Q B_qp, dense constraint matrix, OSQP-style ADMM.  No closed forms, no swing elimination
// in the reference-faithful path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/a1mpc.h"

namespace {

typedef long double ld;
const double OSQP_INFTY = 1e30;  // OsqpEigen::INFTY (osqp's OSQP_INFTY), used by ConvexMpc.cpp:230-237

// ------------------------------------------------------------------------------------------
// tiny dense helpers (row-major)
// ------------------------------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

static void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j];
      C[3 * i + j] = s;
    }
}
static void mat3_T(const double* A, double* B) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[3 * j + i] = A[3 * i + j];
}
static void mat3_inv(const double* m, double* o) {
  double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
               m[2] * (m[3] * m[7] - m[4] * m[6]);
  double id = 1.0 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// Utils::skew, utils/Utils.cpp:35-41
static void skew(const double* v, double* S) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

// ------------------------------------------------------------------------------------------
// ConvexMpc restated (ConvexMpc.cpp:7-260), horizon N at run time
// ------------------------------------------------------------------------------------------
struct ConvexMpcRestated {
  int N;
  double mu, fz_min, fz_max;
  std::vector<double> q_weights_mpc, r_weights_mpc;  // tiled, un-doubled (ConvexMpc.cpp:16-19, 37-40)
  Mat linear_constraints;                            // 20N x 12N (ConvexMpc.cpp:46-58)
  Mat A_mat_c, B_mat_c, A_mat_d, B_mat_d, B_mat_d_list, A_qp, B_qp, hessian, QB_;
  std::vector<double> gradient, lb, ub;

  ConvexMpcRestated(int N_, const double* q, const double* r, double mu_, double fzmin, double fzmax)
      : N(N_), mu(mu_), fz_min(fzmin), fz_max(fzmax) {
    q_weights_mpc.resize(13 * N);
    r_weights_mpc.resize(12 * N);
    for (int i = 0; i < N; ++i) {
      for (int k = 0; k < 13; ++k) q_weights_mpc[13 * i + k] = q[k];
      for (int k = 0; k < 12; ++k) r_weights_mpc[12 * i + k] = r[k];
    }
    linear_constraints = Mat(20 * N, 12 * N);
    for (int i = 0; i < 4 * N; ++i) {
      linear_constraints(0 + 5 * i, 0 + 3 * i) = 1;
      linear_constraints(1 + 5 * i, 0 + 3 * i) = 1;
      linear_constraints(2 + 5 * i, 1 + 3 * i) = 1;
      linear_constraints(3 + 5 * i, 1 + 3 * i) = 1;
      linear_constraints(4 + 5 * i, 2 + 3 * i) = 1;
      linear_constraints(0 + 5 * i, 2 + 3 * i) = mu;
      linear_constraints(1 + 5 * i, 2 + 3 * i) = -mu;
      linear_constraints(2 + 5 * i, 2 + 3 * i) = mu;
      linear_constraints(3 + 5 * i, 2 + 3 * i) = -mu;
    }
    reset();
  }
  void reinit(const double* q, const double* r) {  // the per-tick part of the constructor (ConvexMpc.cpp:16-58)
    for (int i = 0; i < N; ++i) {
      for (int k = 0; k < 13; ++k) q_weights_mpc[13 * i + k] = q[k];
      for (int k = 0; k < 12; ++k) r_weights_mpc[12 * i + k] = r[k];
    }
    for (int i = 0; i < 4 * N; ++i) {
      linear_constraints(0 + 5 * i, 0 + 3 * i) = 1; linear_constraints(1 + 5 * i, 0 + 3 * i) = 1;
      linear_constraints(2 + 5 * i, 1 + 3 * i) = 1; linear_constraints(3 + 5 * i, 1 + 3 * i) = 1;
      linear_constraints(4 + 5 * i, 2 + 3 * i) = 1;
      linear_constraints(0 + 5 * i, 2 + 3 * i) = mu; linear_constraints(1 + 5 * i, 2 + 3 * i) = -mu;
      linear_constraints(2 + 5 * i, 2 + 3 * i) = mu; linear_constraints(3 + 5 * i, 2 + 3 * i) = -mu;
    }
  }
  void reset() {  // ConvexMpc.cpp:70-108 (setZero on fixed-size members: no reallocation)
    auto z = [](Mat& m, int r, int c) {
      if (m.r != r || m.c != c) m = Mat(r, c);
      else std::fill(m.a.begin(), m.a.end(), 0.0);
    };
    z(A_mat_c, 13, 13); z(B_mat_c, 13, 12); z(A_mat_d, 13, 13); z(B_mat_d, 13, 12);
    z(B_mat_d_list, 13 * N, 12); z(A_qp, 13 * N, 13); z(B_qp, 13 * N, 12 * N); z(hessian, 12 * N, 12 * N);
    gradient.assign(12 * N, 0.0); lb.assign(20 * N, 0.0); ub.assign(20 * N, 0.0);
  }
  void calculate_A_mat_c(const double* root_euler) {  // ConvexMpc.cpp:110-130
    double cy = std::cos(root_euler[2]), sy = std::sin(root_euler[2]);
    A_mat_c(0, 6) = cy;  A_mat_c(0, 7) = sy; A_mat_c(0, 8) = 0;
    A_mat_c(1, 6) = -sy; A_mat_c(1, 7) = cy; A_mat_c(1, 8) = 0;
    A_mat_c(2, 6) = 0;   A_mat_c(2, 7) = 0;  A_mat_c(2, 8) = 1;
    for (int k = 0; k < 3; ++k) A_mat_c(3 + k, 9 + k) = 1;
    A_mat_c(11, 12) = 1;
  }
  void calculate_B_mat_c(double mass, const double* inertia, const double* R, const double* foot /*3x4 row-major*/) {
    // ConvexMpc.cpp:132-143
    double RT[9], t[9], Iw[9], Iwinv[9];
    mat3_T(R, RT);
    mat3_mul(R, inertia, t);
    mat3_mul(t, RT, Iw);
    for (int i = 0; i < 4; ++i) {
      mat3_inv(Iw, Iwinv);  // the reference inverts once per leg
      double v[3] = {foot[0 * 4 + i], foot[1 * 4 + i], foot[2 * 4 + i]}, S[9], M[9];
      skew(v, S);
      mat3_mul(Iwinv, S, M);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          B_mat_c(6 + a, 3 * i + b) = M[3 * a + b];
          B_mat_c(9 + a, 3 * i + b) = (a == b) ? (1 / mass) : 0.0;
        }
    }
  }
  void state_space_discretization(double dt) {  // ConvexMpc.cpp:145-156 (forward Euler)
    for (int i = 0; i < 13; ++i)
      for (int j = 0; j < 13; ++j) A_mat_d(i, j) = (i == j ? 1.0 : 0.0) + A_mat_c(i, j) * dt;
    for (int i = 0; i < 13; ++i)
      for (int j = 0; j < 12; ++j) B_mat_d(i, j) = B_mat_c(i, j) * dt;
  }
  void store_B(int i) {  // A1RobotControl.cpp:513
    for (int a = 0; a < 13; ++a)
      for (int b = 0; b < 12; ++b) B_mat_d_list(13 * i + a, b) = B_mat_d(a, b);
  }
  // ConvexMpc.cpp:158-245
  void calculate_qp_mats(const double* mpc_states, const double* mpc_states_d, const bool* contacts) {
    const int nx = 13, nu = 12;
    // rollout :184-202
    for (int i = 0; i < N; ++i) {
      if (i == 0) {
        for (int a = 0; a < nx; ++a)
          for (int b = 0; b < nx; ++b) A_qp(a, b) = A_mat_d(a, b);
      } else {
        for (int a = 0; a < nx; ++a)
          for (int b = 0; b < nx; ++b) {
            double s = 0;
            for (int k = 0; k < nx; ++k) s += A_qp(nx * (i - 1) + a, k) * A_mat_d(k, b);
            A_qp(nx * i + a, b) = s;
          }
      }
      for (int j = 0; j < i + 1; ++j) {
        if (i - j == 0) {
          for (int a = 0; a < nx; ++a)
            for (int b = 0; b < nu; ++b) B_qp(nx * i + a, nu * j + b) = B_mat_d_list(nx * j + a, b);
        } else {
          for (int a = 0; a < nx; ++a)
            for (int b = 0; b < nu; ++b) {
              double s = 0;
              for (int k = 0; k < nx; ++k) s += A_qp(nx * (i - j - 1) + a, k) * B_mat_d_list(nx * j + k, b);
              B_qp(nx * i + a, nu * j + b) = s;
            }
        }
      }
    }
    // hessian :207-211   dense_hessian = B_qp^T * Q * B_qp ; += R   with Q = 2 q, R = 2 r
    const int n = nu * N, ns = nx * N;
    if (QB_.r != ns || QB_.c != n) QB_ = Mat(ns, n);
    Mat& QB = QB_;
    for (int i = 0; i < ns; ++i) {
      double w = 2 * q_weights_mpc[i];
      for (int j = 0; j < n; ++j) QB(i, j) = w * B_qp(i, j);
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) hessian(i, j) = 0;
    for (int k = 0; k < ns; ++k) {
      const double* bk = &B_qp.a[(size_t)k * n];
      const double* qk = &QB.a[(size_t)k * n];
      for (int i = 0; i < n; ++i) {
        double bi = bk[i];
        if (bi == 0.0) continue;
        double* hi = &hessian.a[(size_t)i * n];
        for (int j = 0; j < n; ++j) hi[j] += bi * qk[j];
      }
    }
    for (int i = 0; i < n; ++i) hessian(i, i) += 2 * r_weights_mpc[i];
    // gradient :215-217
    std::vector<double> tmp(ns);
    for (int i = 0; i < ns; ++i) {
      double s = 0;
      for (int k = 0; k < nx; ++k) s += A_qp(i, k) * mpc_states[k];
      tmp[i] = (s - mpc_states_d[i]) * 2 * q_weights_mpc[i];
    }
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int i = 0; i < ns; ++i) s += B_qp(i, j) * tmp[i];
      gradient[j] = s;
    }
    // bounds :223-245
    double lb1[20], ub1[20];
    for (int i = 0; i < 4; ++i) {
      double c = contacts[i] ? 1.0 : 0.0;
      lb1[5 * i + 0] = 0;           ub1[5 * i + 0] = OSQP_INFTY;
      lb1[5 * i + 1] = -OSQP_INFTY; ub1[5 * i + 1] = 0;
      lb1[5 * i + 2] = 0;           ub1[5 * i + 2] = OSQP_INFTY;
      lb1[5 * i + 3] = -OSQP_INFTY; ub1[5 * i + 3] = 0;
      lb1[5 * i + 4] = fz_min * c;  ub1[5 * i + 4] = fz_max * c;
    }
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < 20; ++k) { lb[20 * i + k] = lb1[k]; ub[20 * i + k] = ub1[k]; }
  }
};

// one robot's record, unpacked from the SoA batch
struct RobotState {
  double euler[3], pos[3], ang_vel[3], lin_vel[3];
  double R[9];
  double foot[12];  // 3x4 row-major like Eigen's operator<< listing: foot[a*4 + leg]
  double euler_d01[2], ang_vel_d[3], lin_vel_d[3], pos_d_z;
  bool contacts[4];
};
static void unpack(const a1mpc_inputs* in, int b, RobotState* s) {
  size_t ld_ = in->ld;
  for (int k = 0; k < 3; ++k) {
    s->euler[k] = in->x0[(0 + k) * ld_ + b];
    s->pos[k] = in->x0[(3 + k) * ld_ + b];
    s->ang_vel[k] = in->x0[(6 + k) * ld_ + b];
    s->lin_vel[k] = in->x0[(9 + k) * ld_ + b];
  }
  for (int k = 0; k < 9; ++k) s->R[k] = in->rot[k * ld_ + b];
  for (int leg = 0; leg < 4; ++leg)
    for (int a = 0; a < 3; ++a) s->foot[a * 4 + leg] = in->foot[(3 * leg + a) * ld_ + b];
  s->euler_d01[0] = in->ref[0 * ld_ + b];
  s->euler_d01[1] = in->ref[1 * ld_ + b];
  for (int k = 0; k < 3; ++k) {
    s->ang_vel_d[k] = in->ref[(2 + k) * ld_ + b];
    s->lin_vel_d[k] = in->ref[(5 + k) * ld_ + b];
  }
  s->pos_d_z = in->ref[8 * ld_ + b];
  for (int i = 0; i < 4; ++i) s->contacts[i] = (in->contact[b] >> i) & 1u;
}

// A1RobotControl::compute_grf, MPC branch up to calculate_qp_mats (A1RobotControl.cpp:447-518)
static void drive_convex_mpc(ConvexMpcRestated& mpc, const a1mpc_config& cfg, const RobotState& st,
                             std::vector<double>& mpc_states, std::vector<double>& mpc_states_d) {
  const int N = cfg.horizon;
  mpc.reset();
  mpc_states.assign(13, 0.0);
  for (int k = 0; k < 3; ++k) {
    mpc_states[k] = st.euler[k];
    mpc_states[3 + k] = st.pos[k];
    mpc_states[6 + k] = st.ang_vel[k];
    mpc_states[9 + k] = st.lin_vel[k];
  }
  mpc_states[12] = -9.8;
  double mpc_dt = cfg.dt;
  double vdw[3];
  for (int a = 0; a < 3; ++a)
    vdw[a] = st.R[3 * a + 0] * st.lin_vel_d[0] + st.R[3 * a + 1] * st.lin_vel_d[1] + st.R[3 * a + 2] * st.lin_vel_d[2];
  mpc_states_d.assign(13 * N, 0.0);
  for (int i = 0; i < N; ++i) {  // :472-488
    double* d = &mpc_states_d[13 * i];
    d[0] = st.euler_d01[0];
    d[1] = st.euler_d01[1];
    d[2] = st.euler[2] + st.ang_vel_d[2] * mpc_dt * (i + 1);
    d[3] = st.pos[0] + vdw[0] * mpc_dt * (i + 1);
    d[4] = st.pos[1] + vdw[1] * mpc_dt * (i + 1);
    d[5] = st.pos_d_z;
    d[6] = st.ang_vel_d[0];
    d[7] = st.ang_vel_d[1];
    d[8] = st.ang_vel_d[2];
    d[9] = vdw[0];
    d[10] = vdw[1];
    d[11] = 0;
    d[12] = -9.8;
  }
  mpc.calculate_A_mat_c(st.euler);
  for (int i = 0; i < N; ++i) {  // :498-514, same inputs every i
    mpc.calculate_B_mat_c(cfg.mass, cfg.inertia, st.R, st.foot);
    mpc.state_space_discretization(mpc_dt);
    mpc.store_B(i);
  }
  mpc.calculate_qp_mats(mpc_states.data(), mpc_states_d.data(), st.contacts);
}

// ------------------------------------------------------------------------------------------
// OSQP algorithm restated (osqp 0.6.x defaults).  Dense P, row-sparse A.
// The KKT system [P+sigma I, A'; A, -1/rho] is solved in its reduced form
// (P + sigma I + A' diag(rho) A) x = rhs_x + A' diag(rho) rhs_z, which is the same linear map.
// ------------------------------------------------------------------------------------------
struct OsqpSettings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6, eps_abs = 1e-3, eps_rel = 1e-3;
  int max_iter = 4000, scaling = 10, check_termination = 25, adaptive_rho = 1, adaptive_rho_interval = 25;
  double adaptive_rho_tolerance = 5.0;
};
struct OsqpInfo { int iter = 0; int status = 0; /*1 solved, 2 max_iter*/ double pri_res = 0, dua_res = 0; int rho_updates = 0; };

struct SparseRows {  // A in row lists
  int m = 0, n = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
  void from_dense(const Mat& A) {
    m = A.r; n = A.c; ptr.assign(1, 0); idx.clear(); val.clear();
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < n; ++j)
        if (A(i, j) != 0.0) { idx.push_back(j); val.push_back(A(i, j)); }
      ptr.push_back((int)idx.size());
    }
  }
};

static bool chol_factor(std::vector<double>& K, int n) {  // in place lower, row-major
  for (int j = 0; j < n; ++j) {
    double d = K[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    K[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = K[(size_t)i * n + j];
      const double* ri = &K[(size_t)i * n];
      const double* rj = &K[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      K[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
static void chol_solve(const std::vector<double>& L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    const double* ri = &L[(size_t)i * n];
    for (int k = 0; k < i; ++k) s -= ri[k] * b[k];
    b[i] = s / ri[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}
static double norm_inf(const double* v, int n) {
  double m = 0;
  for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(v[i]));
  return m;
}
static void limit_scaling(double* v, int n) {
  for (int i = 0; i < n; ++i) {
    v[i] = v[i] < 1e-4 ? 1.0 : v[i];
    v[i] = v[i] > 1e4 ? 1e4 : v[i];
  }
}

static void osqp_restated_solve(int n, int m, const Mat& P_in, const double* q_in, const SparseRows& A_in,
                                const double* l_in, const double* u_in, const OsqpSettings& st,
                                double* x_out, OsqpInfo* info) {
  // ---- scale_data (osqp/src/scaling.c) ----
  Mat P = P_in;
  SparseRows A = A_in;
  std::vector<double> q(q_in, q_in + n), l(l_in, l_in + m), u(u_in, u_in + m);
  std::vector<double> D(n, 1.0), E(m, 1.0), Dt(n), Et(m);
  double c = 1.0;
  for (int it = 0; it < st.scaling; ++it) {
    for (int j = 0; j < n; ++j) Dt[j] = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Dt[j] = std::max(Dt[j], std::fabs(P(i, j)));
    for (int i = 0; i < m; ++i) {
      double rn = 0;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        rn = std::max(rn, std::fabs(A.val[k]));
        Dt[A.idx[k]] = std::max(Dt[A.idx[k]], std::fabs(A.val[k]));
      }
      Et[i] = rn;
    }
    limit_scaling(Dt.data(), n);
    limit_scaling(Et.data(), m);
    for (int j = 0; j < n; ++j) Dt[j] = 1.0 / std::sqrt(Dt[j]);
    for (int i = 0; i < m; ++i) Et[i] = 1.0 / std::sqrt(Et[i]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) P(i, j) *= Dt[i] * Dt[j];
    for (int i = 0; i < m; ++i)
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) A.val[k] *= Et[i] * Dt[A.idx[k]];
    for (int j = 0; j < n; ++j) { q[j] *= Dt[j]; D[j] *= Dt[j]; }
    for (int i = 0; i < m; ++i) E[i] *= Et[i];
    // cost scaling
    double mean = 0;
    for (int j = 0; j < n; ++j) {
      double cn = 0;
      for (int i = 0; i < n; ++i) cn = std::max(cn, std::fabs(P(i, j)));
      mean += cn;
    }
    mean /= n;
    double nq = norm_inf(q.data(), n);
    limit_scaling(&nq, 1);
    double ct = std::max(mean, nq);
    limit_scaling(&ct, 1);
    ct = 1.0 / ct;
    for (auto& v : P.a) v *= ct;
    for (auto& v : q) v *= ct;
    c *= ct;
  }
  for (int i = 0; i < m; ++i) { l[i] *= E[i]; u[i] *= E[i]; }
  const double cinv = 1.0 / c;

  // ---- rho vector (osqp/src/auxil.c set_rho_vec) ----
  std::vector<int> ctype(m);
  std::vector<double> rho_vec(m);
  double rho = st.rho;
  auto set_rho_vec = [&]() {
    for (int i = 0; i < m; ++i) {
      if (l[i] < -OSQP_INFTY * 1e-4 && u[i] > OSQP_INFTY * 1e-4) { ctype[i] = -1; rho_vec[i] = 1e-6; }
      else if (u[i] - l[i] < 1e-4) { ctype[i] = 1; rho_vec[i] = 1e3 * rho; }
      else { ctype[i] = 0; rho_vec[i] = rho; }
    }
  };
  set_rho_vec();
  std::vector<double> K((size_t)n * n);
  auto factor = [&]() {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) K[(size_t)i * n + j] = P(i, j) + (i == j ? st.sigma : 0.0);
    for (int i = 0; i < m; ++i)
      for (int k1 = A.ptr[i]; k1 < A.ptr[i + 1]; ++k1)
        for (int k2 = A.ptr[i]; k2 < A.ptr[i + 1]; ++k2)
          K[(size_t)A.idx[k1] * n + A.idx[k2]] += rho_vec[i] * A.val[k1] * A.val[k2];
    chol_factor(K, n);
  };
  factor();

  std::vector<double> x(n, 0.0), z(m, 0.0), y(m, 0.0), xp(n), zp(m), xt(n), zt(m), Ax(m), Px(n), Aty(n), rhs(n);
  auto mulA = [&](const double* v, double* out) {
    for (int i = 0; i < m; ++i) {
      double s = 0;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += A.val[k] * v[A.idx[k]];
      out[i] = s;
    }
  };
  auto mulAt = [&](const double* v, double* out) {
    for (int j = 0; j < n; ++j) out[j] = 0;
    for (int i = 0; i < m; ++i)
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) out[A.idx[k]] += A.val[k] * v[i];
  };
  auto mulP = [&](const double* v, double* out) {
    for (int i = 0; i < n; ++i) {
      double s = 0;
      const double* pi = &P.a[(size_t)i * n];
      for (int j = 0; j < n; ++j) s += pi[j] * v[j];
      out[i] = s;
    }
  };
  double pri_res = 0, dua_res = 0;
  auto residuals = [&](bool scaled, double* eps_pri, double* eps_dua) {
    // update_info + check_termination thresholds (osqp/src/auxil.c compute_pri_res/compute_dua_res/...)
    mulA(x.data(), Ax.data());
    mulP(x.data(), Px.data());
    mulAt(y.data(), Aty.data());
    double pr = 0, nAx = 0, nz = 0;
    for (int i = 0; i < m; ++i) {
      double s = scaled ? 1.0 : 1.0 / E[i];
      pr = std::max(pr, std::fabs(s * (Ax[i] - z[i])));
      nAx = std::max(nAx, std::fabs(s * Ax[i]));
      nz = std::max(nz, std::fabs(s * z[i]));
    }
    double dr = 0, nPx = 0, nAty = 0, nq = 0;
    for (int j = 0; j < n; ++j) {
      double s = scaled ? 1.0 : 1.0 / D[j];
      dr = std::max(dr, std::fabs(s * (Px[j] + q[j] + Aty[j])));
      nPx = std::max(nPx, std::fabs(s * Px[j]));
      nAty = std::max(nAty, std::fabs(s * Aty[j]));
      nq = std::max(nq, std::fabs(s * q[j]));
    }
    if (!scaled) { dr *= cinv; nPx *= cinv; nAty *= cinv; nq *= cinv; }
    pri_res = pr; dua_res = dr;
    if (eps_pri) *eps_pri = st.eps_abs + st.eps_rel * std::max(nAx, nz);
    if (eps_dua) *eps_dua = st.eps_abs + st.eps_rel * std::max(std::max(nPx, nAty), nq);
    return std::make_pair(std::max(nAx, nz), std::max(std::max(nPx, nAty), nq));
  };
  int iter;
  bool solved = false;
  for (iter = 1; iter <= st.max_iter; ++iter) {
    xp = x; zp = z;
    // update_xz_tilde
    for (int j = 0; j < n; ++j) rhs[j] = st.sigma * xp[j] - q[j];
    for (int i = 0; i < m; ++i) {
      double rz = zp[i] - y[i] / rho_vec[i];
      double w = rho_vec[i] * rz;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) rhs[A.idx[k]] += A.val[k] * w;
    }
    chol_solve(K, n, rhs.data());
    xt = rhs;
    mulA(xt.data(), zt.data());  // reduced form: z_tilde = A x_tilde
    // update_x, update_z, update_y
    for (int j = 0; j < n; ++j) x[j] = st.alpha * xt[j] + (1 - st.alpha) * xp[j];
    for (int i = 0; i < m; ++i) {
      double v = st.alpha * zt[i] + (1 - st.alpha) * zp[i];
      double zz = v + y[i] / rho_vec[i];
      zz = std::min(std::max(zz, l[i]), u[i]);
      z[i] = zz;
      y[i] += rho_vec[i] * (v - zz);
    }
    bool can_check = st.check_termination && (iter % st.check_termination == 0);
    if (can_check) {
      double ep, ed;
      residuals(false, &ep, &ed);
      if (pri_res <= ep && dua_res <= ed) { solved = true; break; }
    }
    if (st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0)) {
      auto nn = residuals(true, nullptr, nullptr);  // compute_rho_estimate works on scaled quantities
      double pr = pri_res / (nn.first + 1e-10), dr = dua_res / (nn.second + 1e-10);
      double rho_new = rho * std::sqrt(pr / (dr + 1e-10));
      rho_new = std::min(std::max(rho_new, 1e-6), 1e6);
      if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
        rho = rho_new;
        set_rho_vec();
        factor();
        info->rho_updates++;
      }
    }
  }
  if (!solved) {
    if (iter > st.max_iter) iter = st.max_iter;
    double ep, ed;
    residuals(false, &ep, &ed);
    solved = pri_res <= ep && dua_res <= ed;
  }
  for (int j = 0; j < n; ++j) x_out[j] = D[j] * x[j];  // unscale_solution
  info->iter = iter;
  info->status = solved ? 1 : 2;
  info->pri_res = pri_res;
  info->dua_res = dua_res;
}

// ------------------------------------------------------------------------------------------
// Exact solver (long double): Mehrotra IPM on the swing-eliminated QP followed by an active-set
// finish, then a KKT certificate on the LITERAL problem (P, q, A, l, u).
// ------------------------------------------------------------------------------------------
static bool ld_chol(std::vector<ld>& K, int n) {
  for (int j = 0; j < n; ++j) {
    ld d = K[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = sqrtl(d);
    K[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      ld s = K[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= K[(size_t)i * n + k] * K[(size_t)j * n + k];
      K[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
static void ld_chol_solve(const std::vector<ld>& L, int n, ld* b) {
  for (int i = 0; i < n; ++i) {
    ld s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    ld s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

struct ExactInfo { int ipm_iters = 0, rounds = 0, verified = 0; double kkt_stat = 0, kkt_prim = 0, kkt_dual = 0; };

// reduced problem: n = 3K variables, K foot-steps; constraints per foot-step, C u <= d:
//   -fx - mu fz <= 0, fx - mu fz <= 0, -fy - mu fz <= 0, fy - mu fz <= 0, fz <= fzmax, -fz <= -fzmin
static bool exact_reduced(int n, const std::vector<ld>& H, const std::vector<ld>& g, ld mu_f, ld fzmin, ld fzmax,
                          std::vector<ld>& u_out, ExactInfo* info) {
  const int K = n / 3, m = 6 * K;
  // scaling as in any sane IPM: forces in units of fs, cost by max|H|
  const ld fs = 100.0L;
  ld cs = 0;
  for (auto v : H) cs = std::max(cs, fabsl(v));
  cs *= fs * fs;
  std::vector<ld> Hs((size_t)n * n), gs(n);
  for (size_t i = 0; i < Hs.size(); ++i) Hs[i] = H[i] * fs * fs / cs;
  for (int i = 0; i < n; ++i) gs[i] = g[i] * fs / cs;
  const ld dmax = fzmax / fs, dmin = fzmin / fs;
  auto Crow = [&](int r, ld* c3) {  // row r%6 of a foot-step block
    switch (r) {
      case 0: c3[0] = -1; c3[1] = 0; c3[2] = -mu_f; break;
      case 1: c3[0] = 1; c3[1] = 0; c3[2] = -mu_f; break;
      case 2: c3[0] = 0; c3[1] = -1; c3[2] = -mu_f; break;
      case 3: c3[0] = 0; c3[1] = 1; c3[2] = -mu_f; break;
      case 4: c3[0] = 0; c3[1] = 0; c3[2] = 1; break;
      default: c3[0] = 0; c3[1] = 0; c3[2] = -1; break;
    }
  };
  auto dval = [&](int r) -> ld { return r == 4 ? dmax : (r == 5 ? -dmin : 0.0L); };
  std::vector<ld> x(n, 0.0L), s(m), lam(m), rd(n), rp(m), w(m), Kmat((size_t)n * n), rhs(n), dxa(n), dsa(m), dla(m),
      dx(n), ds(m), dl(m), rc(m);
  for (int k = 0; k < K; ++k) x[3 * k + 2] = 0.5L * (dmin + dmax) * 0.5L + 0.5L * dmin;
  ld gmax = 0;
  for (auto v : gs) gmax = std::max(gmax, fabsl(v));
  for (int k = 0; k < K; ++k)
    for (int r = 0; r < 6; ++r) {
      ld c3[3];
      Crow(r, c3);
      ld cu = c3[0] * x[3 * k] + c3[1] * x[3 * k + 1] + c3[2] * x[3 * k + 2];
      s[6 * k + r] = std::max(dval(r) - cu, (ld)1e-2L);
      lam[6 * k + r] = gmax + 1e-3L;
    }
  auto amax = [&](const std::vector<ld>& v, const std::vector<ld>& dv) {
    ld a = 1;
    for (int i = 0; i < m; ++i)
      if (dv[i] < 0) a = std::min(a, -v[i] / dv[i]);
    return a;
  };
  // face states for the finisher
  std::vector<int> zx(K), zy(K), zz(K);
  auto finisher = [&](int maxround) -> bool {
    const ld tol = 1e-15L;
    std::vector<ld> Z((size_t)n * n), c(n), M((size_t)n * n), y(n), u(n), r(n), t1(n);
    std::vector<char> fixed(n);
    for (int rnd = 0; rnd < maxround; ++rnd) {
      info->rounds++;
      std::fill(Z.begin(), Z.end(), 0.0L);
      std::fill(c.begin(), c.end(), 0.0L);
      std::fill(fixed.begin(), fixed.end(), 0);
      for (int k = 0; k < K; ++k) {
        int ix = 3 * k, iy = ix + 1, iz = ix + 2;
        if (zz[k] == -1) {  // pinned at fz = fzmin ... only a vertex when fzmin == 0
          fixed[iz] = 1; c[iz] = dmin;
          if (dmin == 0) { fixed[ix] = fixed[iy] = 1; continue; }
          if (zx[k]) { fixed[ix] = 1; c[ix] = zx[k] * mu_f * dmin; } else Z[(size_t)ix * n + ix] = 1;
          if (zy[k]) { fixed[iy] = 1; c[iy] = zy[k] * mu_f * dmin; } else Z[(size_t)iy * n + iy] = 1;
          continue;
        }
        if (zz[k] == 0) {
          Z[(size_t)iz * n + iz] = 1;
          if (zx[k]) { Z[(size_t)ix * n + iz] = zx[k] * mu_f; fixed[ix] = 1; } else Z[(size_t)ix * n + ix] = 1;
          if (zy[k]) { Z[(size_t)iy * n + iz] = zy[k] * mu_f; fixed[iy] = 1; } else Z[(size_t)iy * n + iy] = 1;
        } else {
          fixed[iz] = 1; c[iz] = dmax;
          if (zx[k]) { c[ix] = zx[k] * mu_f * dmax; fixed[ix] = 1; } else Z[(size_t)ix * n + ix] = 1;
          if (zy[k]) { c[iy] = zy[k] * mu_f * dmax; fixed[iy] = 1; } else Z[(size_t)iy * n + iy] = 1;
        }
      }
      // M = Z' Hs Z + diag(fixed), rhs = -Z'(gs + Hs c)
      std::vector<ld> HZ((size_t)n * n, 0.0L);
      for (int i = 0; i < n; ++i)
        for (int kk = 0; kk < n; ++kk) {
          ld h = Hs[(size_t)i * n + kk];
          if (h == 0) continue;
          int k0 = 3 * (kk / 3);
          for (int j = k0; j < k0 + 3; ++j)
            if (Z[(size_t)kk * n + j] != 0) HZ[(size_t)i * n + j] += h * Z[(size_t)kk * n + j];
        }
      std::fill(M.begin(), M.end(), 0.0L);
      for (int kk = 0; kk < n; ++kk) {
        int k0 = 3 * (kk / 3);
        for (int i = k0; i < k0 + 3; ++i) {
          ld zki = Z[(size_t)kk * n + i];
          if (zki == 0) continue;
          for (int j = 0; j < n; ++j) M[(size_t)i * n + j] += zki * HZ[(size_t)kk * n + j];
        }
      }
      for (int i = 0; i < n; ++i)
        if (fixed[i]) M[(size_t)i * n + i] += 1;
      for (int i = 0; i < n; ++i) {
        ld sacc = gs[i];
        for (int j = 0; j < n; ++j) sacc += Hs[(size_t)i * n + j] * c[j];
        t1[i] = sacc;
      }
      for (int j = 0; j < n; ++j) {
        ld sacc = 0;
        int k0 = 3 * (j / 3);
        for (int i = k0; i < k0 + 3; ++i) sacc += Z[(size_t)i * n + j] * t1[i];
        y[j] = -sacc;
      }
      if (!ld_chol(M, n)) return false;
      ld_chol_solve(M, n, y.data());
      for (int i = 0; i < n; ++i) {
        int k0 = 3 * (i / 3);
        ld sacc = c[i];
        for (int j = k0; j < k0 + 3; ++j) sacc += Z[(size_t)i * n + j] * y[j];
        u[i] = sacc;
      }
      for (int i = 0; i < n; ++i) {
        ld sacc = gs[i];
        for (int j = 0; j < n; ++j) sacc += Hs[(size_t)i * n + j] * u[j];
        r[i] = -sacc;
      }
      bool pv = false;
      for (int k = 0; k < K; ++k) {
        ld fx = u[3 * k], fy = u[3 * k + 1], fz = u[3 * k + 2];
        if (zz[k] == 0 && (fz > dmax + tol || fz < dmin - tol)) pv = true;
        bool xfree = (zx[k] == 0) && !(zz[k] == -1 && dmin == 0), yfree = (zy[k] == 0) && !(zz[k] == -1 && dmin == 0);
        if ((xfree && fabsl(fx) > mu_f * fz + tol) || (yfree && fabsl(fy) > mu_f * fz + tol)) pv = true;
      }
      int changed = 0;
      for (int k = 0; k < K; ++k) {
        ld fx = u[3 * k], fy = u[3 * k + 1], fz = u[3 * k + 2];
        ld rx = r[3 * k], ry = r[3 * k + 1], rz = r[3 * k + 2];
        if (zz[k] == -1 && dmin == 0) {
          if (!pv && -rz / mu_f < fabsl(rx) + fabsl(ry) - tol) {
            zz[k] = 0;
            zx[k] = fabsl(rx) > tol ? (rx > 0 ? 1 : -1) : 0;
            zy[k] = fabsl(ry) > tol ? (ry > 0 ? 1 : -1) : 0;
            changed++;
          }
          continue;
        }
        ld lx = zx[k] ? zx[k] * rx : 0.0L, ly = zy[k] ? zy[k] * ry : 0.0L;
        ld l5 = rz + mu_f * (lx + ly);  // multiplier of fz<=max (if >0) or of fz>=min (if <0)
        int nzx = zx[k], nzy = zy[k], nzz = zz[k];
        if (!pv) {
          if (zx[k] && lx < -tol) nzx = 0;
          if (zy[k] && ly < -tol) nzy = 0;
          if (zz[k] == 1 && l5 < -tol) nzz = 0;
          if (zz[k] == -1 && l5 > tol) nzz = 0;
        }
        if (zz[k] == 0) {
          if (fz > dmax + tol) nzz = 1;
          else if (fz < dmin - tol) nzz = -1;
        }
        if (!(nzz == -1 && dmin == 0)) {
          if (zx[k] == 0 && fabsl(fx) > mu_f * fz + tol) nzx = fx > 0 ? 1 : -1;
          if (zy[k] == 0 && fabsl(fy) > mu_f * fz + tol) nzy = fy > 0 ? 1 : -1;
        }
        if (nzx != zx[k] || nzy != zy[k] || nzz != zz[k]) {
          changed++;
          zx[k] = nzx; zy[k] = nzy; zz[k] = nzz;
          if (nzz == -1 && dmin == 0) zx[k] = zy[k] = 0;
        }
      }
      if (changed == 0) {
        for (int i = 0; i < n; ++i) u_out[i] = u[i] * fs;
        return true;
      }
    }
    return false;
  };

  ld mu_target = 1e-10L;
  for (int attempt = 0; attempt < 4; ++attempt) {
    bool ipm_ok = false;
    for (; info->ipm_iters < 80;) {
      // residuals
      ld mu = 0, rdmax = 0, rpmax = 0;
      for (int i = 0; i < n; ++i) {
        ld sacc = gs[i];
        for (int j = 0; j < n; ++j) sacc += Hs[(size_t)i * n + j] * x[j];
        rd[i] = sacc;
      }
      for (int k = 0; k < K; ++k)
        for (int r = 0; r < 6; ++r) {
          ld c3[3];
          Crow(r, c3);
          int i = 6 * k + r;
          for (int a = 0; a < 3; ++a) rd[3 * k + a] += c3[a] * lam[i];
          rp[i] = c3[0] * x[3 * k] + c3[1] * x[3 * k + 1] + c3[2] * x[3 * k + 2] + s[i] - dval(r);
          mu += s[i] * lam[i];
          rpmax = std::max(rpmax, fabsl(rp[i]));
        }
      mu /= m;
      for (int i = 0; i < n; ++i) rdmax = std::max(rdmax, fabsl(rd[i]));
      if (mu < mu_target && rdmax < 1e-9L && rpmax < 1e-9L) { ipm_ok = true; break; }
      for (int i = 0; i < m; ++i) w[i] = lam[i] / s[i];
      Kmat = Hs;
      for (int k = 0; k < K; ++k)
        for (int r = 0; r < 6; ++r) {
          ld c3[3];
          Crow(r, c3);
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Kmat[(size_t)(3 * k + a) * n + 3 * k + b] += w[6 * k + r] * c3[a] * c3[b];
        }
      if (!ld_chol(Kmat, n)) break;
      auto solve = [&](const std::vector<ld>& rc_, std::vector<ld>& dx_, std::vector<ld>& ds_, std::vector<ld>& dl_) {
        for (int i = 0; i < n; ++i) rhs[i] = -rd[i];
        for (int k = 0; k < K; ++k)
          for (int r = 0; r < 6; ++r) {
            ld c3[3];
            Crow(r, c3);
            int i = 6 * k + r;
            ld t = rc_[i] / s[i] - w[i] * rp[i];
            for (int a = 0; a < 3; ++a) rhs[3 * k + a] += c3[a] * t;
          }
        ld_chol_solve(Kmat, n, rhs.data());
        dx_ = rhs;
        for (int k = 0; k < K; ++k)
          for (int r = 0; r < 6; ++r) {
            ld c3[3];
            Crow(r, c3);
            int i = 6 * k + r;
            ds_[i] = -rp[i] - (c3[0] * dx_[3 * k] + c3[1] * dx_[3 * k + 1] + c3[2] * dx_[3 * k + 2]);
            dl_[i] = -(rc_[i] + lam[i] * ds_[i]) / s[i];
          }
      };
      for (int i = 0; i < m; ++i) rc[i] = s[i] * lam[i];
      solve(rc, dxa, dsa, dla);
      ld aa = std::min(amax(s, dsa), amax(lam, dla));
      ld mu_aff = 0;
      for (int i = 0; i < m; ++i) mu_aff += (s[i] + aa * dsa[i]) * (lam[i] + aa * dla[i]);
      mu_aff /= m;
      ld sigma = powl(mu_aff / mu, 3);
      for (int i = 0; i < m; ++i) rc[i] = s[i] * lam[i] + dsa[i] * dla[i] - sigma * mu;
      solve(rc, dx, ds, dl);
      ld ap = amax(s, ds), ad = amax(lam, dl);
      ld a = std::min(ap < 1 ? 0.995L * ap : 1.0L, ad < 1 ? 0.995L * ad : 1.0L);
      for (int i = 0; i < n; ++i) x[i] += a * dx[i];
      for (int i = 0; i < m; ++i) { s[i] += a * ds[i]; lam[i] += a * dl[i]; }
      info->ipm_iters++;
    }
    // guess faces from the iterate
    for (int k = 0; k < K; ++k) {
      bool act[6];
      for (int r = 0; r < 6; ++r) act[r] = lam[6 * k + r] > s[6 * k + r];
      zx[k] = zy[k] = zz[k] = 0;
      if (dmin == 0 && ((act[0] && act[1]) || (act[2] && act[3]) || act[5])) { zz[k] = -1; continue; }
      zx[k] = act[0] ? -1 : (act[1] ? 1 : 0);
      zy[k] = act[2] ? -1 : (act[3] ? 1 : 0);
      zz[k] = act[4] ? 1 : (act[5] ? -1 : 0);
    }
    if (finisher(6)) { info->verified = 1; return true; }
    if (!ipm_ok) break;
    mu_target *= 1e-3L;
  }
  for (int i = 0; i < n; ++i) u_out[i] = x[i] * fs;
  info->verified = 0;
  return false;
}

// KKT certificate on the literal problem: given x, find multipliers y per foot-step in closed form
// and report stationarity / feasibility residuals (all in the problem's own units).
static void kkt_certificate(int N, const Mat& P, const double* q, const double* lbv, const double* ubv, double mu_f,
                            const std::vector<ld>& x, ExactInfo* info) {
  const int n = 12 * N;
  ld stat = 0, prim = 0, dual = 0;
  std::vector<ld> r(n);
  for (int i = 0; i < n; ++i) {
    ld s = q[i];
    for (int j = 0; j < n; ++j) s += (ld)P(i, j) * x[j];
    r[i] = -s;  // must equal A' y
  }
  for (int k = 0; k < 4 * N; ++k) {
    ld fx = x[3 * k], fy = x[3 * k + 1], fz = x[3 * k + 2];
    ld rx = r[3 * k], ry = r[3 * k + 1], rz = r[3 * k + 2];
    ld lo5 = lbv[5 * k + 4], up5 = ubv[5 * k + 4];
    // primal feasibility of the 5 rows
    prim = std::max(prim, -(fx + mu_f * fz));
    prim = std::max(prim, (fx - mu_f * fz));
    prim = std::max(prim, -(fy + mu_f * fz));
    prim = std::max(prim, (fy - mu_f * fz));
    prim = std::max(prim, lo5 - fz);
    prim = std::max(prim, fz - up5);
    const ld tolact = 1e-9L;
    // rows: y0 <= 0 on fx+mu fz >= 0 ; y1 >= 0 on fx-mu fz <= 0 ; y2, y3 likewise ; y4 on lo<=fz<=up
    // stationarity: rx = y0 + y1 ; ry = y2 + y3 ; rz = mu(y0 - y1 + y2 - y3) + y4
    bool a0 = fabsl(fx + mu_f * fz) <= tolact, a1 = fabsl(fx - mu_f * fz) <= tolact;
    bool a2 = fabsl(fy + mu_f * fz) <= tolact, a3 = fabsl(fy - mu_f * fz) <= tolact;
    bool alo = fabsl(fz - lo5) <= tolact, aup = fabsl(fz - up5) <= tolact;
    ld y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    // x pair
    if (a0 && a1) { y0 = std::min(rx, (ld)0); y1 = std::max(rx, (ld)0); }
    else if (a0) y0 = rx; else if (a1) y1 = rx;
    if (a2 && a3) { y2 = std::min(ry, (ld)0); y3 = std::max(ry, (ld)0); }
    else if (a2) y2 = ry; else if (a3) y3 = ry;
    stat = std::max(stat, fabsl(rx - y0 - y1));
    stat = std::max(stat, fabsl(ry - y2 - y3));
    dual = std::max(dual, y0);    // y0 must be <= 0
    dual = std::max(dual, -y1);   // y1 >= 0
    dual = std::max(dual, y2);
    dual = std::max(dual, -y3);
    ld y4 = rz - mu_f * (y0 - y1 + y2 - y3);
    // at the vertex fx=fy=fz=0 the pairs (y0,y1), (y2,y3) are determined only up to a common shift t>=0 that
    // RAISES y4 by 2 mu t; the smallest y4 (t=0, computed above) is the one to test at a lower bound.
    if (alo && aup) { /* equality row: free sign */ }
    else if (alo) dual = std::max(dual, y4);       // need y4 <= 0
    else if (aup) dual = std::max(dual, -y4);      // need y4 >= 0
    else stat = std::max(stat, fabsl(y4));         // inactive: must vanish
  }
  info->kkt_stat = (double)stat;
  info->kkt_prim = (double)std::max(prim, (ld)0);
  info->kkt_dual = (double)std::max(dual, (ld)0);
}

// exact solve of the literal MPC QP (P 12N x 12N, q, pyramid constraints); step_masks[k] = contact mask of horizon step k
static bool exact_solve_literal_sched(int N, const Mat& P, const double* q, const double* lbv, const double* ubv,
                                      const uint32_t* step_masks, double mu_f, double fzmin, double fzmax, double* x_out,
                                      ExactInfo* info) {
  const int nfull = 12 * N;
  std::vector<int> idx;
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < 4; ++i)
      if ((step_masks[k] >> i) & 1u)
        for (int a = 0; a < 3; ++a) idx.push_back(12 * k + 3 * i + a);
  const int n = (int)idx.size();
  std::vector<ld> xfull(nfull, 0.0L);
  bool ok = true;
  if (n > 0) {
    std::vector<ld> H((size_t)n * n), g(n), u(n);
    for (int i = 0; i < n; ++i) {
      g[i] = q[idx[i]];
      for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = P(idx[i], idx[j]);
    }
    ok = exact_reduced(n, H, g, mu_f, fzmin, fzmax, u, info);
    for (int i = 0; i < n; ++i) xfull[idx[i]] = u[i];
  } else {
    info->verified = 1;
  }
  kkt_certificate(N, P, q, lbv, ubv, mu_f, xfull, info);
  for (int i = 0; i < nfull; ++i) x_out[i] = (double)xfull[i];
  return ok;
}
static bool exact_solve_literal(int N, const Mat& P, const double* q, const double* lbv, const double* ubv,
                                const bool* contacts, double mu_f, double fzmin, double fzmax, double* x_out,
                                ExactInfo* info) {
  uint32_t m = 0;
  for (int i = 0; i < 4; ++i) m |= contacts[i] ? (1u << i) : 0u;
  std::vector<uint32_t> masks(N, m);
  return exact_solve_literal_sched(N, P, q, lbv, ubv, masks.data(), mu_f, fzmin, fzmax, x_out, info);
}

// rotation taking world z to the unit normal n (about the horizontal axis z x n); row-major 3x3
static void terrain_frame(const double* n_in, double* R) {
  double nx = n_in[0], ny = n_in[1], nz = n_in[2];
  const double inv = 1.0 / std::sqrt(nx * nx + ny * ny + nz * nz);
  nx *= inv; ny *= inv; nz *= inv;
  const double k = 1.0 / (1.0 + nz);
  R[0] = 1 - nx * nx * k; R[1] = -nx * ny * k;    R[2] = nx;
  R[3] = -nx * ny * k;    R[4] = 1 - ny * ny * k; R[5] = ny;
  R[6] = -nx;             R[7] = -ny;             R[8] = nz;
}

}  // namespace

// ==========================================================================================
// C entry points (loaded with ctypes by tests/ and bench.py only)
// ==========================================================================================
extern "C" {

enum { ORACLE_MODE_EXACT = 0, ORACLE_MODE_OSQP_DEFAULT = 1, ORACLE_MODE_OSQP_TIGHT = 2, ORACLE_MODE_BUILD_ONLY = 3 };

// ConvexMpc members after compute_grf drove it, for QP `b` of the batch.  Any output may be NULL.
int oracle_build_qp(const a1mpc_config* cfg, const a1mpc_inputs* in, int b, double* H, double* g, double* A,
                    double* lb, double* ub) {
  const int N = cfg->horizon;
  RobotState st;
  unpack(in, b, &st);
  ConvexMpcRestated mpc(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  std::vector<double> x0, xd;
  drive_convex_mpc(mpc, *cfg, st, x0, xd);
  const int n = 12 * N, m = 20 * N;
  if (H) std::memcpy(H, mpc.hessian.a.data(), sizeof(double) * n * n);
  if (g) std::memcpy(g, mpc.gradient.data(), sizeof(double) * n);
  if (A) std::memcpy(A, mpc.linear_constraints.a.data(), sizeof(double) * m * n);
  if (lb) std::memcpy(lb, mpc.lb.data(), sizeof(double) * m);
  if (ub) std::memcpy(ub, mpc.ub.data(), sizeof(double) * m);
  return 0;
}

// the intermediate members for QP `b`: mpc_states [13], mpc_states_d [13N] (A1RobotControl.cpp:452-488), A_mat_d [13x13],
// B_mat_d_list [13N x 12], A_qp [13N x 13], B_qp [13N x 12N] (ConvexMpc.cpp:145-202), all row-major.  Any output may be NULL.
int oracle_rollout(const a1mpc_config* cfg, const a1mpc_inputs* in, int b, double* mpc_states, double* mpc_states_d, double* A_d,
                   double* B_d_list, double* A_qp, double* B_qp) {
  const int N = cfg->horizon;
  RobotState st;
  unpack(in, b, &st);
  ConvexMpcRestated mpc(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  std::vector<double> x0, xd;
  drive_convex_mpc(mpc, *cfg, st, x0, xd);
  if (mpc_states) std::memcpy(mpc_states, x0.data(), sizeof(double) * 13);
  if (mpc_states_d) std::memcpy(mpc_states_d, xd.data(), sizeof(double) * 13 * N);
  if (A_d) std::memcpy(A_d, mpc.A_mat_d.a.data(), sizeof(double) * 13 * 13);
  if (B_d_list) std::memcpy(B_d_list, mpc.B_mat_d_list.a.data(), sizeof(double) * 13 * N * 12);
  if (A_qp) std::memcpy(A_qp, mpc.A_qp.a.data(), sizeof(double) * 13 * N * 13);
  if (B_qp) std::memcpy(B_qp, mpc.B_qp.a.data(), sizeof(double) * 13 * N * 12 * N);
  return 0;
}

// general ConvexMpc::calculate_qp_mats with caller-supplied A_d, B_d_list (test_mpc.cpp:106-125)
int oracle_qp_mats(const a1mpc_config* cfg, const double* A_d, const double* B_d_list, const double* x0,
                   const double* x_d, double* H, double* g) {
  const int N = cfg->horizon;
  ConvexMpcRestated mpc(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  for (int i = 0; i < 13; ++i)
    for (int j = 0; j < 13; ++j) mpc.A_mat_d(i, j) = A_d[13 * i + j];
  for (int i = 0; i < 13 * N; ++i)
    for (int j = 0; j < 12; ++j) mpc.B_mat_d_list(i, j) = B_d_list[12 * i + j];
  bool contacts[4] = {true, true, true, true};
  mpc.calculate_qp_mats(x0, x_d, contacts);
  const int n = 12 * N;
  if (H) std::memcpy(H, mpc.hessian.a.data(), sizeof(double) * n * n);
  if (g) std::memcpy(g, mpc.gradient.data(), sizeof(double) * n);
  return 0;
}

// info[8]: iters, status/verified, kkt_stat, kkt_prim, kkt_dual, rounds, pri_res, dua_res
// The reference constructs a ConvexMpc per control tick (A1RobotControl.cpp:447) out of fixed-size Eigen members, i.e.
// without heap traffic; the restatement keeps one object per worker thread so that the multi-threaded baseline is not
// throttled by mmap/munmap of its 100 KB matrices (the constructor's work -- tiling weights, filling the pyramid -- is
// repeated per QP through reinit()).
static int solve_one(const a1mpc_config* cfg, const a1mpc_inputs* in, int b, int mode, double eps, double* f_body,
                     double* u_full, double* info8, ConvexMpcRestated* ws = nullptr) {
  const int N = cfg->horizon;
  const int n = 12 * N, m = 20 * N;
  RobotState st;
  unpack(in, b, &st);
  ConvexMpcRestated local_or_ws = ws ? ConvexMpcRestated(1, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max) : ConvexMpcRestated(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  ConvexMpcRestated& mpc = ws ? *ws : local_or_ws;
  if (ws) ws->reinit(cfg->q, cfg->r);
  std::vector<double> x0, xd, sol(n, 0.0);
  drive_convex_mpc(mpc, *cfg, st, x0, xd);
  if (info8) std::fill(info8, info8 + 8, 0.0);
  if (mode == ORACLE_MODE_EXACT) {
    ExactInfo ei;
    exact_solve_literal(N, mpc.hessian, mpc.gradient.data(), mpc.lb.data(), mpc.ub.data(), st.contacts, cfg->mu,
                        cfg->fz_min, cfg->fz_max, sol.data(), &ei);
    if (info8) {
      info8[0] = ei.ipm_iters; info8[1] = ei.verified; info8[2] = ei.kkt_stat; info8[3] = ei.kkt_prim;
      info8[4] = ei.kkt_dual; info8[5] = ei.rounds;
    }
  } else if (mode == ORACLE_MODE_OSQP_DEFAULT || mode == ORACLE_MODE_OSQP_TIGHT) {
    OsqpSettings os;
    if (mode == ORACLE_MODE_OSQP_TIGHT) { os.eps_abs = os.eps_rel = (eps > 0 ? eps : 1e-11); os.max_iter = 400000; }
    SparseRows A;
    A.from_dense(mpc.linear_constraints);
    OsqpInfo oi;
    osqp_restated_solve(n, m, mpc.hessian, mpc.gradient.data(), A, mpc.lb.data(), mpc.ub.data(), os, sol.data(), &oi);
    if (info8) { info8[0] = oi.iter; info8[1] = oi.status; info8[5] = oi.rho_updates; info8[6] = oi.pri_res; info8[7] = oi.dua_res; }
  }
  // A1RobotControl.cpp:555-561: f_body = R^T * solution[3i:3i+3]
  if (f_body)
    for (int i = 0; i < 4; ++i)
      for (int a = 0; a < 3; ++a)
        f_body[3 * i + a] = st.R[0 * 3 + a] * sol[3 * i] + st.R[1 * 3 + a] * sol[3 * i + 1] + st.R[2 * 3 + a] * sol[3 * i + 2];
  if (u_full) std::memcpy(u_full, sol.data(), sizeof(double) * n);
  return 0;
}

// Batched compute_grf (MPC branch) on `nthreads` host threads, one QP per task.
// f_body [12][B] SoA (ld = B), u_full [B][12N] QP-major or NULL, info [B][8] or NULL.
int oracle_compute_grf_batch(const a1mpc_config* cfg, int B, const a1mpc_inputs* in, int mode, double eps,
                             int nthreads, double* f_body, double* u_full, double* info) {
  if (nthreads < 1) nthreads = 1;
  const int n = 12 * cfg->horizon;
  std::atomic<int> next(0);
  auto work = [&]() {
    ConvexMpcRestated ws(cfg->horizon, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      double f[12];
      solve_one(cfg, in, b, mode, eps, f, u_full ? u_full + (size_t)b * n : nullptr, info ? info + (size_t)b * 8 : nullptr, &ws);
      if (f_body)
        for (int k = 0; k < 12; ++k) f_body[(size_t)k * B + b] = f[k];
    }
  };
  if (nthreads == 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
  return 0;
}

// BASELINE config 4 (extension beyond the reference): per-step contact masks sched[N] and per-foot terrain normals (12) for
// QP b.  The literal problem keeps all 12N world-frame forces; with normals the pyramid rows act on Rf^T f, which is the
// same as solving for local forces with P' = Rb^T P Rb, q' = Rb^T q and the unrotated pyramid.
static int solve_one_ext(const a1mpc_config* cfg, const a1mpc_inputs* in, int b, const uint32_t* sched, size_t sched_ld,
                         const double* normals, size_t normals_ld, int mode, double* f_body, double* u_full, double* info8) {
  const int N = cfg->horizon;
  const int n = 12 * N, m = 20 * N;
  RobotState st;
  unpack(in, b, &st);
  ConvexMpcRestated mpc(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  std::vector<double> x0, xd, sol(n, 0.0);
  drive_convex_mpc(mpc, *cfg, st, x0, xd);
  std::vector<uint32_t> masks(N);
  for (int k = 0; k < N; ++k) masks[k] = (sched ? sched[(size_t)k * sched_ld + b] : in->contact[b]) & 15u;
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < 4; ++i) {
      const double c = ((masks[k] >> i) & 1u) ? 1.0 : 0.0;
      mpc.lb[20 * k + 5 * i + 4] = cfg->fz_min * c;
      mpc.ub[20 * k + 5 * i + 4] = cfg->fz_max * c;
    }
  double Rf[4][9];
  for (int i = 0; i < 4; ++i) {
    double nn[3] = {0, 0, 1};
    if (normals)
      for (int a = 0; a < 3; ++a) nn[a] = normals[(size_t)(3 * i + a) * normals_ld + b];
    terrain_frame(nn, Rf[i]);
  }
  Mat P = mpc.hessian;
  std::vector<double> q = mpc.gradient;
  if (normals) {  // P' = Rb^T P Rb, q' = Rb^T q  (Rb = blockdiag of the foot frames, the same frame at every step)
    Mat T(n, n);
    for (int r = 0; r < n; ++r)
      for (int kf = 0; kf < 4 * N; ++kf) {
        const double* R = Rf[kf % 4];
        for (int c2 = 0; c2 < 3; ++c2) {
          double sacc = 0;
          for (int a = 0; a < 3; ++a) sacc += mpc.hessian(r, 3 * kf + a) * R[3 * a + c2];
          T(r, 3 * kf + c2) = sacc;
        }
      }
    for (int kf = 0; kf < 4 * N; ++kf) {
      const double* R = Rf[kf % 4];
      for (int c2 = 0; c2 < 3; ++c2) {
        for (int col = 0; col < n; ++col) {
          double sacc = 0;
          for (int a = 0; a < 3; ++a) sacc += R[3 * a + c2] * T(3 * kf + a, col);
          P(3 * kf + c2, col) = sacc;
        }
        double sq = 0;
        for (int a = 0; a < 3; ++a) sq += R[3 * a + c2] * mpc.gradient[3 * kf + a];
        q[3 * kf + c2] = sq;
      }
    }
  }
  if (info8) std::fill(info8, info8 + 8, 0.0);
  if (mode == ORACLE_MODE_EXACT) {
    ExactInfo ei;
    exact_solve_literal_sched(N, P, q.data(), mpc.lb.data(), mpc.ub.data(), masks.data(), cfg->mu, cfg->fz_min, cfg->fz_max, sol.data(), &ei);
    if (info8) { info8[0] = ei.ipm_iters; info8[1] = ei.verified; info8[2] = ei.kkt_stat; info8[3] = ei.kkt_prim; info8[4] = ei.kkt_dual; info8[5] = ei.rounds; }
  } else {
    OsqpSettings os;
    if (mode == ORACLE_MODE_OSQP_TIGHT) { os.eps_abs = os.eps_rel = 1e-11; os.max_iter = 400000; }
    SparseRows A;
    A.from_dense(mpc.linear_constraints);
    OsqpInfo oi;
    osqp_restated_solve(n, m, P, q.data(), A, mpc.lb.data(), mpc.ub.data(), os, sol.data(), &oi);
    if (info8) { info8[0] = oi.iter; info8[1] = oi.status; }
  }
  // local -> world
  std::vector<double> uw(n, 0.0);
  for (int kf = 0; kf < 4 * N; ++kf) {
    const double* R = Rf[kf % 4];
    for (int a = 0; a < 3; ++a) uw[3 * kf + a] = R[3 * a] * sol[3 * kf] + R[3 * a + 1] * sol[3 * kf + 1] + R[3 * a + 2] * sol[3 * kf + 2];
  }
  if (f_body)
    for (int i = 0; i < 4; ++i)
      for (int a = 0; a < 3; ++a)
        f_body[3 * i + a] = st.R[0 * 3 + a] * uw[3 * i] + st.R[1 * 3 + a] * uw[3 * i + 1] + st.R[2 * 3 + a] * uw[3 * i + 2];
  if (u_full) std::memcpy(u_full, uw.data(), sizeof(double) * n);
  return 0;
}

// sched [N][B] (ld = B) or NULL, normals [12][B] (ld = B) or NULL; outputs as oracle_compute_grf_batch
int oracle_compute_grf_batch_ext(const a1mpc_config* cfg, int B, const a1mpc_inputs* in, const uint32_t* sched, const double* normals,
                                 int mode, int nthreads, double* f_body, double* u_full, double* info) {
  if (nthreads < 1) nthreads = 1;
  const int n = 12 * cfg->horizon;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      double f[12];
      solve_one_ext(cfg, in, b, sched, (size_t)B, normals, (size_t)B, mode, f, u_full ? u_full + (size_t)b * n : nullptr, info ? info + (size_t)b * 8 : nullptr);
      if (f_body)
        for (int k = 0; k < 12; ++k) f_body[(size_t)k * B + b] = f[k];
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
  return 0;
}

// Solve a caller-supplied literal MPC QP (H 12N x 12N row-major, g) with contact pattern.
int oracle_solve_dense(const a1mpc_config* cfg, const double* H, const double* g, uint32_t contact, int mode,
                       double* u, double* info8) {
  const int N = cfg->horizon, n = 12 * N, m = 20 * N;
  ConvexMpcRestated mpc(N, cfg->q, cfg->r, cfg->mu, cfg->fz_min, cfg->fz_max);
  std::memcpy(mpc.hessian.a.data(), H, sizeof(double) * n * n);
  bool contacts[4];
  for (int i = 0; i < 4; ++i) contacts[i] = (contact >> i) & 1u;
  // bounds as ConvexMpc.cpp:223-245
  for (int k = 0; k < 4 * N; ++k) {
    double c = contacts[k % 4] ? 1.0 : 0.0;
    mpc.lb[5 * k + 0] = 0; mpc.ub[5 * k + 0] = OSQP_INFTY;
    mpc.lb[5 * k + 1] = -OSQP_INFTY; mpc.ub[5 * k + 1] = 0;
    mpc.lb[5 * k + 2] = 0; mpc.ub[5 * k + 2] = OSQP_INFTY;
    mpc.lb[5 * k + 3] = -OSQP_INFTY; mpc.ub[5 * k + 3] = 0;
    mpc.lb[5 * k + 4] = cfg->fz_min * c; mpc.ub[5 * k + 4] = cfg->fz_max * c;
  }
  if (info8) std::fill(info8, info8 + 8, 0.0);
  if (mode == ORACLE_MODE_EXACT) {
    ExactInfo ei;
    exact_solve_literal(N, mpc.hessian, g, mpc.lb.data(), mpc.ub.data(), contacts, cfg->mu, cfg->fz_min, cfg->fz_max, u, &ei);
    if (info8) { info8[0] = ei.ipm_iters; info8[1] = ei.verified; info8[2] = ei.kkt_stat; info8[3] = ei.kkt_prim; info8[4] = ei.kkt_dual; info8[5] = ei.rounds; }
  } else {
    OsqpSettings os;
    if (mode == ORACLE_MODE_OSQP_TIGHT) { os.eps_abs = os.eps_rel = 1e-11; os.max_iter = 400000; }
    SparseRows A;
    A.from_dense(mpc.linear_constraints);
    OsqpInfo oi;
    osqp_restated_solve(n, m, mpc.hessian, g, A, mpc.lb.data(), mpc.ub.data(), os, u, &oi);
    if (info8) { info8[0] = oi.iter; info8[1] = oi.status; }
  }
  return 0;
}

// Generic dense QP  min 1/2 x'Px + q'x  s.t. l <= Ax <= u  through the OSQP-algorithm restatement, with the signature of the
// solve hook of oracle/ref_shim/OsqpEigen/OsqpEigen.h (P n x n, A m x n, both COLUMN-major): this is what stands in for the absent
// OSQP when the reference's own compute_grf / test_mpc.cpp run from oracle/_ref.  Always a cold start (x = y = 0): run to
// eps 1e-11 the starting point does not matter, the optimum is unique (P positive definite).  y is not produced.
static int qp_hook_common(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u,
                          bool tight, double* x) {
  Mat Pm(n, n), Am(m, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Pm(i, j) = P[(size_t)j * n + i];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) Am(i, j) = A[(size_t)j * m + i];
  OsqpSettings os;
  if (tight) { os.eps_abs = os.eps_rel = 1e-11; os.max_iter = 400000; }
  SparseRows As;
  As.from_dense(Am);
  OsqpInfo oi;
  osqp_restated_solve(n, m, Pm, q, As, l, u, os, x, &oi);
  return oi.status == 1 ? 0 : (tight ? 2 : 0);  // default settings may stop at max_iter like OSQP does; still "a solution"
}
int oracle_qp_hook_osqp_tight(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u,
                              int /*warm*/, double* x, double* /*y*/) {
  return qp_hook_common(n, m, P, q, A, l, u, true, x);
}
int oracle_qp_hook_osqp_default(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u,
                                int /*warm*/, double* x, double* /*y*/) {
  return qp_hook_common(n, m, P, q, A, l, u, false, x);
}

// compute_grf's QP branch (A1RobotControl.cpp:11-48, 377-445): 12 variables, 20 rows.
// root_acc[6], rot_z[9], rot[9] row-major, foot[12] leg-major, contact mask.  mode as above.
int oracle_grf_qp_single(const double* root_acc, const double* rot_z, const double* rot, const double* foot,
                         uint32_t contact, int mode, double* f_body, double* info8) {
  const double Qd[6] = {1.0, 1.0, 1.0, 400.0, 400.0, 100.0};
  const double Rw = 1e-3, mu = 0.7, F_min = 0, F_max = 180;
  // inertia_inv 6x12 (:394-399)
  double Minv[6][12];
  double RzT[9];
  mat3_T(rot_z, RzT);
  for (int i = 0; i < 4; ++i) {
    double S[9], M3[9];
    skew(&foot[3 * i], S);
    mat3_mul(RzT, S, M3);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        Minv[a][3 * i + b] = (a == b) ? 1.0 : 0.0;
        Minv[3 + a][3 * i + b] = M3[3 * a + b];
      }
  }
  Mat P(12, 12);
  double q[12];
  for (int i = 0; i < 12; ++i) {
    for (int j = 0; j < 12; ++j) {
      double s = (i == j) ? Rw : 0.0;
      for (int k = 0; k < 6; ++k) s += Minv[k][i] * Qd[k] * Minv[k][j];
      P(i, j) = s;
    }
    double s = 0;
    for (int k = 0; k < 6; ++k) s += Minv[k][i] * Qd[k] * root_acc[k];
    q[i] = -s;
  }
  // linearMatrix (:28-48) 20 x 12, bounds
  Mat A(20, 12);
  double lb[20], ub[20];
  for (int i = 0; i < 4; ++i) {
    double c = ((contact >> i) & 1u) ? 1.0 : 0.0;
    A(i, 2 + 3 * i) = 1; lb[i] = c * F_min; ub[i] = c * F_max;
    A(4 + 4 * i, 3 * i) = 1;      A(4 + 4 * i, 2 + 3 * i) = -mu;
    A(4 + 4 * i + 1, 3 * i) = -1; A(4 + 4 * i + 1, 2 + 3 * i) = -mu;
    A(4 + 4 * i + 2, 1 + 3 * i) = 1;  A(4 + 4 * i + 2, 2 + 3 * i) = -mu;
    A(4 + 4 * i + 3, 1 + 3 * i) = -1; A(4 + 4 * i + 3, 2 + 3 * i) = -mu;
    for (int k = 0; k < 4; ++k) { lb[4 + 4 * i + k] = -OSQP_INFTY; ub[4 + 4 * i + k] = 0; }
  }
  double sol[12] = {0};
  if (info8) std::fill(info8, info8 + 8, 0.0);
  if (mode == ORACLE_MODE_EXACT) {
    // same feasible set as a 1-step pyramid with contact-gated fz bounds: reuse the exact solver
    bool contacts[4];
    for (int i = 0; i < 4; ++i) contacts[i] = (contact >> i) & 1u;
    double lb5[20], ub5[20];
    for (int i = 0; i < 4; ++i) {
      double c = contacts[i] ? 1.0 : 0.0;
      lb5[5 * i + 0] = 0; ub5[5 * i + 0] = OSQP_INFTY; lb5[5 * i + 1] = -OSQP_INFTY; ub5[5 * i + 1] = 0;
      lb5[5 * i + 2] = 0; ub5[5 * i + 2] = OSQP_INFTY; lb5[5 * i + 3] = -OSQP_INFTY; ub5[5 * i + 3] = 0;
      lb5[5 * i + 4] = F_min * c; ub5[5 * i + 4] = F_max * c;
    }
    ExactInfo ei;
    exact_solve_literal(1, P, q, lb5, ub5, contacts, mu, F_min, F_max, sol, &ei);
    if (info8) { info8[0] = ei.ipm_iters; info8[1] = ei.verified; info8[2] = ei.kkt_stat; info8[3] = ei.kkt_prim; info8[4] = ei.kkt_dual; info8[5] = ei.rounds; }
  } else {
    OsqpSettings os;
    if (mode == ORACLE_MODE_OSQP_TIGHT) { os.eps_abs = os.eps_rel = 1e-11; os.max_iter = 400000; }
    SparseRows As;
    As.from_dense(A);
    OsqpInfo oi;
    osqp_restated_solve(12, 20, P, q, As, lb, ub, os, sol, &oi);
    if (info8) { info8[0] = oi.iter; info8[1] = oi.status; }
  }
  for (int i = 0; i < 4; ++i)
    for (int a = 0; a < 3; ++a)
      f_body[3 * i + a] = rot[0 * 3 + a] * sol[3 * i] + rot[1 * 3 + a] * sol[3 * i + 1] + rot[2 * 3 + a] * sol[3 * i + 2];
  return 0;
}

// A1RobotControl::compute_joint_torques (A1RobotControl.cpp:289-319) for one robot.  jac: four 3x3 row-major blocks,
// f_grf / f_kin leg-major, tau in/out (NaN results keep the previous value).  Partial-pivot LU like Eigen's lu().
int oracle_joint_torques(const double* f_grf, const double* f_kin, const double* jac, uint32_t contact, const double* km_foot,
                         const double* torques_gravity, double* tau) {
  for (int leg = 0; leg < 4; ++leg) {
    const double* J = jac + 9 * leg;
    double t[3];
    if ((contact >> leg) & 1u) {
      for (int a = 0; a < 3; ++a) t[a] = J[a] * -f_grf[3 * leg] + J[3 + a] * -f_grf[3 * leg + 1] + J[6 + a] * -f_grf[3 * leg + 2];
    } else {
      long double A[3][4];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) A[i][j] = J[3 * i + j];
        A[i][3] = (long double)km_foot[i] * f_kin[3 * leg + i];
      }
      for (int c = 0; c < 3; ++c) {
        int p = c;
        for (int i = c + 1; i < 3; ++i)
          if (fabsl(A[i][c]) > fabsl(A[p][c])) p = i;
        for (int j = 0; j < 4; ++j) std::swap(A[c][j], A[p][j]);
        for (int i = c + 1; i < 3; ++i) {
          long double m = A[i][c] / A[c][c];
          for (int j = c; j < 4; ++j) A[i][j] -= m * A[c][j];
        }
      }
      long double x[3];
      for (int i = 2; i >= 0; --i) {
        long double s = A[i][3];
        for (int j = i + 1; j < 3; ++j) s -= A[i][j] * x[j];
        x[i] = s / A[i][i];
      }
      for (int a = 0; a < 3; ++a) t[a] = (double)x[a];
    }
    for (int a = 0; a < 3; ++a) {
      double v = t[a] + torques_gravity[3 * leg + a];
      if (!std::isnan(v)) tau[3 * leg + a] = v;
    }
  }
  return 0;
}

// A1RobotControl::update_plan (A1RobotControl.cpp:148-202) for one robot; arrays as in the reference (3 x NUM_LEG row-major).
// sched_out[N]: planned contact mask i ticks ahead (what update_plan would set after i more calls) -- the extension's input.
int oracle_update_plan(double counter_per_gait, double counter_per_swing, double control_dt, const double* default_foot_pos,
                       double dx_lim, double dy_lim, int movement_mode, double* gait_counter, const double* gait_counter_speed,
                       const double* lin_vel, const double* lin_vel_d, const double* rot_z, const double* rot, const double* root_pos,
                       int N, uint32_t* plan_contacts, uint32_t* sched_out, double* t_rel, double* t_abs, double* t_world) {
  bool plan[4];
  if (!movement_mode) {
    for (int i = 0; i < 4; ++i) plan[i] = true;
    gait_counter[0] = 0; gait_counter[1] = 120; gait_counter[2] = 120; gait_counter[3] = 0;
  } else {
    for (int i = 0; i < 4; ++i) {
      gait_counter[i] = gait_counter[i] + gait_counter_speed[i];
      gait_counter[i] = std::fmod(gait_counter[i], counter_per_gait);
      plan[i] = gait_counter[i] <= counter_per_swing;
    }
  }
  uint32_t m = 0;
  for (int i = 0; i < 4; ++i) m |= plan[i] ? (1u << i) : 0u;
  *plan_contacts = m;
  // look-ahead: repeat the counter update i more times on a copy
  double gc[4] = {gait_counter[0], gait_counter[1], gait_counter[2], gait_counter[3]};
  for (int st = 0; st < N; ++st) {
    uint32_t ms = 0;
    for (int i = 0; i < 4; ++i) {
      if (!movement_mode) { ms |= 1u << i; continue; }
      if (st > 0) gc[i] = std::fmod(gc[i] + gait_counter_speed[i], counter_per_gait);
      if (gc[i] <= counter_per_swing) ms |= 1u << i;
    }
    sched_out[st] = ms;
  }
  double vrel[3];
  for (int a = 0; a < 3; ++a) vrel[a] = rot_z[0 * 3 + a] * lin_vel[0] + rot_z[1 * 3 + a] * lin_vel[1] + rot_z[2 * 3 + a] * lin_vel[2];
  for (int i = 0; i < 4; ++i) {
    double delta_x = std::sqrt(std::abs(default_foot_pos[2 * 4 + 0]) / 9.8) * (vrel[0] - lin_vel_d[0]) +
                     ((counter_per_swing / gait_counter_speed[i]) * control_dt) / 2.0 * lin_vel_d[0];
    double delta_y = std::sqrt(std::abs(default_foot_pos[2 * 4 + 0]) / 9.8) * (vrel[1] - lin_vel_d[1]) +
                     ((counter_per_swing / gait_counter_speed[i]) * control_dt) / 2.0 * lin_vel_d[1];
    if (delta_x < -dx_lim) delta_x = -dx_lim;
    if (delta_x > dx_lim) delta_x = dx_lim;
    if (delta_y < -dy_lim) delta_y = -dy_lim;
    if (delta_y > dy_lim) delta_y = dy_lim;
    double f[3] = {default_foot_pos[0 * 4 + i] + delta_x, default_foot_pos[1 * 4 + i] + delta_y, default_foot_pos[2 * 4 + i]};
    for (int a = 0; a < 3; ++a) {
      double fa = rot[3 * a] * f[0] + rot[3 * a + 1] * f[1] + rot[3 * a + 2] * f[2];
      t_rel[3 * i + a] = f[a];
      t_abs[3 * i + a] = fa;
      t_world[3 * i + a] = fa + root_pos[a];
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Leg kinematics.  A1Kinematics::fk / jac (legKinematics/A1Kinematics.cpp:7-18) evaluate Matlab-generated
// expansions (:39-131) of the A1 leg chain.  Restated here AS THE CHAIN ITSELF -- hip at (ox, oy, 0), roll q0 about x, thigh
// offset d along y, pitch q1 about y, upper link lt down, pitch q2 about y, lower link lc down, contact offset (cx,cy,cz) --
// with the geometric Jacobian (axis x lever arm), i.e. independently of the closed form the CUDA kernel uses.
// PINNED: oracle/_ref (built in the authoring container from the reference's own A1Kinematics.cpp against a stub of the
// two Eigen types it touches) agrees with this function to 1e-15, and tests/golden/kinematics_v1.json holds vectors
// generated from that build.
// p[3], J[9] row-major (J[3*a + k] = d p_a / d q_k).
// ---------------------------------------------------------------------------------------------------------------------
static void rot_x(double t, double (&R)[3][3]) { const double c = std::cos(t), s = std::sin(t); double M[3][3] = {{1, 0, 0}, {0, c, -s}, {0, s, c}}; std::memcpy(R, M, sizeof(M)); }
static void rot_y(double t, double (&R)[3][3]) { const double c = std::cos(t), s = std::sin(t); double M[3][3] = {{c, 0, s}, {0, 1, 0}, {-s, 0, c}}; std::memcpy(R, M, sizeof(M)); }
static void mat_vec3(const double (&R)[3][3], const double* v, double* o) { for (int a = 0; a < 3; ++a) o[a] = R[a][0] * v[0] + R[a][1] * v[1] + R[a][2] * v[2]; }
static void mat_mat3(const double (&A)[3][3], const double (&B)[3][3], double (&C)[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}
static void cross3(const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }

int oracle_leg_kinematics(const double* q, const double* rho_opt, const double* rho_fix, double* p, double* J) {
  const double ox = rho_fix[0], oy = rho_fix[1], d = rho_fix[2], lt = rho_fix[3], lc = rho_fix[4];
  double Rx[3][3], Ry1[3][3], Ry2[3][3], R01[3][3], R012[3][3];
  rot_x(q[0], Rx); rot_y(q[1], Ry1); rot_y(q[2], Ry2);
  mat_mat3(Rx, Ry1, R01); mat_mat3(R01, Ry2, R012);
  const double hip[3] = {ox, oy, 0.0};
  const double off_thigh[3] = {0.0, d, 0.0}, link_up[3] = {0.0, 0.0, -lt}, tip[3] = {rho_opt[0], rho_opt[1], rho_opt[2] - lc};
  double t0[3], t1[3], t2[3];
  mat_vec3(Rx, off_thigh, t0); mat_vec3(R01, link_up, t1); mat_vec3(R012, tip, t2);
  double o1[3], o2[3];
  for (int a = 0; a < 3; ++a) { o1[a] = hip[a] + t0[a]; o2[a] = o1[a] + t1[a]; p[a] = o2[a] + t2[a]; }
  const double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
  double ax1[3];
  mat_vec3(Rx, ey, ax1);               // both pitch joints turn about the rolled y axis
  double l0[3], l1[3], l2[3], c0[3], c1[3], c2[3];
  for (int a = 0; a < 3; ++a) { l0[a] = p[a] - hip[a]; l1[a] = p[a] - o1[a]; l2[a] = p[a] - o2[a]; }
  cross3(ex, l0, c0); cross3(ax1, l1, c1); cross3(ax1, l2, c2);
  for (int a = 0; a < 3; ++a) { J[3 * a] = c0[a]; J[3 * a + 1] = c1[a]; J[3 * a + 2] = c2[a]; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// A1BasicEKF, literal restatement with the dense matrices of the reference: constructor :7-41 (C, Q, R, A, B),
// init_state :56-68, update_estimation :70-164.  x[18], P[18*18] row-major in/out.  The two linear solves with S (:135, :139,
// fullPivHouseholderQr in the reference) are done by Gaussian elimination with partial pivoting in long double: S is
// symmetric positive definite, any backward-stable solver returns the same vectors to rounding.  (parity unpinned: the
// reference's EKF needs Eigen and ROS headers and has no test vectors.)
// ---------------------------------------------------------------------------------------------------------------------
int oracle_ekf_init(const double* foot_pos_rel, const double* rot, double* x, double* P) {
  for (int i = 0; i < 18 * 18; ++i) P[i] = 0.0;
  for (int i = 0; i < 18; ++i) { P[i * 18 + i] = 3.0; x[i] = 0.0; }
  x[2] = 0.09;
  for (int leg = 0; leg < 4; ++leg)
    for (int a = 0; a < 3; ++a)
      x[6 + 3 * leg + a] = rot[3 * a] * foot_pos_rel[3 * leg] + rot[3 * a + 1] * foot_pos_rel[3 * leg + 1] + rot[3 * a + 2] * foot_pos_rel[3 * leg + 2] + x[a];
  return 0;
}

int oracle_ekf_update(double dt, int assume_flat_ground, uint32_t movement_mode, const double* imu_acc, const double* imu_ang_vel,
                      const double* rot, const double* foot_pos_rel, const double* foot_vel_rel, const double* foot_force, double* x, double* P,
                      double* root_pos, double* root_lin_vel, uint32_t* est_contacts) {
  constexpr int NX = 18, NY = 28;
  typedef long double LD;
  std::vector<LD> A(NX * NX, 0), Bm(NX * 3, 0), C(NY * NX, 0), Q(NX * NX, 0), Rn(NY * NY, 0);
  for (int i = 0; i < NX; ++i) A[i * NX + i] = 1;
  for (int i = 0; i < 4; ++i)
    for (int a = 0; a < 3; ++a) {
      C[(3 * i + a) * NX + a] = -1;                 // -pos
      C[(3 * i + a) * NX + 6 + 3 * i + a] = 1;      // foot pos
      C[(12 + 3 * i + a) * NX + 3 + a] = 1;         // vel
    }
  for (int i = 0; i < 4; ++i) C[(24 + i) * NX + 6 + 3 * i + 2] = 1;   // height z of foot
  for (int a = 0; a < 3; ++a) { A[a * NX + 3 + a] = dt; Bm[(3 + a) * 3 + a] = dt; }
  LD u[3];
  for (int a = 0; a < 3; ++a) u[a] = (LD)rot[3 * a] * imu_acc[0] + (LD)rot[3 * a + 1] * imu_acc[1] + (LD)rot[3 * a + 2] * imu_acc[2];
  u[2] += -9.81L;
  double ec[4];
  for (int i = 0; i < 4; ++i) ec[i] = (movement_mode == 0) ? 1.0 : std::min(std::max(foot_force[i] / (100.0 - 0.0), 0.0), 1.0);
  for (int a = 0; a < 3; ++a) { Q[a * NX + a] = 0.01L * dt / 20.0L; Q[(3 + a) * NX + 3 + a] = 0.01L * dt * 9.8L / 20.0L; }
  for (int i = 0; i < 4; ++i)
    for (int a = 0; a < 3; ++a) {
      const LD sc = 1 + (1 - (LD)ec[i]) * 1e3L;
      Q[(6 + 3 * i + a) * NX + 6 + 3 * i + a] = sc * dt * 0.01L;
      Rn[(3 * i + a) * NY + 3 * i + a] = sc * 0.001L;
      Rn[(12 + 3 * i + a) * NY + 12 + 3 * i + a] = sc * 0.1L;
    }
  for (int i = 0; i < 4; ++i) Rn[(24 + i) * NY + 24 + i] = assume_flat_ground ? (1 + (1 - (LD)ec[i]) * 1e3L) * 0.001L : 1e5L;
  // process update
  std::vector<LD> xbar(NX, 0), AP(NX * NX, 0), Pbar(NX * NX, 0);
  for (int i = 0; i < NX; ++i) {
    LD s = 0;
    for (int j = 0; j < NX; ++j) s += A[i * NX + j] * x[j];
    for (int j = 0; j < 3; ++j) s += Bm[i * 3 + j] * u[j];
    xbar[i] = s;
  }
  for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) { LD s = 0; for (int k = 0; k < NX; ++k) s += A[i * NX + k] * P[k * NX + j]; AP[i * NX + j] = s; }
  for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) { LD s = Q[i * NX + j]; for (int k = 0; k < NX; ++k) s += AP[i * NX + k] * A[j * NX + k]; Pbar[i * NX + j] = s; }
  // measurement
  std::vector<LD> yhat(NY, 0), y(NY, 0), ey(NY, 0);
  for (int r = 0; r < NY; ++r) { LD s = 0; for (int j = 0; j < NX; ++j) s += C[r * NX + j] * xbar[j]; yhat[r] = s; }
  for (int i = 0; i < 4; ++i) {
    const double* fk = foot_pos_rel + 3 * i;
    LD sk[3] = {(LD)imu_ang_vel[1] * fk[2] - (LD)imu_ang_vel[2] * fk[1], (LD)imu_ang_vel[2] * fk[0] - (LD)imu_ang_vel[0] * fk[2],
                (LD)imu_ang_vel[0] * fk[1] - (LD)imu_ang_vel[1] * fk[0]};
    LD lv[3];
    for (int a = 0; a < 3; ++a) lv[a] = -(LD)foot_vel_rel[3 * i + a] - sk[a];
    for (int a = 0; a < 3; ++a) {
      y[3 * i + a] = (LD)rot[3 * a] * fk[0] + (LD)rot[3 * a + 1] * fk[1] + (LD)rot[3 * a + 2] * fk[2];
      const LD rl = (LD)rot[3 * a] * lv[0] + (LD)rot[3 * a + 1] * lv[1] + (LD)rot[3 * a + 2] * lv[2];
      y[12 + 3 * i + a] = (1 - (LD)ec[i]) * x[3 + a] + (LD)ec[i] * rl;
    }
    y[24 + i] = (1 - (LD)ec[i]) * ((LD)x[2] + fk[2]) + (LD)ec[i] * 0;
  }
  for (int r = 0; r < NY; ++r) ey[r] = y[r] - yhat[r];
  std::vector<LD> CP(NY * NX, 0), S(NY * NY, 0), PCt(NX * NY, 0);
  for (int r = 0; r < NY; ++r) for (int j = 0; j < NX; ++j) { LD s = 0; for (int k = 0; k < NX; ++k) s += C[r * NX + k] * Pbar[k * NX + j]; CP[r * NX + j] = s; }
  for (int j = 0; j < NX; ++j) for (int r = 0; r < NY; ++r) { LD s = 0; for (int k = 0; k < NX; ++k) s += Pbar[j * NX + k] * C[r * NX + k]; PCt[j * NY + r] = s; }
  for (int r = 0; r < NY; ++r) for (int c = 0; c < NY; ++c) { LD s = Rn[r * NY + c]; for (int k = 0; k < NX; ++k) s += CP[r * NX + k] * C[c * NX + k]; S[r * NY + c] = s; }
  { std::vector<LD> St(S); for (int r = 0; r < NY; ++r) for (int c = 0; c < NY; ++c) S[r * NY + c] = 0.5L * (St[r * NY + c] + St[c * NY + r]); }
  // solve S [z | SC] = [ey | C]: elimination with partial pivoting on the augmented system
  const int NR = 1 + NX;
  std::vector<LD> M(NY * (NY + NR));
  for (int r = 0; r < NY; ++r) {
    for (int c = 0; c < NY; ++c) M[r * (NY + NR) + c] = S[r * NY + c];
    M[r * (NY + NR) + NY] = ey[r];
    for (int j = 0; j < NX; ++j) M[r * (NY + NR) + NY + 1 + j] = C[r * NX + j];
  }
  const int W = NY + NR;
  for (int c = 0; c < NY; ++c) {
    int pv = c;
    for (int r = c + 1; r < NY; ++r) if (fabsl(M[r * W + c]) > fabsl(M[pv * W + c])) pv = r;
    if (!(fabsl(M[pv * W + c]) > 0)) return 3;
    if (pv != c) for (int k = 0; k < W; ++k) std::swap(M[c * W + k], M[pv * W + k]);
    for (int r = c + 1; r < NY; ++r) {
      const LD m = M[r * W + c] / M[c * W + c];
      if (m != 0) for (int k = c; k < W; ++k) M[r * W + k] -= m * M[c * W + k];
    }
  }
  std::vector<LD> Z(NY * NR);
  for (int k = 0; k < NR; ++k)
    for (int r = NY - 1; r >= 0; --r) {
      LD s = M[r * W + NY + k];
      for (int c = r + 1; c < NY; ++c) s -= M[r * W + c] * Z[c * NR + k];
      Z[r * NR + k] = s / M[r * W + r];
    }
  // x = xbar + Pbar C' z ;  P = Pbar - Pbar C' SC Pbar ; symmetrise
  std::vector<LD> xn(NX), T1(NX * NX, 0), Pn(NX * NX, 0);
  for (int j = 0; j < NX; ++j) { LD s = xbar[j]; for (int r = 0; r < NY; ++r) s += PCt[j * NY + r] * Z[r * NR]; xn[j] = s; }
  for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) { LD s = 0; for (int r = 0; r < NY; ++r) s += PCt[i * NY + r] * Z[r * NR + 1 + j]; T1[i * NX + j] = s; }   // Pbar C' SC
  for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) { LD s = Pbar[i * NX + j]; for (int k = 0; k < NX; ++k) s -= T1[i * NX + k] * Pbar[k * NX + j]; Pn[i * NX + j] = s; }
  std::vector<LD> Ps(NX * NX);
  for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) Ps[i * NX + j] = 0.5L * (Pn[i * NX + j] + Pn[j * NX + i]);
  if (Ps[0] * Ps[NX + 1] - Ps[1] * Ps[NX] > 1e-6L) {
    for (int i = 0; i < 2; ++i) for (int j = 2; j < NX; ++j) { Ps[i * NX + j] = 0; Ps[j * NX + i] = 0; }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) Ps[i * NX + j] /= 10.0L;
  }
  for (int i = 0; i < NX; ++i) x[i] = (double)xn[i];
  for (int i = 0; i < NX * NX; ++i) P[i] = (double)Ps[i];
  if (est_contacts) { uint32_t m = 0; for (int i = 0; i < 4; ++i) m |= (ec[i] < 0.5 ? 0u : 1u) << i; *est_contacts = m; }
  if (root_pos) for (int a = 0; a < 3; ++a) root_pos[a] = x[a];
  if (root_lin_vel) for (int a = 0; a < 3; ++a) root_lin_vel[a] = x[3 + a];
  return 0;
}

// wall-clock timing of the reference-faithful path (build + OSQP default, cold start) on nthreads
double oracle_time_reference_path(const a1mpc_config* cfg, int B, const a1mpc_inputs* in, int nthreads, double* f_body) {
  auto t0 = std::chrono::steady_clock::now();
  oracle_compute_grf_batch(cfg, B, in, ORACLE_MODE_OSQP_DEFAULT, 0.0, nthreads, f_body, nullptr, nullptr);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

// ref_shim "OsqpEigen/OsqpEigen.h" -- TEST INFRASTRUCTURE.  NOT osqp-eigen, not part of the product.
//
// The call surface of OsqpEigen::Solver that the reference uses (A1RobotControl.cpp:416-439, :522-555, test/test_mpc.cpp:131-151),
// so that those files compile unmodified into oracle/_ref/ (osqp-eigen and OSQP are neither vendored by the reference nor
// installable offline).  The solver records the problem exactly as the reference hands it over -- the UPPER triangle of the
// Hessian (osqp-eigen's setHessianMatrix keeps triangularView<Upper>), constraint matrix, gradient, bounds -- and solve()
// forwards it to a hook installed by oracle/ref_mpc_wrap.cpp (which captures the problem for the golden vectors and lets the
// oracle's OSQP-algorithm restatement or exact solver produce the solution).  No arithmetic of its own.
#pragma once
#include <Eigen/Dense>
#include <memory>
#include <vector>

namespace OsqpEigen {
const double INFTY = 1e30;  // osqp's OSQP_INFTY

// n, m, P (n x n, column-major, symmetric completion of the upper triangle), q, A (m x n, column-major), l, u,
// warm start flag, x (in: previous primal if warm, out: primal), y (in/out: dual); returns 0 on success
typedef int (*SolveHook)(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u,
                         int warm, double* x, double* y);
inline SolveHook& solve_hook() { static SolveHook h = nullptr; return h; }

class Settings {
 public:
  void setVerbosity(bool) {}
  void setWarmStart(bool w) { warm = w; }
  bool warm = true;  // OSQP default warm_start = 1
};

class Data {
 public:
  void setNumberOfVariables(int n_) { n = n_; }
  void setNumberOfConstraints(int m_) { m = m_; }
  bool setHessianMatrix(const Eigen::SparseMatrix<double>& H) {
    if (H.rows() != n || H.cols() != n) return false;
    P.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i <= j; ++i) { P[(size_t)j * n + i] = H.coeff(i, j); P[(size_t)i * n + j] = H.coeff(i, j); }
    return true;
  }
  bool setLinearConstraintsMatrix(const Eigen::SparseMatrix<double>& Ac) {
    if (Ac.rows() != m || Ac.cols() != n) return false;
    A.assign((size_t)m * n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) A[(size_t)j * m + i] = Ac.coeff(i, j);
    return true;
  }
  template <class V> bool setGradient(const Eigen::MatrixBase<V>& g) { return copy(g, q, n); }
  template <class V> bool setLowerBound(const Eigen::MatrixBase<V>& v) { return copy(v, l, m); }
  template <class V> bool setUpperBound(const Eigen::MatrixBase<V>& v) { return copy(v, u, m); }
  int n = 0, m = 0;
  std::vector<double> P, A, q, l, u;
 private:
  template <class V> static bool copy(const Eigen::MatrixBase<V>& v, std::vector<double>& d, int len) {
    if (v.size() != len) return false;
    d.resize((size_t)len);
    for (int k = 0; k < len; ++k) d[(size_t)k] = v.coeff(k);
    return true;
  }
};

class Solver {
 public:
  Solver() : settings_(new Settings), data_(new Data) {}
  const std::unique_ptr<Settings>& settings() const { return settings_; }
  const std::unique_ptr<Data>& data() const { return data_; }
  bool isInitialized() const { return init_; }
  bool initSolver() {
    if (data_->P.empty() || data_->A.empty() || data_->q.empty() || data_->l.empty() || data_->u.empty()) return false;
    x_.assign((size_t)data_->n, 0.0);
    y_.assign((size_t)data_->m, 0.0);
    init_ = true;
    have_prev_ = false;
    return true;
  }
  bool updateHessianMatrix(const Eigen::SparseMatrix<double>& H) { return data_->setHessianMatrix(H); }
  template <class V> bool updateGradient(const Eigen::MatrixBase<V>& g) { return data_->setGradient(g); }
  template <class V> bool updateLowerBound(const Eigen::MatrixBase<V>& v) { return data_->setLowerBound(v); }
  template <class V> bool updateUpperBound(const Eigen::MatrixBase<V>& v) { return data_->setUpperBound(v); }
  bool solve() {
    if (!init_ || !solve_hook()) return false;
    int warm = settings_->warm && have_prev_;
    if (!warm) { x_.assign(x_.size(), 0.0); y_.assign(y_.size(), 0.0); }
    int rc = solve_hook()(data_->n, data_->m, data_->P.data(), data_->q.data(), data_->A.data(), data_->l.data(), data_->u.data(),
                          warm, x_.data(), y_.data());
    have_prev_ = (rc == 0);
    return rc == 0;
  }
  Eigen::VectorXd getSolution() const {
    Eigen::VectorXd s((int)x_.size());
    for (size_t k = 0; k < x_.size(); ++k) s(k) = x_[k];
    return s;
  }
 private:
  std::unique_ptr<Settings> settings_;
  std::unique_ptr<Data> data_;
  std::vector<double> x_, y_;
  bool init_ = false, have_prev_ = false;
};
}  // namespace OsqpEigen

// oracle/ref_test_mpc_hook.cpp -- TEST INFRASTRUCTURE.  Linked next to the reference's unmodified test/test_mpc.cpp
// (oracle/_ref/ref_test_mpc): installs, before main() runs, the solver that the OsqpEigen stand-in forwards to -- the oracle's
// OSQP-algorithm restatement run to eps 1e-11 (OSQP itself is absent offline).
#include "OsqpEigen/OsqpEigen.h"
extern "C" int oracle_qp_hook_osqp_tight(int n, int m, const double* P, const double* q, const double* A, const double* l,
                                         const double* u, int warm, double* x, double* y);
namespace {
struct Install { Install() { OsqpEigen::solve_hook() = oracle_qp_hook_osqp_tight; } } install_;
}

// emu_driver.cpp -- runs the UNCHANGED device code of a1mpc_device.cuh (pack + fused solve kernels) on the CPU block
// emulator of cuda_emu.h and exposes it to the tests through a small C interface.  TEST INFRASTRUCTURE ONLY: this is a
// checker for the warp-level algorithms (shuffle/DMMA fragment maps, shared-memory layouts, barriers) on machines
// without a GPU.  It mirrors enqueue_solve() of a1mpc_api.cu (same kernels, same template classes, same records).
#define A1MPC_EMU 1
#include "cuda_emu.h"

#include <atomic>
#include <thread>

#include "../../a1-qp-mpc-controller_b200/csrc/a1mpc_misc.cuh"
#include "../../a1-qp-mpc-controller_b200/csrc/a1mpc_estim.cuh"
#include "../../a1-qp-mpc-controller_b200/csrc/a1mpc_sched.cuh"
#include "../../a1-qp-mpc-controller_b200/csrc/a1mpc_dense.cu"   // kernels only (host launchers are compiled out under A1MPC_EMU)

using namespace a1mpc;

namespace {

DevParams make_params(const a1mpc_config* cfg) {   // a1mpc_create() in a1mpc_api.cu
  DevParams P;
  P.N = cfg->horizon;
  P.max_iter = cfg->max_iter > 0 ? cfg->max_iter : 40;
  P.dt = cfg->dt; P.mu = cfg->mu; P.fzmax = cfg->fz_max; P.mass = cfg->mass;
  P.mu_switch = cfg->tol > 0.0 ? cfg->tol : MU_SWITCH_DEFAULT;
  for (int i = 0; i < 9; ++i) P.inertia[i] = cfg->inertia[i];
  for (int i = 0; i < 13; ++i) P.q2[i] = 2.0 * cfg->q[i];
  for (int i = 0; i < 12; ++i) P.r2[i] = 2.0 * cfg->r[i];
  return P;
}

struct Stats { std::atomic<unsigned long> collectives{0}, mma{0}; };

template <int NS, int N, int WPC, int LSM, bool EXT>
void run_class(const DevParams& P, const double* rec, const int* count, const DevOutputs& out, int order_mode, int nthreads, Stats& st,
               uint32_t* warm = nullptr, int shift = 0) {
  using G = Geo<NS, N, LSM>;
  const int nq = count[EXT ? 5 : NS];
  if (nq == 0) return;
  const int grid = std::max(1, (nq + 2 * WPC - 1) / (2 * WPC));   // ~2 QPs per warp: exercises the persistent loop (and the CTA rendezvous)
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (;;) {
      const int bx = next.fetch_add(1);
      if (bx >= grid) break;
      unsigned long nc = 0, nm = 0;
      a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)grid, 1, 1}, 32 * WPC * G::TW, G::smem_bytes(WPC), order_mode,
                       [&]() {
                         if constexpr (!EXT) {
                           if (warm) { solve_kernel_warm<NS, N, WPC, LSM>(P, rec, count, out, warm, shift); return; }
                         }
                         solve_kernel<NS, N, WPC, LSM, EXT>(P, rec, count, out);
                       }, &nc, &nm);
      st.collectives += nc;
      st.mma += nm;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
}

template <int N>
void run_all(const DevParams& P, const double* rec, size_t cap, const int* count, const DevOutputs& out, int order_mode, int nthreads, Stats& st,
             uint32_t* warm, int shift);
template <>
void run_all<10>(const DevParams& P, const double* rec, size_t cap, const int* count, const DevOutputs& out, int order_mode, int nthreads, Stats& st,
                 uint32_t* warm, int shift) {
  run_class<1, 10, A1MPC_WPC1, 0, false>(P, rec + 0 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st, warm, shift);
  run_class<2, 10, A1MPC_WPC2, 0, false>(P, rec + 1 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st, warm, shift);
  run_class<3, 10, A1MPC_WPC34, 1, false>(P, rec + 2 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st, warm, shift);
  run_class<4, 10, A1MPC_WPC34, 1, false>(P, rec + 3 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st, warm, shift);
}
template <>
void run_all<20>(const DevParams& P, const double* rec, size_t cap, const int* count, const DevOutputs& out, int order_mode, int nthreads, Stats& st,
                 uint32_t*, int) {
  run_class<1, 20, 2, 0, false>(P, rec + 0 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st);
  run_class<2, 20, 1, 0, false>(P, rec + 1 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st);
  run_class<3, 20, 1, 1, false>(P, rec + 2 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st);
  run_class<4, 20, 1, 1, false>(P, rec + 3 * cap * REC_DOUBLES, count, out, order_mode, nthreads, st);
}

template <int NS, int N>
void run_dense(const DevParams& P, int B, const double* H, const double* g, const uint32_t* contact, const int* list, const int* count,
               double* u, int32_t* status, int order_mode) {
  const int nq = count[NS];
  for (int bx = 0; bx < nq; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)nq, 1, 1}, 32 * Geo<NS, N>::TW, DenseGeo<NS, N>::smem_bytes(), order_mode,
                     [&]() { dense_solve_kernel<NS, N>(P, H, g, contact, list + (size_t)(NS - 1) * B, count, u, status); });
}

template <int NS>
void run_grf(const DevParams& P, int B, const double* root_acc, const double* rot_z, const double* rot, const double* foot,
             const uint32_t* contact, const int* list, const int* count, double* f_body, int32_t* status, int order_mode) {
  const int nq = count[NS];
  for (int bx = 0; bx < nq; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)nq, 1, 1}, 32, DenseGeo<NS, 1>::smem_bytes() + 72 * 8, order_mode,
                     [&]() { grf_qp_kernel<NS>(P, root_acc, rot_z, rot, foot, contact, list + (size_t)(NS - 1) * B, count, f_body, status); });
}

void classify(int B, const uint32_t* contact, int* list, int* count, double* out, int out_per_qp, int32_t* status) {
  const int pb = 128, pgrid = (B + pb - 1) / pb;
  for (int bx = 0; bx < pgrid; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                     [&]() { classify_kernel(B, contact, list, count, out, out_per_qp, status); });
}

}  // namespace

extern "C" {

// a1mpc_solve_dense_batch on the emulator (QP-major H [B,n,n], g [B,n], u [B,n]); N = 10 only
int emu_solve_dense(const a1mpc_config* cfg, int B, const double* H, const double* g, const uint32_t* contact, double* u, int32_t* status,
                    int order_mode) {
  if (cfg->horizon != 10 && cfg->horizon != 20) return -1;
  const DevParams P = make_params(cfg);
  std::vector<int> list((size_t)4 * B + 8, 0);
  int count[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0..7] class counts, [8..15] queue counters (next_qp)
  classify(B, contact, list.data(), count, u, 12 * cfg->horizon, status);
  if (cfg->horizon == 20) {   // a1mpc_dense.cu serves the direct classes at N = 20 (one or two stance feet)
    run_dense<2, 20>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
    run_dense<1, 20>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
    return 0;
  }
  run_dense<4, 10>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
  run_dense<3, 10>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
  run_dense<2, 10>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
  run_dense<1, 10>(P, B, H, g, contact, list.data(), count, u, status, order_mode);
  return 0;
}

// a1mpc_grf_qp_batch on the emulator (grf_qp_launch of a1mpc_dense.cu)
int emu_grf_qp(int B, const double* root_acc, const double* rot_z, const double* rot, const double* foot, const uint32_t* contact,
               double* f_body, int32_t* status, int order_mode) {
  DevParams P;
  std::memset(&P, 0, sizeof(P));
  P.N = 1; P.max_iter = 40; P.mu = 0.7; P.fzmax = 180.0; P.mu_switch = 1e-9;
  std::vector<int> list((size_t)4 * B + 8, 0);
  int count[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0..7] class counts, [8..15] queue counters (next_qp)
  classify(B, contact, list.data(), count, f_body, 12, status);
  run_grf<4>(P, B, root_acc, rot_z, rot, foot, contact, list.data(), count, f_body, status, order_mode);
  run_grf<3>(P, B, root_acc, rot_z, rot, foot, contact, list.data(), count, f_body, status, order_mode);
  run_grf<2>(P, B, root_acc, rot_z, rot, foot, contact, list.data(), count, f_body, status, order_mode);
  run_grf<1>(P, B, root_acc, rot_z, rot, foot, contact, list.data(), count, f_body, status, order_mode);
  return 0;
}

// the compacted two-feet-per-step kernel of a1mpc_sched.cuh; every QP of the batch must have exactly two stance feet in every step
int emu_solve_sched2(const a1mpc_config* cfg, int B, const a1mpc_inputs* in, const uint32_t* sched, const double* normals,
                     const a1mpc_outputs* o, int order_mode, int nthreads) {
  if (cfg->horizon != 10 || !sched) return -1;
  for (int st = 0; st < 10; ++st)
    for (int b = 0; b < B; ++b)
      if (__builtin_popcount(sched[(size_t)st * in->ld + b] & 15u) != 2) return -2;
  const DevParams P = make_params(cfg);
  const DevInputs din{in->x0, in->rot, in->foot, in->ref, in->contact, in->ld};
  const DevOutputs dout{o->f_body, o->status, o->iters, o->u_full, o->ld};
  int count[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0..7] class counts, [8..15] queue counters (next_qp)
  std::vector<double> rec((size_t)B * REC_EXT_DOUBLES + 2);
  const int pb = 128, pgrid = (B + pb - 1) / pb;
  for (int bx = 0; bx < pgrid; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                     [&]() { pack_ext_kernel(din, sched, normals, B, rec.data(), count, dout, cfg->horizon); });
  count[6] = count[5];
  constexpr int WPC = 4;
  const int nq = count[6], grid = std::max(1, (nq + 2 * WPC - 1) / (2 * WPC));
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (;;) {
      const int bx = next.fetch_add(1);
      if (bx >= grid) break;
      a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)grid, 1, 1}, 32 * WPC, SchedGeo<10>::smem_bytes(WPC), order_mode,
                       [&]() { solve_kernel_sched2<10, WPC>(P, rec.data(), count, dout); });
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < std::max(1, nthreads); ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return 0;
}

// a1mpc_leg_kinematics_batch on the emulator
int emu_leg_kinematics(int B, const double* joint_pos, const double* joint_vel, const double* rot, const double* rho_opt, const double* rho_fix,
                       double* foot_pos_rel, double* jac, double* foot_vel_rel, double* foot_pos_abs, double* foot_vel_abs) {
  LegParams P;
  for (int i = 0; i < 12; ++i) P.rho_opt[i] = rho_opt[i];
  for (int i = 0; i < 20; ++i) P.rho_fix[i] = rho_fix[i];
  const int pb = 128, pgrid = (B + pb - 1) / pb;
  for (int bx = 0; bx < pgrid; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                     [&]() { leg_kinematics_kernel(B, joint_pos, joint_vel, rot, P, foot_pos_rel, jac, foot_vel_rel, foot_pos_abs, foot_vel_abs); });
  return 0;
}

// a1mpc_ekf_init_batch / a1mpc_ekf_update_batch on the emulator (state: B x 342 doubles on the host)
int emu_ekf_init(int B, double* state, const double* foot_pos_rel, const double* rot) {
  const int pb = 128, pgrid = (B + pb - 1) / pb;
  for (int bx = 0; bx < pgrid; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                     [&]() { ekf_init_kernel(B, state, foot_pos_rel, rot); });
  return 0;
}

int emu_ekf_update(int B, double* state, double dt, int assume_flat_ground, const uint32_t* movement_mode, const double* imu_acc,
                   const double* imu_ang_vel, const double* rot, const double* foot_pos_rel, const double* foot_vel_rel, const double* foot_force,
                   double* root_pos, double* root_lin_vel, uint32_t* est_contacts, int32_t* status, int order_mode) {
  EkfParams P{dt, assume_flat_ground ? 1 : 0};
  const int grid = std::max(1, (B + 2 * EKF_WPC - 1) / (2 * EKF_WPC));   // ~2 robots per warp: exercises the persistent loop
  for (int bx = 0; bx < grid; ++bx)
    a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)grid, 1, 1}, 32 * EKF_WPC, (size_t)EKF_WPC * EKF_WARP_DOUBLES * 8, order_mode,
                     [&]() {
                       ekf_update_kernel(B, P, state, movement_mode, imu_acc, imu_ang_vel, rot, foot_pos_rel, foot_vel_rel, foot_force, root_pos,
                                         root_lin_vel, est_contacts, status);
                     });
  return 0;
}

// Same contract as a1mpc_solve_batch / a1mpc_solve_batch_ext with host pointers.  order_mode: lane order between
// collectives (0 ascending, 1 descending, 2 pseudo-random).  stats[0] = warp collectives executed, stats[1] = DMMAs.
// warm: host buffer of B * (4 + 4 * horizon) u32 (zero = no guess), N = 10 only -- the emulated a1mpc_solve_batch_warm
int emu_solve_batch(const a1mpc_config* cfg, int B, const a1mpc_inputs* in, const uint32_t* sched, const double* normals,
                    const a1mpc_outputs* o, int order_mode, int nthreads, unsigned long* stats, uint32_t* warm, int shift) {
  if (cfg->horizon != 10 && cfg->horizon != 20) return -1;
  const DevParams P = make_params(cfg);
  const DevInputs din{in->x0, in->rot, in->foot, in->ref, in->contact, in->ld};
  const DevOutputs dout{o->f_body, o->status, o->iters, o->u_full, o->ld};
  const size_t cap = (size_t)B;
  int count[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0..7] class counts, [8..15] queue counters (next_qp)
  Stats st;
  const bool ext = sched || normals;
  std::vector<double> rec(cap * (ext ? (size_t)REC_EXT_DOUBLES : 4 * (size_t)REC_DOUBLES) + 2);
  const int pb = 128, pgrid = (B + pb - 1) / pb;
  for (int bx = 0; bx < pgrid; ++bx) {
    if (!ext)
      a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                       [&]() { pack_kernel(din, B, rec.data(), (int)cap, count, dout, cfg->horizon); });
    else
      a1emu::run_block(a1emu::Dim3{(unsigned)bx, 0, 0}, a1emu::Dim3{(unsigned)pgrid, 1, 1}, pb, 0, 0,
                       [&]() { pack_ext_kernel(din, sched, normals, B, rec.data(), count, dout, cfg->horizon); });
  }
  if (nthreads < 1) nthreads = 1;
  if (!ext) {
    if (warm && cfg->horizon != 10) return -2;
    if (cfg->horizon == 10) run_all<10>(P, rec.data(), cap, count, dout, order_mode, nthreads, st, warm, shift);
    else run_all<20>(P, rec.data(), cap, count, dout, order_mode, nthreads, st, nullptr, 0);
  } else {
    if (cfg->horizon == 10) run_class<4, 10, A1MPC_WPC34, 1, true>(P, rec.data(), count, dout, order_mode, nthreads, st);
    else run_class<4, 20, 1, 1, true>(P, rec.data(), count, dout, order_mode, nthreads, st);
  }
  if (stats) { stats[0] = st.collectives; stats[1] = st.mma; }
  return 0;
}

}  // extern "C"

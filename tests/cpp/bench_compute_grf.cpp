// bench_compute_grf.cpp -- times the plugin-level call a user of the reference would make, A1RobotControlBatch::compute_grf
// (host/A1MpcBatch.h, the batched mirror of A1RobotControl::compute_grf, A1RobotControl.h:44), with PAGEABLE std::vector
// states: AoS -> SoA repack on the host, H2D copies from pageable memory, pack + solve kernels, D2H, SoA -> AoS.
// Prints one JSON object; bench.py runs it on rank 0 and reports it next to the pinned-buffer C-ABI figure.
//   usage: bench_compute_grf [B=1024] [steps=200] [warmup=5]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "A1MpcBatch.h"

using namespace a1mpc_host;

int main(int argc, char** argv) {
  const int B = argc > 1 ? std::atoi(argv[1]) : 1024, K = argc > 2 ? std::atoi(argv[2]) : 200, W = argc > 3 ? std::atoi(argv[3]) : 5;
  a1mpc_config cfg;
  a1mpc_default_config(&cfg);
  const int NR = 8;   // distinct batches, cycled
  std::vector<std::vector<A1CtrlStatesLite>> rings(NR, std::vector<A1CtrlStatesLite>(B));
  std::vector<double> x0(12 * (size_t)B), rot(9 * (size_t)B), foot(12 * (size_t)B), ref(9 * (size_t)B);
  std::vector<uint32_t> contact(B);
  for (int r = 0; r < NR; ++r) {
    if (a1mpc_gen_states(2, 9000 + r, B, x0.data(), rot.data(), foot.data(), ref.data(), contact.data()) != A1MPC_OK) return 2;
    for (int b = 0; b < B; ++b) {
      A1CtrlStatesLite& s = rings[r][b];
      for (int k = 0; k < 3; ++k) {
        s.root_euler[k] = x0[(size_t)k * B + b]; s.root_pos[k] = x0[(size_t)(3 + k) * B + b];
        s.root_ang_vel[k] = x0[(size_t)(6 + k) * B + b]; s.root_lin_vel[k] = x0[(size_t)(9 + k) * B + b];
        s.root_ang_vel_d[k] = ref[(size_t)(2 + k) * B + b]; s.root_lin_vel_d[k] = ref[(size_t)(5 + k) * B + b];
      }
      for (int k = 0; k < 9; ++k) s.root_rot_mat[k] = rot[(size_t)k * B + b];
      for (int leg = 0; leg < 4; ++leg)
        for (int a = 0; a < 3; ++a) s.foot_pos_abs[a * 4 + leg] = foot[(size_t)(3 * leg + a) * B + b];
      s.root_euler_d[0] = ref[b]; s.root_euler_d[1] = ref[(size_t)B + b]; s.root_pos_d[2] = ref[(size_t)8 * B + b];
      for (int i = 0; i < 4; ++i) s.contacts[i] = (contact[b] >> i) & 1u;
    }
  }
  A1RobotControlBatch ctrl(cfg.mass, cfg.inertia, cfg.q, cfg.r, cfg.horizon, 0);
  std::vector<std::array<double, 12>> grf;
  std::vector<int32_t> status;
  long bad = 0;
  for (int i = 0; i < W; ++i) ctrl.compute_grf(rings[i % NR], cfg.dt, grf, &status);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < K; ++i) {
    ctrl.compute_grf(rings[i % NR], cfg.dt, grf, &status);
    for (int32_t s : status) bad += (s != 0);
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("{\"call\": \"A1RobotControlBatch::compute_grf (pageable std::vector states, AoS<->SoA repack included)\", \"batch\": %d, \"steps\": %d, "
              "\"value\": %.1f, \"unit\": \"QPs/s\", \"ms_per_call\": %.4f, \"non_optimal\": %ld, \"timing\": \"host wall clock (the call is synchronous)\"}\n",
              B, K, (double)B * K / sec, 1e3 * sec / K, bad);
  return bad == 0 ? 0 : 1;
}

// a1mpc_gen.cpp -- deterministic synthetic workload of the benchmark.  Host only.
// Distribution: A1 trot-gait states around a 0.30 m stance; limits from A1Params.h:19-21, 44-45 and the
// default footholds of config/gazebo_a1_mpc.yaml:18-32.
#include <cmath>
#include <cstdint>

#include "../../include/a1mpc.h"

namespace {
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
  double normal(double sigma) {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};
}  // namespace

extern "C" int a1mpc_gen_states(int config_id, uint64_t stream, int B, double* x0, double* rot, double* foot, double* ref,
                                uint32_t* contact) {
  if (B <= 0 || !x0 || !rot || !foot || !ref || !contact) return A1MPC_EINVAL;
  const bool wide = (config_id == 4);
  const uint64_t seed = 0xA1C0FFEEull + (uint64_t)config_id + stream * 0x100000001B3ull;
  const size_t ld = (size_t)B;
  for (int b = 0; b < B; ++b) {
    SplitMix64 rng(seed ^ ((uint64_t)(b + 1) * 0xD1342543DE82EF95ull));
    const double yaw = rng.uniform(-3.141592653589793, 3.141592653589793);
    const double roll = rng.normal(0.02), pitch = rng.normal(0.02);
    const double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch), cy = std::cos(yaw), sy = std::sin(yaw);
    // R = Rz * Ry * Rx
    const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                         sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                         -sp,     cp * sr,                cp * cr};
    const double px = rng.normal(1.0), py = rng.normal(1.0);
    const double pz = wide ? rng.uniform(0.25, 0.32) : rng.uniform(0.296, 0.304);
    const double w[3] = {rng.normal(0.1), rng.normal(0.1), rng.normal(0.1)};
    const double vc[3] = {rng.uniform(-0.6, 0.6), rng.uniform(-0.3, 0.3), 0.0};
    const double yawrate_d = rng.uniform(-0.8, 0.8);
    double v[3];
    for (int a = 0; a < 3; ++a) v[a] = R[3 * a] * vc[0] + R[3 * a + 1] * vc[1] + R[3 * a + 2] * vc[2] + rng.normal(0.03);
    x0[0 * ld + b] = roll; x0[1 * ld + b] = pitch; x0[2 * ld + b] = yaw;
    x0[3 * ld + b] = px; x0[4 * ld + b] = py; x0[5 * ld + b] = pz;
    for (int a = 0; a < 3; ++a) { x0[(6 + a) * ld + b] = w[a]; x0[(9 + a) * ld + b] = v[a]; }
    for (int k = 0; k < 9; ++k) rot[k * ld + b] = R[k];
    const double dx[4] = {0.17, 0.17, -0.17, -0.17}, dy[4] = {0.15, -0.15, 0.15, -0.15};
    for (int leg = 0; leg < 4; ++leg) {
      const double f[3] = {dx[leg] + rng.uniform(-0.1, 0.1), dy[leg] + rng.uniform(-0.1, 0.1), -0.35 + rng.normal(0.01)};
      for (int a = 0; a < 3; ++a) foot[(3 * leg + a) * ld + b] = R[3 * a] * f[0] + R[3 * a + 1] * f[1] + R[3 * a + 2] * f[2];
    }
    ref[0 * ld + b] = 0.0; ref[1 * ld + b] = 0.0;
    ref[2 * ld + b] = 0.0; ref[3 * ld + b] = 0.0; ref[4 * ld + b] = yawrate_d;
    ref[5 * ld + b] = vc[0]; ref[6 * ld + b] = vc[1]; ref[7 * ld + b] = 0.0;
    ref[8 * ld + b] = 0.30;
    const double u = rng.uniform();
    contact[b] = (u < 0.45) ? 0b1001u : (u < 0.90 ? 0b0110u : 0b1111u);  // {FL,RR} | {FR,RL} | all four
  }
  return A1MPC_OK;
}

extern "C" int a1mpc_gen_schedule(int config_id, uint64_t stream, int B, int horizon, uint32_t* sched, double* normals) {
  if (B <= 0 || horizon <= 0 || horizon > A1MPC_MAX_HORIZON || !sched || !normals) return A1MPC_EINVAL;
  const uint64_t seed = 0x5C4ED0000ull + 0xA1C0FFEEull + (uint64_t)config_id + stream * 0x100000001B3ull;
  const size_t ld = (size_t)B;
  for (int b = 0; b < B; ++b) {
    SplitMix64 rng(seed ^ ((uint64_t)(b + 1) * 0xD1342543DE82EF95ull));
    const int gait = (int)(rng.next() % 3ull);       // 0 trot, 1 bound, 2 rotary gallop
    const int phase0 = (int)(rng.next() % 16ull);
    for (int st = 0; st < horizon; ++st) {
      const int p = (phase0 + st) & 15;
      uint32_t m;
      if (gait == 0) m = (p < 8) ? 0b1001u : 0b0110u;            // {FL,RR} / {FR,RL}
      else if (gait == 1) m = (p < 8) ? 0b0011u : 0b1100u;       // {FL,FR} / {RL,RR}
      else {
        const int off[4] = {0, 4, 12, 8};                          // FL, FR, RL, RR touch down a quarter period apart (rotary)
        m = 0;
        for (int leg = 0; leg < 4; ++leg)
          if (((p + 16 - off[leg]) & 15) < 8) m |= 1u << leg;
      }
      sched[(size_t)st * ld + b] = m;
    }
    for (int leg = 0; leg < 4; ++leg) {
      const double ang = rng.normal(0.2), az = rng.uniform(0.0, 6.283185307179586);
      // rotate z by `ang` about the horizontal axis (cos az, sin az, 0)
      const double ax = std::cos(az), ay = std::sin(az);
      normals[(size_t)(3 * leg) * ld + b] = ay * std::sin(ang);
      normals[(size_t)(3 * leg + 1) * ld + b] = -ax * std::sin(ang);
      normals[(size_t)(3 * leg + 2) * ld + b] = std::cos(ang);
    }
  }
  return A1MPC_OK;
}

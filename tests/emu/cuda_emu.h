// cuda_emu.h -- lane-accurate CPU emulation of one CUDA thread block.  TEST INFRASTRUCTURE ONLY.
//
// The device code of a1-qp-mpc-controller_b200/csrc/a1mpc_device.cuh is compiled UNCHANGED by g++ against this
// shim (A1MPC_EMU) so that the warp-level algorithms (shuffles, ballots, DMMA fragments, shared-memory layouts) can be
// checked against the oracle on a machine without a GPU.  It is not a CPU fallback: nothing in the product links it,
// it is ~10^4 times slower than the GPU, and it exists only under tests/.
//
// Model: every CUDA thread is a fibre (own stack, hand-written context switch) on ONE host thread per block.  A fibre
// runs until it reaches a warp collective (__shfl_sync, __any_sync, __syncwarp, mma, ...) or __syncthreads, publishes
// its operand and yields; the last lane to arrive releases the warp.  Between two collectives the lanes of a warp run
// one after the other in a configurable order (ascending / descending / pseudo-random): code that relies on implicit
// lock-step (a missing __syncwarp around shared memory) gives order-dependent results, which the tests detect by
// comparing the orders bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static thread_local   // one block per host thread at a time: shared by all fibres of the block

namespace a1emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = false;
};

struct Warp {
  unsigned gen = 0;
  int count = 0, nlanes = 32;
  uint64_t slot[2][32];
  uint64_t slot2[2][32];
};

struct NamedBarrier { unsigned gen = 0; int count = 0, acc_or = 0, acc_and = 1, res_or[2] = {0, 0}, res_and[2] = {1, 1}; };

struct Block {
  NamedBarrier nbar[16];
  int nthreads = 0;
  std::vector<Fiber> fib;
  std::vector<Warp> warps;
  unsigned bgen = 0;
  int bcount = 0;
  void* sched_sp = nullptr;
  int cur = 0;
  unsigned long progress = 0;
  std::function<void()> body;
  std::vector<double> smem;
  int order_mode = 0;       // 0 ascending, 1 descending, 2 pseudo-random
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  unsigned long n_collectives = 0, n_mma = 0;
};

extern thread_local Block* g_blk;
extern thread_local int g_f32;   // experiment (A1MPC_EMU_F32): > 0 while a factorisation runs whose arithmetic is rounded to fp32
// round to A1EMU_MANT mantissa bits (environment; default 23 = fp32, 10 = tf32, 7 = bf16) -- experiment only
inline double lowp(double x) {
  static const int mant = std::getenv("A1EMU_MANT") ? std::atoi(std::getenv("A1EMU_MANT")) : 23;
  if (mant >= 23) return (double)(float)x;
  if (x == 0.0 || !(x == x)) return x;
  int e;
  const double m = std::frexp(x, &e);
  return std::ldexp(std::nearbyint(std::ldexp(m, mant + 1)), e - mant - 1);
}
void yield_to_scheduler();
// runs `body` once per thread of one block (threadIdx/blockDim set), with dynamic shared memory of smem_bytes
void run_block(Dim3 block_idx, Dim3 grid_dim, int nthreads, size_t smem_bytes, int order_mode, const std::function<void()>& body,
               unsigned long* n_collectives = nullptr, unsigned long* n_mma = nullptr);

}  // namespace a1emu

extern thread_local a1emu::Dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace a1emu {

inline int lane_id() { return (int)(threadIdx.x & 31u); }
inline Warp& my_warp() { return g_blk->warps[threadIdx.x >> 5]; }

// publish (v, v2), wait for the whole warp, return the generation whose slots hold this collective's operands
inline unsigned warp_arrive(uint64_t v, uint64_t v2 = 0) {
  Warp& w = my_warp();
  const unsigned g = w.gen;
  const int lane = lane_id();
  w.slot[g & 1][lane] = v;
  w.slot2[g & 1][lane] = v2;
  if (++w.count == w.nlanes) {
    w.count = 0;
    w.gen = g + 1;
    ++g_blk->progress;
    ++g_blk->n_collectives;
  } else {
    while (w.gen == g) yield_to_scheduler();
  }
  return g;
}

inline uint64_t d2u(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
inline double u2d(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }

}  // namespace a1emu

// ------------------------------------------------------------------------------------------------
// warp collectives (full-mask forms only: that is all the device code uses)
// ------------------------------------------------------------------------------------------------
inline void __syncwarp(unsigned = 0xffffffffu) { a1emu::warp_arrive(0); }
inline double __shfl_sync(unsigned, double v, int src) {
  a1emu::Warp& w = a1emu::my_warp();
  const unsigned g = a1emu::warp_arrive(a1emu::d2u(v));
  return a1emu::u2d(w.slot[g & 1][src & 31]);
}
inline int __shfl_sync(unsigned, int v, int src) {
  a1emu::Warp& w = a1emu::my_warp();
  const unsigned g = a1emu::warp_arrive((uint64_t)(uint32_t)v);
  return (int)(uint32_t)w.slot[g & 1][src & 31];
}
inline double __shfl_xor_sync(unsigned m, double v, int x) { return __shfl_sync(m, v, a1emu::lane_id() ^ x); }
inline int __shfl_xor_sync(unsigned m, int v, int x) { return __shfl_sync(m, v, a1emu::lane_id() ^ x); }
inline unsigned __ballot_sync(unsigned, int pred) {
  a1emu::Warp& w = a1emu::my_warp();
  const unsigned g = a1emu::warp_arrive(pred ? 1u : 0u);
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (w.slot[g & 1][l] ? 1u : 0u) << l;
  return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline int __reduce_add_sync(unsigned, int v) {
  a1emu::Warp& w = a1emu::my_warp();
  const unsigned g = a1emu::warp_arrive((uint64_t)(uint32_t)v);
  int s = 0;
  for (int l = 0; l < 32; ++l) s += (int)(uint32_t)w.slot[g & 1][l];
  return s;
}
namespace a1emu {
// PTX bar.sync / bar.red.{or,and}.pred on named barrier `id` over `nthreads` threads: mode 0 barrier only, 1 returns the OR of the
// predicates, 2 their AND.  Every participating thread must use the same mode (as on the GPU).
inline int named_barrier(int id, int nthreads, int mode, int pred) {
  NamedBarrier& nb = g_blk->nbar[id & 15];
  const unsigned g = nb.gen;
  nb.acc_or |= pred ? 1 : 0;
  nb.acc_and &= pred ? 1 : 0;
  if (++nb.count == nthreads) {
    nb.res_or[g & 1] = nb.acc_or; nb.res_and[g & 1] = nb.acc_and;
    nb.count = 0; nb.acc_or = 0; nb.acc_and = 1;
    nb.gen = g + 1;
    ++g_blk->progress;
    ++g_blk->n_collectives;
  } else {
    while (nb.gen == g) yield_to_scheduler();
  }
  return mode == 1 ? nb.res_or[g & 1] : (mode == 2 ? nb.res_and[g & 1] : 0);
}
}  // namespace a1emu
inline void __syncthreads() {
  a1emu::Block& b = *a1emu::g_blk;
  const unsigned g = b.bgen;
  if (++b.bcount == b.nthreads) {
    b.bcount = 0;
    b.bgen = g + 1;
    ++b.progress;
  } else {
    while (b.bgen == g) a1emu::yield_to_scheduler();
  }
}

// mma.sync.aligned.m8n8k4.row.col.f64 (PTX ISA fragment layout): A: lane l holds A[l>>2][l&3]; B: lane l holds
// B[l&3][l>>2]; C/D: lane l holds [l>>2][2(l&3)], [l>>2][2(l&3)+1].
inline void a1emu_dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  a1emu::Warp& w = a1emu::my_warp();
  const unsigned g = a1emu::warp_arrive(a1emu::d2u(a), a1emu::d2u(b));
  ++a1emu::g_blk->n_mma;
  const int lane = a1emu::lane_id(), row = lane >> 2, col = 2 * (lane & 3);
  double r0 = c0, r1 = c1;
  // arithmetic model of the tensor unit (PTX leaves the accumulation order open): 0 = chain of fused multiply-adds,
  // 1 = products rounded, then added in k order, 2 = rounded products summed pairwise, then added to C
  static const int model = std::getenv("A1EMU_DMMA_MODEL") ? std::atoi(std::getenv("A1EMU_DMMA_MODEL")) : 0;
  double p0[4], p1[4];
  for (int k = 0; k < 4; ++k) {
    const double av = a1emu::u2d(w.slot[g & 1][row * 4 + k]);
    const double b0 = a1emu::u2d(w.slot2[g & 1][col * 4 + k]), b1 = a1emu::u2d(w.slot2[g & 1][(col + 1) * 4 + k]);
    if (a1emu::g_f32 > 0) {   // low-precision inputs, fp32 accumulation (what a tf32 / bf16 MMA does)
      r0 = (double)((float)r0 + (float)(a1emu::lowp(av) * a1emu::lowp(b0))); r1 = (double)((float)r1 + (float)(a1emu::lowp(av) * a1emu::lowp(b1)));
    }
    else if (model == 0) { r0 = std::fma(av, b0, r0); r1 = std::fma(av, b1, r1); }
    else { volatile double q0 = av * b0, q1 = av * b1; p0[k] = q0; p1[k] = q1; }
  }
  if (a1emu::g_f32 > 0) { d0 = r0; d1 = r1; return; }
  if (model == 1) { for (int k = 0; k < 4; ++k) { r0 += p0[k]; r1 += p1[k]; } }
  if (model == 2) { r0 += (p0[0] + p0[1]) + (p0[2] + p0[3]); r1 += (p1[0] + p1[1]) + (p1[2] + p1[3]); }
  // the operands of this collective stay valid until every lane has arrived at the NEXT one, so no second barrier
  d0 = r0;
  d1 = r1;
}

// ------------------------------------------------------------------------------------------------
// scalar intrinsics
// ------------------------------------------------------------------------------------------------
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::max;
using std::min;
using std::sqrt;
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __drcp_rn(double x) { return 1.0 / x; }
inline double __dsqrt_rn(double x) { return std::sqrt(x); }
// vector types used by the fused-collect store path of the device code (never executed by the emulator: npeer == 0)
struct float2 { float x, y; };
struct double2 { double x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int __double2hiint(double x) { return (int)(uint32_t)(a1emu::d2u(x) >> 32); }
inline int __double2loint(double x) { return (int)(uint32_t)(a1emu::d2u(x) & 0xffffffffull); }
inline long long __double_as_longlong(double x) { return (long long)a1emu::d2u(x); }
inline double __longlong_as_double(long long x) { return a1emu::u2d((uint64_t)x); }
inline double __hiloint2double(int hi, int lo) { return a1emu::u2d(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// a1mpc_misc.cuh -- non-template kernels; include from exactly one translation unit (a1mpc_api.cu)
#pragma once
#include "a1mpc_device.cuh"

namespace a1mpc {

// -------------------------------------------------------------------------------------------
// pack kernel: SoA batch -> per-class records.  Thread-per-QP, every load is a coalesced 64-bit
// batch-major access (32 consecutive QPs per warp instruction).
// -------------------------------------------------------------------------------------------
__global__ void pack_kernel(DevInputs in, int B, double* __restrict__ rec, int cap, int* __restrict__ count,
                            DevOutputs out, int horizon) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t mask = in.contact[b] & 15u;
  const int ns = __popc(mask);
  if (ns == 0) {  // every foot is pinned to zero by fz in [0,0] (ConvexMpc.cpp:233,238)
    for (int k = 0; k < 12; ++k) out.f_body[(size_t)k * out.ld + b] = 0.0;
    out.status[b] = A1MPC_STATUS_NO_CONTACT;
    if (out.iters) out.iters[b] = 0;
    if (out.u_full)
      for (int k = 0; k < 12 * horizon; ++k) out.u_full[(size_t)k * out.ld + b] = 0.0;
    return;
  }
  const int slot = atomicAdd(&count[ns], 1);
  double* r = rec + ((size_t)(ns - 1) * cap + slot) * REC_DOUBLES;
#pragma unroll
  for (int k = 0; k < 12; ++k) r[k] = in.x0[(size_t)k * in.ld + b];
#pragma unroll
  for (int k = 0; k < 9; ++k) r[12 + k] = in.rot[(size_t)k * in.ld + b];
#pragma unroll
  for (int k = 0; k < 12; ++k) r[21 + k] = in.foot[(size_t)k * in.ld + b];
#pragma unroll
  for (int k = 0; k < 9; ++k) r[33 + k] = in.ref[(size_t)k * in.ld + b];
  r[42] = __hiloint2double((int)mask, b);
  r[43] = 0.0;
}

// Classes whose factor does not fit in shared memory (N=20 with four stance feet in fp64):
// reported, never silently approximated.
__global__ void unsupported_kernel(const double* __restrict__ rec, const int* __restrict__ count, int cls, DevOutputs out, int horizon) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count[cls]) return;
  const int b = __double2loint(rec[(size_t)q * REC_DOUBLES + 42]);
  for (int k = 0; k < 12; ++k) out.f_body[(size_t)k * out.ld + b] = 0.0;
  out.status[b] = A1MPC_STATUS_NUMERICAL;
  if (out.iters) out.iters[b] = 0;
  if (out.u_full)
    for (int k = 0; k < 12 * horizon; ++k) out.u_full[(size_t)k * out.ld + b] = 0.0;
}

// fp64 FMA pipe peak probe: 8 independent dependent-free DFMA chains per thread
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, k = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, k); a1 = fma(a1, m, k); a2 = fma(a2, m, k); a3 = fma(a3, m, k);
    a4 = fma(a4, m, k); a5 = fma(a5, m, k); a6 = fma(a6, m, k); a7 = fma(a7, m, k);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void flush_kernel(double* buf, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}

}  // namespace a1mpc

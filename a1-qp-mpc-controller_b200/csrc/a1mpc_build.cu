// a1mpc_build.cu -- ConvexMpc member parity kernel (a1mpc_build_qp_batch), both horizons
#include "a1mpc_internal.h"

namespace a1mpc {

cudaError_t build_dense_launch(const DevParams& P, const DevInputs& in, int B, double* H, double* g, double* lb, double* ub, cudaStream_t st) {
  static bool attr_done_dev[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_done = attr_done_dev[dev & 63];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(build_dense_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_dense_smem<10>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(build_dense_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_dense_smem<20>());
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (P.N == 10) build_dense_kernel<10><<<B, 128, build_dense_smem<10>(), st>>>(P, in, B, H, g, lb, ub);
  else build_dense_kernel<20><<<B, 128, build_dense_smem<20>(), st>>>(P, in, B, H, g, lb, ub);
  return cudaGetLastError();
}

}  // namespace a1mpc

// a1mpc_nccl.cpp -- optional final collect over NVLink: ncclAllGather of the [12][B_local]
// force blocks.  NCCL is dlopen'ed so that the library has no link-time dependency on it; the data path
// of the solver never needs a collective (independent QPs).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/a1mpc.h"

extern "C" {
void* a1mpc_internal_stream(a1mpc_handle* h);
void* a1mpc_internal_gather_begin(a1mpc_handle* h);
void a1mpc_internal_gather_end(a1mpc_handle* h);
int a1mpc_internal_device(a1mpc_handle* h);
void** a1mpc_internal_nccl_slot(a1mpc_handle* h);
void a1mpc_internal_set_error(const char* msg);
}

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_getuid)(ncclUniqueId_t*);
typedef int (*fn_initrank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_destroy)(void*);
struct Nccl {
  void* lib = nullptr;
  fn_getuid get_uid = nullptr;
  fn_initrank init_rank = nullptr;
  fn_allgather all_gather = nullptr;
  fn_errstr err = nullptr;
  fn_destroy destroy = nullptr;
} g;
bool load() {
  if (g.lib) return true;
  const char* env = std::getenv("A1MPC_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    g.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g.lib) break;
  }
  if (!g.lib) return false;
  g.get_uid = (fn_getuid)dlsym(g.lib, "ncclGetUniqueId");
  g.init_rank = (fn_initrank)dlsym(g.lib, "ncclCommInitRank");
  g.all_gather = (fn_allgather)dlsym(g.lib, "ncclAllGather");
  g.err = (fn_errstr)dlsym(g.lib, "ncclGetErrorString");
  g.destroy = (fn_destroy)dlsym(g.lib, "ncclCommDestroy");
  return g.get_uid && g.init_rank && g.all_gather;
}
int nccl_fail(const char* what, int code) {
  std::string m = std::string(what) + ": " + (g.err ? g.err(code) : "nccl error");
  a1mpc_internal_set_error(m.c_str());
  return A1MPC_ENCCL;
}
}  // namespace

extern "C" {

int a1mpc_nccl_unique_id(void* unique_id128) {
  if (!unique_id128) return A1MPC_EINVAL;
  if (!load()) { a1mpc_internal_set_error("libnccl.so.2 not loadable (set A1MPC_NCCL_LIB)"); return A1MPC_ENCCL; }
  ncclUniqueId_t id;
  int rc = g.get_uid(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  std::memcpy(unique_id128, &id, 128);
  return A1MPC_OK;
}

int a1mpc_nccl_init(a1mpc_handle* h, int nranks, int rank, const void* unique_id128) {
  if (!h || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return A1MPC_EINVAL;
  if (!load()) { a1mpc_internal_set_error("libnccl.so.2 not loadable (set A1MPC_NCCL_LIB)"); return A1MPC_ENCCL; }
  cudaSetDevice(a1mpc_internal_device(h));
  ncclUniqueId_t id;
  std::memcpy(&id, unique_id128, 128);
  void* comm = nullptr;
  int rc = g.init_rank(&comm, nranks, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *a1mpc_internal_nccl_slot(h) = comm;
  return A1MPC_OK;
}

// a1mpc_destroy: release the communicator (peers blocked in their own destroy otherwise wait for it)
void a1mpc_internal_nccl_destroy(void* comm) {
  if (comm && g.destroy) g.destroy(comm);
}

int a1mpc_allgather_forces(a1mpc_handle* h, const double* f_local, double* f_all, int B_local) {
  if (!h || !f_local || !f_all || B_local <= 0) return A1MPC_EINVAL;
  void* comm = *a1mpc_internal_nccl_slot(h);
  if (!comm) { a1mpc_internal_set_error("a1mpc_nccl_init was not called"); return A1MPC_ENCCL; }
  cudaSetDevice(a1mpc_internal_device(h));
  // ncclFloat64 = 8
  // own stream, forked after the solve that produced f_local: the collect overlaps the next batch's kernels;
  // a1mpc_sync / a1mpc_event_record join it back
  void* gs = a1mpc_internal_gather_begin(h);
  if (!gs) { a1mpc_internal_set_error("could not create the collect stream"); return A1MPC_ECUDA; }
  int rc = g.all_gather(f_local, f_all, (size_t)12 * B_local, 8, comm, (cudaStream_t)gs);
  if (rc) return nccl_fail("ncclAllGather", rc);
  a1mpc_internal_gather_end(h);
  return A1MPC_OK;
}

}  // extern "C"

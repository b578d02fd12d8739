// Collective under the C ABI: a plan can own an NCCL communicator, and tdb200_loss_grad then
// enqueues the all-reduce of the [loss terms | gradient] vector on the caller's stream right after the reduction
// kernel - no host code between the last kernel and the collective, graph-capturable, and usable from a host that is
// not Python.  NCCL is bound lazily with dlopen / dlsym (the library torch has already loaded is reused; a process
// that never shards never touches NCCL), so libtedeous_b200.so has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>
#include <string>

#include "tdb200.h"

extern "C" void tdb200_set_error_(const char* msg);

namespace tdb {

namespace {
typedef struct { char internal[128]; } nccl_unique_id;      // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm*, int, nccl_unique_id, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef const char* (*fn_get_error_string)(int);

struct Nccl {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_get_error_string error_string = nullptr;
  bool tried = false;
};
Nccl g_nccl;

bool load_nccl() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);         // the copy the process already uses (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_nccl.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
  g_nccl.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy) return false;
  g_nccl.handle = h;
  return true;
}
int nccl_fail(int rc, const char* what) {
  std::string m = std::string(what) + ": NCCL error " + std::to_string(rc);
  if (g_nccl.error_string) m += std::string(" (") + g_nccl.error_string(rc) + ")";
  tdb200_set_error_(m.c_str());
  return TDB200_ERR_CUDA;
}
}  // namespace

int comm_create(const void* unique_id, int rank, int world, void** comm_out) {
  if (!load_nccl()) { tdb200_set_error_("NCCL (libnccl.so.2) could not be loaded"); return TDB200_ERR_NO_DEVICE; }
  nccl_unique_id id;
  memcpy(id.internal, unique_id, sizeof(id.internal));
  nccl_comm c = nullptr;
  const int rc = g_nccl.comm_init_rank(&c, world, id, rank);
  if (rc != 0) return nccl_fail(rc, "ncclCommInitRank");
  *comm_out = c;
  return TDB200_OK;
}
int comm_all_reduce_sum(void* comm, float* buf, size_t n, cudaStream_t s) {
  const int rc = g_nccl.all_reduce(buf, buf, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, s);
  return rc == 0 ? TDB200_OK : nccl_fail(rc, "ncclAllReduce");
}
void comm_destroy(void* comm) {
  if (comm && g_nccl.comm_destroy) g_nccl.comm_destroy(comm);
}

}  // namespace tdb

extern "C" int tdb200_comm_unique_id(void* id_out_128_bytes) {
  if (!id_out_128_bytes) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  if (!tdb::load_nccl()) { tdb200_set_error_("NCCL (libnccl.so.2) could not be loaded"); return TDB200_ERR_NO_DEVICE; }
  tdb::nccl_unique_id id;
  const int rc = tdb::g_nccl.get_unique_id(&id);
  if (rc != 0) return tdb::nccl_fail(rc, "ncclGetUniqueId");
  memcpy(id_out_128_bytes, id.internal, sizeof(id.internal));
  return TDB200_OK;
}

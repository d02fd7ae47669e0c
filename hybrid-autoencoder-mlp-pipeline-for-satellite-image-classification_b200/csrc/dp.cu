// Data-parallel gradient exchange: one NCCL communicator per process, one sum-allreduce of the flat
// fp32 gradient buffer per step (the reference itself is single-device, NB:277).
#include <nccl.h>

#include <cstring>

#include "common.cuh"

struct ae_dp_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

using namespace ae;

#define AE_NCCL(expr)                                                                             \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) {                                                                      \
      ae::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r));   \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

extern "C" {

int ae_dp_get_unique_id(uint8_t* id_host) {
  AE_CHECK(id_host != nullptr, "ae_dp_get_unique_id: null argument");
  static_assert(sizeof(ncclUniqueId) <= AE_DP_UNIQUE_ID_BYTES, "unique id does not fit");
  ncclUniqueId id;
  AE_NCCL(ncclGetUniqueId(&id));
  memset(id_host, 0, AE_DP_UNIQUE_ID_BYTES);
  memcpy(id_host, &id, sizeof(id));
  return 0;
}

int ae_dp_init(const uint8_t* id_host, int rank, int world, ae_dp_comm_t** out) {
  AE_CHECK(id_host && out && world >= 1 && rank >= 0 && rank < world, "ae_dp_init: bad argument");
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  ae_dp_comm* c = new ae_dp_comm();
  c->rank = rank; c->world = world;
  ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ae_dp_init: ncclCommInitRank failed: %s", ncclGetErrorString(r));
    delete c;
    return 1;
  }
  *out = c;
  return 0;
}

int ae_dp_allreduce(ae_dp_comm_t* c, float* buf, int64_t n, ae_stream_t stream) {
  AE_CHECK(c && c->comm && buf && n >= 0, "ae_dp_allreduce: bad argument");
  AE_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, c->comm, (cudaStream_t)stream));
  return 0;
}

int ae_dp_world(const ae_dp_comm_t* c) { return c ? c->world : 1; }

void ae_dp_destroy(ae_dp_comm_t* c) {
  if (!c) return;
  if (c->comm) {
    // finalize flushes outstanding work; a communicator that cannot be finalised cleanly is aborted rather than left to block exit
    if (ncclCommFinalize(c->comm) == ncclSuccess) ncclCommDestroy(c->comm);
    else ncclCommAbort(c->comm);
  }
  delete c;
}

}  // extern "C"

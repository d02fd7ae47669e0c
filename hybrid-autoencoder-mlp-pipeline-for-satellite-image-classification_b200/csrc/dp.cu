// Data-parallel gradient exchange (the reference itself is single-device, NB:277).  Two forms:
//   * NCCL: sum-allreduces of slices of the flat fp32 gradient buffer, issued by the engine on side branches of the step.
//   * fused over NVLink peer memory (ae_dp_peers_attach + k_dp_adam): ONE kernel per step that is reduce-scatter, Adam and
//     all-gather at once.  Every rank owns a 1/world shard of the flat buffer: it sums that shard of every rank's gradient
//     straight out of the peers' memory (fixed rank order: deterministic, every replica identical), applies Adam to its
//     shard (moments are only ever touched for the own shard) and stores the new parameters into every rank's parameter
//     buffer.  Two flag exchanges order it: "my gradients are complete" before the reads, "I have written your
//     parameters" before anyone's next forward.
#include <cuda.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

static constexpr int DP_MAX_WORLD = 8;

namespace ae {
struct DpAttachment {        // peer-mapped views of one model's flat buffers
  float* params[DP_MAX_WORLD];
  float* grads[DP_MAX_WORLD];
  unsigned int* flags[DP_MAX_WORLD];
  int64_t flat_len;
};
}  // namespace ae
using ae::DpAttachment;

struct ae_dp_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  int max_ctas = 0;       // > 0: bound on the SMs one collective may occupy (ncclConfig_t.maxCTAs), 0 = NCCL's default
  std::vector<DpAttachment> attached;
  struct Mapping { cudaIpcMemHandle_t h; void* base; };
  std::vector<Mapping> mappings;   // one cudaIpcOpenMemHandle per distinct peer allocation
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
  // AE_B200_NCCL_MAX_CTAS=n bounds the SMs one collective may occupy (and the step's one-CTA-per-SM kernels then leave n SMs
  // free).  Measured on 2 B200s (profiles/r2_dp_notes.txt): the exchange is on the step's critical path, so every bound
  // below NCCL's own choice made the step slower (2 CTAs: 1.14 ms, 4: 0.87, 8: 0.79, 16: 0.76, unbounded: 0.74) -- default off.
  const char* mc = getenv("AE_B200_NCCL_MAX_CTAS");
  if (mc && atoi(mc) >= 1 && atoi(mc) <= 64) c->max_ctas = atoi(mc);
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  if (c->max_ctas > 0) cfg.maxCTAs = c->max_ctas;
  ncclResult_t r = ncclCommInitRankConfig(&c->comm, world, id, rank, &cfg);
  if (r != ncclSuccess) {
    set_error("ae_dp_init: ncclCommInitRankConfig failed: %s", ncclGetErrorString(r));
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

// -------------------------------------------------------------------------------------------------------------
// fused reduce-scatter + Adam + all-gather over peer memory
// -------------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace ae {

// flags of one rank (uint32, in that rank's memory; peers write their own slot): [0,8) ready[r], [8,16) done[r],
// [16] step sequence number (local), [17] CTA counter (local)
enum { DPF_READY = 0, DPF_DONE = 8, DPF_SEQ = 16, DPF_COUNT = 17, DPF_WORDS = 32 };

struct DpAdam {
  float* params[DP_MAX_WORLD];
  const float* grads[DP_MAX_WORLD];
  unsigned int* flags[DP_MAX_WORLD];
  float *m, *v;
  int64_t lo4, hi4;          // own shard in float4 units
  int rank, world;
  float lr, b1, b2, eps, wd, gscale;
  int* step;
  int bump;                  // advance Adam's step counter (the last launch of a step)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {       // peer memory changes every step: never the read-only path
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// bounded: a rank that never arrives must trap the others, not hang the box
__device__ __forceinline__ void wait_flag(const unsigned int* p, unsigned int t) {
  for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
    if ((int)(ld_acquire_sys(p) - t) >= 0) return;
    __nanosleep(100);
  }
  printf("ae_b200: data-parallel peer flag timed out (block %d thread %d, waiting for step %u)\n", (int)blockIdx.x, (int)threadIdx.x, t);
  __trap();
}

__global__ void __launch_bounds__(512) k_dp_adam(const __grid_constant__ DpAdam a) {
  unsigned int* mine = a.flags[a.rank];
  const unsigned int t = mine[DPF_SEQ] + 1u;                 // advanced by the last CTA of this launch, after everybody read it
  const int tid = threadIdx.x;
  // ---- my gradients are complete (this kernel is stream-ordered behind the backward pass): tell every rank, then wait for all
  if (blockIdx.x == 0 && tid < a.world) st_release_sys(a.flags[tid] + DPF_READY + a.rank, t);
  if (tid < a.world) wait_flag(mine + DPF_READY + tid, t);
  __syncthreads();
  // ---- own shard: sum over ranks (fixed order), Adam (same arithmetic as k_adam), parameters to every rank
  const int st_ = a.step[0] + 1;
  const double bc1 = 1.0 - pow((double)a.b1, (double)st_);
  const double bc2 = 1.0 - pow((double)a.b2, (double)st_);
  const float step_size = (float)((double)a.lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float omb1 = 1.f - a.b1, omb2 = 1.f - a.b2;
  for (int64_t i = a.lo4 + (int64_t)blockIdx.x * blockDim.x + tid; i < a.hi4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
      if (r < a.world) {
        const float4 gv = ld_peer_f4(a.grads[r] + 4 * i);
        gs.x += gv.x; gs.y += gv.y; gs.z += gv.z; gs.w += gv.w;
      }
    }
    float4 pv = reinterpret_cast<float4*>(a.params[a.rank])[i];
    float4 mv = reinterpret_cast<float4*>(a.m)[i];
    float4 vv = reinterpret_cast<float4*>(a.v)[i];
    float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gs.x, gs.y, gs.z, gs.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = ge[k] * a.gscale;
      if (a.wd != 0.f) gg = fmaf(a.wd, pe[k], gg);
      me[k] = me[k] + (gg - me[k]) * omb1;
      ve[k] = ve[k] * a.b2 + (gg * gg) * omb2;
      const float denom = sqrtf(ve[k]) / bc2_sqrt + a.eps;
      pe[k] = pe[k] - step_size * (me[k] / denom);
    }
    const float4 pn = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
      if (r < a.world) reinterpret_cast<float4*>(a.params[r])[i] = pn;
  }
  // ---- all of this rank's parameter stores are out: the last CTA tells every rank and waits for theirs
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(mine + DPF_COUNT, 1u);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      mine[DPF_COUNT] = 0u;
      mine[DPF_SEQ] = t;
      if (a.bump) a.step[0] = st_;
      for (int r = 0; r < a.world; ++r) st_release_sys(a.flags[r] + DPF_DONE + a.rank, t);
      for (int r = 0; r < a.world; ++r) wait_flag(mine + DPF_DONE + r, t);
      __threadfence_system();
    }
  }
}

// the attachment of `c` that describes these flat buffers, or NULL
const DpAttachment* dp_find_attachment(const ae_dp_comm* c, const float* flat_params, const float* flat_grads, int64_t flat_len) {
  if (!c) return nullptr;
  for (const DpAttachment& a : c->attached)
    if (a.params[c->rank] == flat_params && a.grads[c->rank] == flat_grads && a.flat_len == flat_len) return &a;
  return nullptr;
}

// elements [lo, lo + n) of the flat buffers (multiples of 4); every launch is a complete exchange round of its own
int dp_adam_fused(const ae_dp_comm* c, const DpAttachment* at, float* m, float* v, int64_t lo, int64_t n, float lr, float b1, float b2,
                  float eps, float wd, int* step_dev, int bump, cudaStream_t st) {
  AE_CHECK(c && at && lo % 4 == 0 && n % 4 == 0 && lo >= 0 && lo + n <= at->flat_len, "dp_adam_fused: bad argument");
  DpAdam a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < c->world; ++r) { a.params[r] = at->params[r]; a.grads[r] = at->grads[r]; a.flags[r] = at->flags[r]; }
  a.m = m; a.v = v; a.rank = c->rank; a.world = c->world;
  const int64_t n4 = n / 4, lo4 = lo / 4;
  const int64_t per = (n4 + c->world - 1) / c->world;
  const int64_t b4 = per * c->rank < n4 ? per * c->rank : n4;
  a.lo4 = lo4 + b4;
  a.hi4 = lo4 + (b4 + per < n4 ? b4 + per : n4);
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd; a.gscale = 1.f / (float)c->world; a.step = step_dev; a.bump = bump;
  // enough CTAs to keep the remote loads in flight, few enough that all of them are resident at once (they wait for peers)
  const int ctas = per >= 64 * 512 ? 64 : per >= 8 * 512 ? 8 : 1;
  k_dp_adam<<<ctas, 512, 0, st>>>(a);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

extern "C" {

int ae_dp_ipc_export(const void* dev_ptr, uint8_t* handle, int64_t* offset) {
  AE_CHECK(dev_ptr && handle && offset, "ae_dp_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= AE_DP_IPC_HANDLE_BYTES, "IPC handle does not fit");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  AE_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && fp,
           "ae_dp_ipc_export: cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  AE_CHECK(((RangeFn)fp)(&base, &size, (CUdeviceptr)dev_ptr) == CUDA_SUCCESS, "ae_dp_ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  AE_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  memset(handle, 0, AE_DP_IPC_HANDLE_BYTES);
  memcpy(handle, &h, sizeof(h));
  *offset = (int64_t)((CUdeviceptr)dev_ptr - base);
  return 0;
}

int ae_dp_peers_attach(ae_dp_comm_t* c, const uint8_t* handles, const int64_t* offsets, float* own_params, float* own_grads,
                       void* own_flags, int64_t flat_len) {
  AE_CHECK(c && handles && offsets && own_params && own_grads && own_flags, "ae_dp_peers_attach: null argument");
  AE_CHECK(c->world <= DP_MAX_WORLD, "ae_dp_peers_attach: at most %d ranks", DP_MAX_WORLD);
  AE_CHECK(flat_len % 4 == 0, "ae_dp_peers_attach: flat length must be a multiple of 4");
  DpAttachment at;
  memset(&at, 0, sizeof(at));
  at.flat_len = flat_len;
  for (int r = 0; r < c->world; ++r) {
    void* ptrs[3];
    for (int b = 0; b < 3; ++b) {
      if (r == c->rank) { ptrs[b] = b == 0 ? (void*)own_params : b == 1 ? (void*)own_grads : own_flags; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + ((size_t)r * 3 + b) * AE_DP_IPC_HANDLE_BYTES, sizeof(h));
      void* base = nullptr;
      for (const auto& mp : c->mappings)
        if (memcmp(&mp.h, &h, sizeof(h)) == 0) base = mp.base;
      if (!base) {
        AE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->mappings.push_back({h, base});
      }
      ptrs[b] = (char*)base + offsets[(size_t)r * 3 + b];
    }
    at.params[r] = (float*)ptrs[0]; at.grads[r] = (float*)ptrs[1]; at.flags[r] = (unsigned int*)ptrs[2];
  }
  c->attached.push_back(at);
  return 0;
}

int ae_dp_world(const ae_dp_comm_t* c) { return c ? c->world : 1; }

int ae_dp_max_ctas(const ae_dp_comm_t* c) { return c ? c->max_ctas : 0; }

void ae_dp_destroy(ae_dp_comm_t* c) {
  if (!c) return;
  for (auto& mp : c->mappings) cudaIpcCloseMemHandle(mp.base);
  if (c->comm) {
    // finalize flushes outstanding work; a communicator that cannot be finalised cleanly is aborted rather than left to block exit
    if (ncclCommFinalize(c->comm) == ncclSuccess) ncclCommDestroy(c->comm);
    else ncclCommAbort(c->comm);
  }
  delete c;
}

}  // extern "C"

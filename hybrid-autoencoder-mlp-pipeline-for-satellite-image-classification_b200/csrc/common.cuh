// Shared internal declarations of libae_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ae_b200.h"

namespace ae {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define AE_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ae::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define AE_CHECK(cond, ...)            \
  do {                                 \
    if (!(cond)) {                     \
      ae::set_error(__VA_ARGS__);      \
      return 1;                        \
    }                                  \
  } while (0)

#define AE_TRY(expr)        \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

#define AE_LAUNCH_CHECK()                                                                   \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ae::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

// ---- device-side descriptors ------------------------------------------------------------------
struct Operand {
  const float* src;
  const float* src2;
  const float* bnc;  // [AE_BNC_ROWS][C]
  float scalar;
  int mode;          // AE_OP_*
  int C;             // channel count the coefficient block is indexed by (channel = k % C)
};

struct Epilogue {
  int mode;          // AE_EPI_*
  const float* bias;
  const float* y;
  const float* bnc;
  double* stats;
  int C;             // channels of the BN this epilogue feeds (channel = n % C)
  int nsplit;        // BNRELU_SPLIT / planes: bf16 planes written (2 = hi + lo, 1 = hi only)
  void* planes;      // dense row GEMM, STORE mode: also (out == NULL: only) write the result as split-bf16 planes [nsplit][M*N]
};

// BatchNorm coefficient job a consumer kernel can run in its prologue instead of a separate launch:
//   BN_JOB_FINALIZE: ae_bn_finalize (statistics -> scale/shift/mean/rstd, running-stat update)
//   BN_JOB_BWD     : ae_bn_bwd_reduce (sums -> A/B/C, dgamma, dbeta)
enum { BN_JOB_NONE = 0, BN_JOB_FINALIZE = 1, BN_JOB_BWD = 2 };
struct BnJob {
  int kind;
  const double* stats;
  double count;
  const float* gamma;
  const float* beta;
  float* rmean;
  float* rvar;
  float* bnc;
  float* dgamma;
  float* dbeta;
  int C;
  int training;
  int64_t* nbt;      // FINALIZE, training: this layer's num_batches_tracked, incremented by the job (or NULL)
  float* dzero;      // BWD: C floats set to zero by the job -- the gradient of the bias that feeds this BatchNorm (or NULL)
};

struct Geom {
  int B, Hs, Ws, Cb, Cs;
  int lHs, lWs;      // log2
};

inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

inline Operand make_operand(const ae_operand_t* o, int C) {
  Operand r;
  r.src = o->src; r.src2 = o->src2; r.bnc = o->bnc; r.scalar = o->scalar; r.mode = o->mode; r.C = C;
  return r;
}
inline Operand split_operand(const void* planes, int C) {
  Operand r; r.src = static_cast<const float*>(planes); r.src2 = nullptr; r.bnc = nullptr; r.scalar = 0.f;
  r.mode = AE_OP_SPLIT_BF16; r.C = C;
  return r;
}
inline Operand raw_operand(const float* p) {
  Operand r; r.src = p; r.src2 = nullptr; r.bnc = nullptr; r.scalar = 0.f; r.mode = AE_OP_RAW; r.C = 1;
  return r;
}
inline Operand bnrelu_operand(const float* y, const float* bnc, int C) {
  Operand r; r.src = y; r.src2 = nullptr; r.bnc = bnc; r.scalar = 0.f; r.mode = AE_OP_BNRELU; r.C = C;
  return r;
}
inline Operand bnbwd_operand(const float* dz, const float* y, const float* bnc, int C) {
  Operand r; r.src = dz; r.src2 = y; r.bnc = bnc; r.scalar = 0.f; r.mode = AE_OP_BNBWD; r.C = C;
  return r;
}
inline Epilogue make_epilogue(const ae_epilogue_t* e, int C) {
  Epilogue r;
  if (!e) { r.mode = AE_EPI_STORE; r.bias = nullptr; r.y = nullptr; r.bnc = nullptr; r.stats = nullptr; r.C = C; r.nsplit = 2; r.planes = nullptr; return r; }
  r.mode = e->mode; r.bias = e->bias; r.y = e->y; r.bnc = e->bnc; r.stats = e->stats; r.C = C; r.nsplit = 2; r.planes = nullptr;
  return r;
}
inline Epilogue store_epilogue(const float* bias = nullptr) {
  Epilogue r; r.mode = AE_EPI_STORE; r.bias = bias; r.y = nullptr; r.bnc = nullptr; r.stats = nullptr; r.C = 1; r.nsplit = 2; r.planes = nullptr;
  return r;
}
inline Epilogue bias_stats_epilogue(const float* bias, double* stats, int C) {
  Epilogue r; r.mode = AE_EPI_BIAS_STATS; r.bias = bias; r.y = nullptr; r.bnc = nullptr; r.stats = stats; r.C = C; r.nsplit = 2; r.planes = nullptr;
  return r;
}
inline Epilogue bnrelu_split_epilogue(const float* bias, const float* bnc, int C, int nsplit) {
  Epilogue r; r.mode = AE_EPI_BNRELU_SPLIT; r.bias = bias; r.y = nullptr; r.bnc = bnc; r.stats = nullptr; r.C = C; r.nsplit = nsplit; r.planes = nullptr;
  return r;
}
inline Epilogue relubwd_epilogue(const float* y, const float* bnc, double* stats, int C) {
  Epilogue r; r.mode = AE_EPI_RELUBWD_STATS; r.bias = nullptr; r.y = y; r.bnc = bnc; r.stats = stats; r.C = C; r.nsplit = 2; r.planes = nullptr;
  return r;
}

#ifdef __CUDACC__
// Load 4 consecutive channels [c, c+4) of an activation operand at element offset `off`
// (off and c are multiples of 4) and apply the operand transform.  Invalid (padding) -> 0.
__device__ __forceinline__ float4 load_operand4(const Operand& op, size_t off, int c, bool valid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!valid) return v;
  v = __ldg(reinterpret_cast<const float4*>(op.src + off));
  if (op.mode == AE_OP_RAW) return v;
  if (op.mode == AE_OP_BNRELU) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_SCALE * op.C + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_SHIFT * op.C + c));
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    return v;
  }
  // AE_OP_BNBWD
  const float4 y = __ldg(reinterpret_cast<const float4*>(op.src2 + off));
  // dy = A*dz + B*(y - mean) + C.  (y - mean) is formed first: folding B*mean into C would turn the per-channel
  // rounding of C into an offset that coherent sums (next layer's statistics, weight gradients) amplify.
  const float4 a = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_A * op.C + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_B * op.C + c));
  const float4 k = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_C * op.C + c));
  const float4 m = __ldg(reinterpret_cast<const float4*>(op.bnc + AE_BNC_MEAN * op.C + c));
  v.x = fmaf(a.x, v.x, fmaf(b.x, y.x - m.x, k.x));
  v.y = fmaf(a.y, v.y, fmaf(b.y, y.y - m.y, k.y));
  v.z = fmaf(a.z, v.z, fmaf(b.z, y.z - m.z, k.z));
  v.w = fmaf(a.w, v.w, fmaf(b.w, y.w - m.w, k.w));
  return v;
}

// One BatchNorm coefficient job (what k_bn_finalize / k_bn_bwd_reduce do as launches), run by `nthreads` threads of ONE
// CTA after the statistics are complete.  The statistics were accumulated by other SMs: they are read past L1.
__device__ __forceinline__ void bn_job_run(const BnJob& job, int tid, int nthreads) {
  const int C = job.C;
  for (int c = tid; c < C; c += nthreads) {
    if (job.kind == BN_JOB_FINALIZE) {
      double mean, var;
      if (job.training) {
        mean = __ldcg(job.stats + c) / job.count;
        var = __ldcg(job.stats + C + c) / job.count - mean * mean;
        if (var < 0.0) var = 0.0;
      } else {
        mean = (double)job.rmean[c];
        var = (double)job.rvar[c];
      }
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      const float scale = job.gamma[c] * rstd;
      const float shift = job.beta[c] - (float)mean * scale;
      if (c == 0 && job.training && job.nbt) *job.nbt += 1;
      if (job.training && job.rmean) {
        const double unb = job.count > 1.0 ? var * job.count / (job.count - 1.0) : var;
        job.rmean[c] = (float)(0.9 * (double)job.rmean[c] + 0.1 * mean);
        job.rvar[c] = (float)(0.9 * (double)job.rvar[c] + 0.1 * unb);
      }
      job.bnc[AE_BNC_SCALE * C + c] = scale; job.bnc[AE_BNC_SHIFT * C + c] = shift;
      job.bnc[AE_BNC_MEAN * C + c] = (float)mean; job.bnc[AE_BNC_RSTD * C + c] = rstd;
    } else if (job.kind == BN_JOB_BWD) {
      const double s1 = __ldcg(job.stats + c), s2 = __ldcg(job.stats + C + c);
      const double rstd = (double)job.bnc[AE_BNC_RSTD * C + c];
      const double a = (double)job.gamma[c] * rstd;
      job.bnc[AE_BNC_A * C + c] = (float)a;
      job.bnc[AE_BNC_B * C + c] = (float)(-a * rstd * s2 / job.count);
      job.bnc[AE_BNC_C * C + c] = (float)(-a * s1 / job.count);
      if (job.dgamma) job.dgamma[c] = (float)s2;
      if (job.dbeta) job.dbeta[c] = (float)s1;
      if (job.dzero) job.dzero[c] = 0.f;
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

// ---- internal launchers (implemented across the .cu files) ---------------------------------------
enum { FAM_DENSE = 0, FAM_FPROP = 1, FAM_DGRAD = 2 };

// C[m][n] = sum_k A(m,k) * Bp[k][n]; Bp row-major fp32 [K][N] (for FAM_DGRAD: 4 phase blocks stacked, 9*Cs rows)
struct RowGemm {
  int family;
  Geom g;
  int M, N, K;       // FAM_DGRAD: M = B*Hs*Ws rows per phase, K is per phase (derived), N = Cb
  Operand A;
  const float* Bp;
  Epilogue epi;
  float* out;
  int splitK;        // >1: partial[s][M][N] written instead of out (no bias / stats), DENSE/FPROP only
  float* partial;
  // tcgen05 path: BatchNorm coefficient job of the statistics this launch accumulates, run by the LAST CTA to finish
  // (a self-resetting counter in the caller's memory) instead of a k_bn_finalize / k_bn_bwd_reduce launch behind it
  const BnJob* tail_job;
  unsigned int* tail_counter;
};
int simt_rowgemm(const RowGemm& p, cudaStream_t st);

// C[i][j] = sum_m A(m,i) * B(m,j), m in [0,M): reduction over rows.
struct ColGemm {
  int gather;        // 0: A dense [M][I]; 1: A(m, i=(tap,cb)) gathered from the big image around small pixel m
  Geom g;
  int M, I, J;
  Operand A, B;
  float* out;        // final layout
  int permC, permHW; // i' = (i % permC) * permHW + i / permC when permC > 0
  int transposed;    // 0: out[i'*J + j]; 1: out[j*I + i']
  int splitK;
  float* partial;    // [splitK][I*J] in final layout order
};
int simt_colgemm(const ColGemm& p, cudaStream_t st);
int colgemm_default_split(int M, int I, int J);

// out[idx] = sum_s partial[s*n + idx] (+ bias[idx % bias_n]) (+ addend[idx])
int reduce_partials(const float* partial, int splits, int64_t n, const float* bias, int bias_n,
                    const float* addend, float* out, cudaStream_t st);
// db[perm(n)] = sum_m a[m][n]
int column_sums(const float* a, int M, int N, int permC, int permHW, float* out, cudaStream_t st);

// fp32 [K][N] packs for the SIMT path
int pack_conv_simt(const float* w, int Cs, int Cb, float* fwd, float* dgrad, cudaStream_t st);
// generic permuted transpose: dst[r][c] from torch-layout linear weight (see dense.cu)
int pack_linear(const float* w, int N, int K, int permC, int permHW, int kind, float* dst, cudaStream_t st);
int permute_vector(const float* src, int n, int permC, int permHW, float* dst, cudaStream_t st);

// fused data-parallel exchange over peer memory (dp.cu)
struct DpAttachment;
}  // namespace ae
struct ae_dp_comm;
namespace ae {
const DpAttachment* dp_find_attachment(const ae_dp_comm* c, const float* flat_params, const float* flat_grads, int64_t flat_len);
int dp_adam_fused(const ae_dp_comm* c, const DpAttachment* at, float* m, float* v, int64_t lo, int64_t n, float lr, float b1, float b2,
                  float eps, float wd, int* step_dev, int bump, cudaStream_t st);

// tcgen05 + TMA path (tma_gemm.cu)
size_t tma_packed_bytes(int Cs, int Cb, int nsplit);
int tma_pack_conv(const float* w, int Cs, int Cb, int nsplit, void* fwd, void* dgrad, cudaStream_t st);
int tma_split_operand(const Operand& op, int64_t count, void* planes, int nsplit, const BnJob* job, cudaStream_t st);
int run_bn_job(const BnJob& job, cudaStream_t st);   // the same job as stand-alone launches
int tma_rowgemm(const RowGemm& p, const void* packed, int nsplit, cudaStream_t st);   // p.A: AE_OP_SPLIT_BF16
bool tma_rowgemm_supported(const RowGemm& p);
// second-generation kernel for the training epilogues (rowgemm2.cu); tma_rowgemm dispatches to it
bool rowgemm2_supported(const RowGemm& p);   // shape + epilogue
bool rowgemm2_preferred(const RowGemm& p);   // ... and measured to be the faster generation for this problem size
int tma_rowgemm2(const RowGemm& p, const void* packed, int nsplit, cudaStream_t st);
void rowgemm2_set_sm_reserve(int sms);       // SMs its one-CTA-per-SM grids leave to concurrently running collectives
// dense_tc.cu: Linear over split-bf16 activation planes on tcgen05 (eval-mode encoder bottleneck)
bool dense_tc_supported(int N, int K);
size_t dense_tc_pack_bytes(int N, int K, int nsplit);
int dense_tc_pack(const float* w, int N, int K, int permC, int permHW, int nsplit, void* pack, cudaStream_t st);
int dense_tc_split(int M, int K);
int dense_tc(const void* a_planes, const void* pack, int M, int N, int K, int nsplit, float* out_partial, size_t partial_bytes,
             int* ksplit_out, cudaStream_t st);
bool tma_wgrad_supported(const Geom& g);
int tma_wgrad_slices(const Geom& g);
size_t tma_wgrad_partial_bytes(const Geom& g);
int tma_wgrad(const Geom& g, const void* big_planes, const void* small_planes, float* dw, float* partial,
              size_t partial_bytes, int nsplit, cudaStream_t st);

}  // namespace ae

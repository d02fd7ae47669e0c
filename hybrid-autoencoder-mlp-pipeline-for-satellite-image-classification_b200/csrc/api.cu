// C-ABI entry points for the individual operators (include/ae_b200.h) + error state.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace ae {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int thin_gather_fwd(const Operand& thin, const float* w, const Epilogue& epi, float* out, int batch, cudaStream_t st);
int thin_scatter_sigmoid_fwd(const Operand& wide, const float* w, const float* bias, float* x_hat, const float* x,
                             double* sse, int batch, cudaStream_t st);
int thin_wgrad(const Operand& wide, const Operand& thin, float* dw, float* dbias, void* partials, size_t bytes,
               int batch, cudaStream_t st);
size_t thin_wgrad_workspace_bytes(int batch);
int thin_bwd_fused(const Operand& wide, const Operand& thin, const float* w, const Epilogue& epi, float* out_wide, float* dw,
                   float* dbias, void* partials, size_t bytes, int batch, cudaStream_t st, int phase);
int bn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta, float* rmean, float* rvar,
                float* bnc, int C, int training, cudaStream_t st);
int bn_bwd_reduce(const double* stats, int64_t count, const float* gamma, float* bnc, float* dgamma, float* dbeta, int C,
                  cudaStream_t st);
int softmax_ce(const float* logits, const int64_t* labels, int B, int C, float gscale, float* loss, float* dlogits,
               int* correct, const double* sse, double numel, float alpha, cudaStream_t st);
int sigmoid_mse(const float* x_hat, const float* x, int64_t n, float scale, float* loss, float* d_pre, double* sse_tmp,
                cudaStream_t st);
int adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                   float wd, float gscale, int* step_dev, cudaStream_t st);
int layout_convert(const void* src, void* dst, int N, int C, int H, int W, bool bf16, bool to_nhwc, cudaStream_t st);

static int make_geom(const ae_conv_geom_t* g, Geom* out) {
  AE_CHECK(g != nullptr, "conv geometry is null");
  AE_CHECK(g->batch >= 1, "conv geometry: batch must be >= 1");
  AE_CHECK(is_pow2(g->hs) && is_pow2(g->ws), "conv geometry: hs=%d, ws=%d must be powers of two", g->hs, g->ws);
  AE_CHECK(g->cb % 8 == 0 && g->cs % 16 == 0 && g->cb >= 8 && g->cs >= 16,
           "conv geometry: cb=%d must be a multiple of 8 and cs=%d a multiple of 16", g->cb, g->cs);
  out->B = g->batch; out->Hs = g->hs; out->Ws = g->ws; out->Cb = g->cb; out->Cs = g->cs;
  out->lHs = ilog2(g->hs); out->lWs = ilog2(g->ws);
  return 0;
}

static int nsplit_of(int precision) { return precision == AE_PREC_FP32 ? 2 : 1; }

}  // namespace ae

using namespace ae;

extern "C" {

const char* ae_last_error(void) { return get_error(); }
int ae_abi_version(void) { return AE_ABI_VERSION; }

int ae_device_supported(int ordinal) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ordinal) != cudaSuccess) { cudaGetLastError(); return 0; }
  return prop.major == 10 ? 1 : 0;
}

size_t ae_packed_weight_bytes(int cs, int cb, int precision, int backend) {
  if (backend == AE_BACKEND_SIMT) return (size_t)9 * cs * cb * sizeof(float);
  return tma_packed_bytes(cs, cb, nsplit_of(precision));
}

size_t ae_split_operand_bytes(int64_t count, int precision) { return (size_t)count * 2 * nsplit_of(precision); }

int ae_split_operand(const ae_operand_t* op, int channels, int64_t count, void* planes, int precision, ae_stream_t stream) {
  AE_CHECK(op && op->src && planes && channels >= 8 && count >= 8, "ae_split_operand: bad argument");
  return tma_split_operand(make_operand(op, channels), count, planes, nsplit_of(precision), nullptr, (cudaStream_t)stream);
}

int ae_pack_conv_weight(const float* w, int cs, int cb, void* packed_fwd, void* packed_dgrad, int precision, int backend,
                        ae_stream_t stream) {
  AE_CHECK(w != nullptr, "ae_pack_conv_weight: null weight");
  if (backend == AE_BACKEND_SIMT) return pack_conv_simt(w, cs, cb, (float*)packed_fwd, (float*)packed_dgrad, (cudaStream_t)stream);
  return tma_pack_conv(w, cs, cb, nsplit_of(precision), packed_fwd, packed_dgrad, (cudaStream_t)stream);
}

int ae_conv2d_s2_fwd(const ae_conv_geom_t* g, const ae_operand_t* big, const void* packed_fwd, const ae_epilogue_t* epi,
                     float* out_small, int precision, int backend, ae_stream_t stream) {
  RowGemm r{};
  AE_TRY(make_geom(g, &r.g));
  AE_CHECK(big && packed_fwd && out_small, "ae_conv2d_s2_fwd: null argument");
  r.family = FAM_FPROP; r.M = r.g.B * r.g.Hs * r.g.Ws; r.N = r.g.Cs; r.K = 9 * r.g.Cb;
  r.A = make_operand(big, r.g.Cb); r.Bp = (const float*)packed_fwd; r.epi = make_epilogue(epi, r.g.Cs);
  r.out = out_small; r.splitK = 1;
  if (backend == AE_BACKEND_SIMT) return simt_rowgemm(r, (cudaStream_t)stream);
  AE_CHECK(tma_rowgemm_supported(r), "ae_conv2d_s2_fwd: shape not supported by the tcgen05 path");
  return tma_rowgemm(r, packed_fwd, nsplit_of(precision), (cudaStream_t)stream);
}

int ae_conv2d_s2_dgrad(const ae_conv_geom_t* g, const ae_operand_t* small, const void* packed_dgrad,
                       const ae_epilogue_t* epi, float* out_big, int precision, int backend, ae_stream_t stream) {
  RowGemm r{};
  AE_TRY(make_geom(g, &r.g));
  AE_CHECK(small && packed_dgrad && out_big, "ae_conv2d_s2_dgrad: null argument");
  r.family = FAM_DGRAD; r.M = r.g.B * r.g.Hs * r.g.Ws; r.N = r.g.Cb; r.K = 0;
  r.A = make_operand(small, r.g.Cs); r.Bp = (const float*)packed_dgrad; r.epi = make_epilogue(epi, r.g.Cb);
  r.out = out_big; r.splitK = 1;
  if (backend == AE_BACKEND_SIMT) return simt_rowgemm(r, (cudaStream_t)stream);
  AE_CHECK(tma_rowgemm_supported(r), "ae_conv2d_s2_dgrad: shape not supported by the tcgen05 path");
  return tma_rowgemm(r, packed_dgrad, nsplit_of(precision), (cudaStream_t)stream);
}

size_t ae_conv2d_s2_wgrad_workspace_bytes(const ae_conv_geom_t* g, int precision, int backend) {
  (void)precision;
  if (!g) return 0;
  if (backend == AE_BACKEND_TC) {
    Geom gg;
    if (make_geom(g, &gg) != 0) return 0;
    return tma_wgrad_partial_bytes(gg);
  }
  const int M = g->batch * g->hs * g->ws, I = 9 * g->cb, J = g->cs;
  return (size_t)colgemm_default_split(M, I, J) * I * J * sizeof(float);
}

int ae_conv2d_s2_wgrad(const ae_conv_geom_t* g, const ae_operand_t* big, const ae_operand_t* small, float* dw,
                       void* partials, size_t partials_bytes, int precision, int backend, ae_stream_t stream) {
  ColGemm c{};
  AE_TRY(make_geom(g, &c.g));
  AE_CHECK(big && small && dw, "ae_conv2d_s2_wgrad: null argument");
  if (backend == AE_BACKEND_TC) {
    AE_CHECK(big->mode == AE_OP_SPLIT_BF16 && small->mode == AE_OP_SPLIT_BF16,
             "ae_conv2d_s2_wgrad: the tcgen05 path takes split-bf16 operands (ae_split_operand)");
    return tma_wgrad(c.g, big->src, small->src, dw, (float*)partials, partials_bytes, nsplit_of(precision), (cudaStream_t)stream);
  }
  c.gather = 1; c.M = c.g.B * c.g.Hs * c.g.Ws; c.I = 9 * c.g.Cb; c.J = c.g.Cs;
  c.A = make_operand(big, c.g.Cb); c.B = make_operand(small, c.g.Cs);
  c.out = dw; c.permC = c.g.Cb; c.permHW = 9; c.transposed = 1;
  c.splitK = colgemm_default_split(c.M, c.I, c.J);
  c.partial = (float*)partials;
  AE_CHECK(c.splitK == 1 || (partials && partials_bytes >= (size_t)c.splitK * c.I * c.J * 4),
           "ae_conv2d_s2_wgrad: partial buffer too small (%zu bytes)", partials_bytes);
  return simt_colgemm(c, (cudaStream_t)stream);
}

int ae_thin_gather_fwd(const ae_operand_t* thin, const float* w, const ae_epilogue_t* epi, float* out_wide, int batch,
                       int precision, int backend, ae_stream_t stream) {
  AE_CHECK(thin && w && out_wide && batch >= 1, "ae_thin_gather_fwd: bad argument");
  (void)backend;                                      // both backends run the fp32 CUDA-core kernels (see ae_b200.h)
  Epilogue ep = make_epilogue(epi, 32);
  ep.nsplit = nsplit_of(precision);                   // AE_EPI_BNRELU_SPLIT: planes written
  return thin_gather_fwd(make_operand(thin, 3), w, ep, out_wide, batch, (cudaStream_t)stream);
}

int ae_thin_scatter_sigmoid_fwd(const ae_operand_t* wide, const float* w, const float* bias, float* x_hat, const float* x,
                                double* sse, int batch, int precision, int backend, ae_stream_t stream) {
  AE_CHECK(wide && w && bias && x_hat && batch >= 1, "ae_thin_scatter_sigmoid_fwd: bad argument");
  (void)backend; (void)precision;
  return thin_scatter_sigmoid_fwd(make_operand(wide, 32), w, bias, x_hat, x, sse, batch, (cudaStream_t)stream);
}

size_t ae_thin_wgrad_workspace_bytes(int batch) {
  return thin_wgrad_workspace_bytes(batch);
}

int ae_thin_wgrad(const ae_operand_t* wide, const ae_operand_t* thin, float* dw, float* dbias_thin, void* partials,
                  size_t partials_bytes, int batch, int precision, int backend, ae_stream_t stream) {
  AE_CHECK(wide && thin && dw && partials && batch >= 1, "ae_thin_wgrad: bad argument");
  (void)backend; (void)precision;
  return thin_wgrad(make_operand(wide, 32), make_operand(thin, 3), dw, dbias_thin, partials, partials_bytes, batch,
                    (cudaStream_t)stream);
}

int ae_thin_bwd_fused(const ae_operand_t* wide, const ae_operand_t* thin, const float* w, const ae_epilogue_t* epi,
                      float* out_wide, float* dw, float* dbias_thin, void* partials, size_t partials_bytes, int batch,
                      int precision, int backend, ae_stream_t stream) {
  AE_CHECK(wide && thin && w && epi && out_wide && dw && partials && batch >= 1, "ae_thin_bwd_fused: bad argument");
  (void)backend; (void)precision;
  return thin_bwd_fused(make_operand(wide, 32), make_operand(thin, 3), w, make_epilogue(epi, 32), out_wide, dw, dbias_thin,
                        partials, partials_bytes, batch, (cudaStream_t)stream, 0);
}

int ae_bn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float* bnc, int channels, int training, ae_stream_t stream) {
  AE_CHECK(gamma && beta && bnc && channels >= 1, "ae_bn_finalize: bad argument");
  return bn_finalize(stats, count, gamma, beta, running_mean, running_var, bnc, channels, training, (cudaStream_t)stream);
}

int ae_bn_bwd_reduce(const double* stats, int64_t count, const float* gamma, float* bnc, float* dgamma, float* dbeta,
                     int channels, ae_stream_t stream) {
  AE_CHECK(stats && gamma && bnc && channels >= 1, "ae_bn_bwd_reduce: bad argument");
  return bn_bwd_reduce(stats, count, gamma, bnc, dgamma, dbeta, channels, (cudaStream_t)stream);
}

// ---- dense layers (convenience form: packs into the workspace, then runs the GEMM kernels) -------------------
size_t ae_linear_workspace_bytes(int m, int n, int k) {
  size_t pack = (size_t)n * k * 4 + 256;
  size_t part = (size_t)64 * ((size_t)m * n > (size_t)n * k ? (size_t)m * n : (size_t)n * k) * 4 + 256;
  size_t dap = (size_t)16 * m * k * 4 + 256;
  return 2 * pack + part + dap + (size_t)n * 4 + 1024;
}

static int perm_params(int perm, int dim, int* pc, int* phw) {
  *pc = 0; *phw = 0;
  if (!perm) return 0;
  AE_CHECK(dim % 16 == 0, "linear: permuted dimension %d must be C*16 (4x4 spatial)", dim);
  *pc = dim / 16; *phw = 16;
  return 0;
}

int ae_linear_fwd(const ae_operand_t* a, int a_channels, const float* w, const float* bias, float* out, int m, int n, int k,
                  int perm_k_hw, int perm_n_hw, void* workspace, size_t workspace_bytes, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(a && w && out && workspace, "ae_linear_fwd: null argument");
  AE_CHECK(!(perm_k_hw && perm_n_hw), "ae_linear_fwd: only one axis can be permuted");
  AE_CHECK(n % 4 == 0 && k % 16 == 0, "ae_linear_fwd: n=%d must be a multiple of 4 and k=%d of 16", n, k);
  AE_CHECK(workspace_bytes >= ae_linear_workspace_bytes(m, n, k), "ae_linear_fwd: workspace too small");
  char* ws = (char*)workspace;
  float* pk = (float*)ws; ws += ((size_t)n * k * 4 + 255) & ~(size_t)255;
  float* pb = (float*)ws; ws += ((size_t)n * 4 + 255) & ~(size_t)255;
  int pc, phw;
  const float* bias_use = bias;
  if (perm_n_hw) {
    AE_TRY(perm_params(1, n, &pc, &phw));
    AE_TRY(pack_linear(w, n, k, pc, phw, 2, pk, st));
    if (bias) { AE_TRY(permute_vector(bias, n, pc, phw, pb, st)); bias_use = pb; }
  } else {
    AE_TRY(perm_params(perm_k_hw, k, &pc, &phw));
    AE_TRY(pack_linear(w, n, k, pc, phw, 0, pk, st));
  }
  RowGemm r{};
  r.family = FAM_DENSE; r.M = m; r.N = n; r.K = k;
  r.A = make_operand(a, a_channels > 0 ? a_channels : 1);
  r.Bp = pk; r.epi = store_epilogue(bias_use); r.out = out; r.splitK = 1;
  return simt_rowgemm(r, st);
}

int ae_linear_bwd(const ae_operand_t* a, int a_channels, const float* w, const float* d_out, float* d_a,
                  const ae_epilogue_t* d_a_epi, int d_a_channels, float* dw, float* db, int m, int n, int k,
                  int perm_k_hw, int perm_n_hw, void* workspace, size_t workspace_bytes, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(a && w && d_out && workspace, "ae_linear_bwd: null argument");
  AE_CHECK(!(perm_k_hw && perm_n_hw), "ae_linear_bwd: only one axis can be permuted");
  AE_CHECK(n % 16 == 0 && k % 16 == 0, "ae_linear_bwd: n=%d and k=%d must be multiples of 16", n, k);
  AE_CHECK(workspace_bytes >= ae_linear_workspace_bytes(m, n, k), "ae_linear_bwd: workspace too small");
  char* ws = (char*)workspace;
  float* pk = (float*)ws; ws += ((size_t)n * k * 4 + 255) & ~(size_t)255;
  float* part = (float*)ws;
  int pc, phw;
  if (d_a) {  // dA[m][k] = sum_n dOut[m][n] * w[n][k]   (row GEMM with reduction dim n)
    if (perm_n_hw) { AE_TRY(perm_params(1, n, &pc, &phw)); AE_TRY(pack_linear(w, n, k, pc, phw, 3, pk, st)); }
    else { AE_TRY(perm_params(perm_k_hw, k, &pc, &phw)); AE_TRY(pack_linear(w, n, k, pc, phw, 1, pk, st)); }
    RowGemm r{};
    r.family = FAM_DENSE; r.M = m; r.N = k; r.K = n;
    r.A = raw_operand(d_out); r.Bp = pk; r.epi = make_epilogue(d_a_epi, d_a_channels > 0 ? d_a_channels : 1);
    r.out = d_a; r.splitK = 1;
    AE_TRY(simt_rowgemm(r, st));
  }
  if (dw) {
    ColGemm c{};
    c.M = m; c.out = dw;
    if (perm_n_hw) {        // dw[perm(n')][k] = sum_m dOut[m][n'] * A(m,k)
      AE_TRY(perm_params(1, n, &pc, &phw));
      c.gather = 0; c.I = n; c.J = k; c.A = raw_operand(d_out); c.B = make_operand(a, a_channels > 0 ? a_channels : 1);
      c.permC = pc; c.permHW = phw; c.transposed = 0;
    } else {                // dw[n][perm(k)] = sum_m A(m,k) * dOut[m][n]
      AE_TRY(perm_params(perm_k_hw, k, &pc, &phw));
      c.gather = 0; c.I = k; c.J = n; c.A = make_operand(a, a_channels > 0 ? a_channels : 1); c.B = raw_operand(d_out);
      c.permC = pc; c.permHW = phw; c.transposed = 1;
    }
    c.splitK = colgemm_default_split(c.M, c.I, c.J);
    c.partial = part;
    AE_TRY(simt_colgemm(c, st));
  }
  if (db) {
    if (perm_n_hw) { AE_TRY(perm_params(1, n, &pc, &phw)); } else { pc = 0; phw = 0; }
    AE_TRY(column_sums(d_out, m, n, pc, phw, db, st));
  }
  return 0;
}

int ae_softmax_ce_fwd_bwd(const float* logits, const int64_t* labels, int batch, int classes, float grad_scale, float* loss,
                          float* d_logits, int* correct, ae_stream_t stream) {
  AE_CHECK(logits && labels && loss && batch >= 1 && classes >= 2, "ae_softmax_ce_fwd_bwd: bad argument");
  return softmax_ce(logits, labels, batch, classes, grad_scale, loss, d_logits, correct, nullptr, 1.0, 0.f, (cudaStream_t)stream);
}

int ae_sigmoid_mse_fwd_bwd(const float* x_hat, const float* x, int64_t numel, float scale, float* loss, float* d_pre,
                           ae_stream_t stream) {
  AE_CHECK(x_hat && x && loss && numel >= 1, "ae_sigmoid_mse_fwd_bwd: bad argument");
  // the fp64 accumulator lives right behind the fp32 loss slot the caller provides: loss must have room for 4 floats
  double* tmp = reinterpret_cast<double*>(loss + 2);
  AE_CHECK(((uintptr_t)tmp & 7) == 0, "ae_sigmoid_mse_fwd_bwd: loss must be 8-byte aligned with 4 floats of room");
  return sigmoid_mse(x_hat, x, numel, scale, loss, d_pre, tmp, (cudaStream_t)stream);
}

int ae_adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, float grad_scale, int* step_dev, ae_stream_t stream) {
  AE_CHECK(p && g && m && v && step_dev, "ae_adam_step_flat: null argument");
  return adam_step_flat(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev, (cudaStream_t)stream);
}

int ae_layout_nchw_f32_to_nhwc_f32(const float* src, float* dst, int n, int c, int h, int w, ae_stream_t stream) {
  return layout_convert(src, dst, n, c, h, w, false, true, (cudaStream_t)stream);
}
int ae_layout_nhwc_f32_to_nchw_f32(const float* src, float* dst, int n, int c, int h, int w, ae_stream_t stream) {
  return layout_convert(src, dst, n, c, h, w, false, false, (cudaStream_t)stream);
}
int ae_layout_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int n, int c, int h, int w, ae_stream_t stream) {
  return layout_convert(src, dst, n, c, h, w, true, true, (cudaStream_t)stream);
}
int ae_layout_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int n, int c, int h, int w, ae_stream_t stream) {
  return layout_convert(src, dst, n, c, h, w, true, false, (cudaStream_t)stream);
}

}  // extern "C"

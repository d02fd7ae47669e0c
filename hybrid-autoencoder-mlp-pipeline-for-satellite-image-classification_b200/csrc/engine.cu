// Engine: orchestrates the kernels of the supervised autoencoder (NB:685-702) over caller-owned flat
// parameter / gradient buffers and one caller-owned workspace.  Host-side only; every launch goes to the
// stream the caller passes, so whole steps can be captured in a CUDA graph (ae_step_graph_*).
#include <cstdlib>
#include <vector>

#include "pack.cuh"

namespace ae {

// defined in the other translation units
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
int adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                   float wd, float gscale, int* step_dev, cudaStream_t st);
int adam_step_flat_range(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                         float wd, float gscale, int* step_dev, int bump, cudaStream_t st);
int head_out(const float* hid_pre, const float* w2, const float* b2, float* logits, int B, int H, int C, cudaStream_t st);
size_t head_fused_workspace_floats(int B, int L, int C);
int head_fused_step(const float* z, const int64_t* labels, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* logits, float* dz, float* gw1, float* gb1, float* gw2, float* gb2,
                    float* loss, const double* sse, double numel, float alpha, float* partial, unsigned int* counter,
                    int B, int L, int C, cudaStream_t st_main, cudaStream_t st_finish, int phase);
int head_backward_small(const float* dlogits, const float* w2, const float* hid_pre, float* dhid, float* dw2, float* db2,
                        int B, int H, int C, cudaStream_t st);

static inline int64_t pad4(int64_t v) { return (v + 3) & ~(int64_t)3; }

struct Tensor { int64_t off, size; };

struct MidLayer {   // one of the six 32..256-channel stride-2 layers
  Geom g;
  int w, b;         // tensor indices inside the part
  void* pk_fwd = nullptr;
  void* pk_dgrad = nullptr;
  size_t pk_bytes = 0;
};

struct BN {
  int C;
  int gamma, beta;  // tensor indices
  int64_t count_per_image;
  float* bnc = nullptr;
  double* stats_f = nullptr;
  double* stats_b = nullptr;
};

struct Part {
  std::vector<Tensor> t;
  int64_t flat_len = 0;
  float* params = nullptr;
  float* grads = nullptr;
  float* running = nullptr;
  int64_t* steps = nullptr;
  std::vector<BN> bn;
  std::vector<int64_t> run_off;  // offset of (mean, var) pair start per BN layer inside `running`
  void add(int64_t size) { t.push_back({flat_len, size}); flat_len += pad4(size); }
  float* P(int i) const { return params + t[i].off; }
  float* G(int i) const { return grads + t[i].off; }
  float* rmean(int l) const { return running ? running + run_off[l] : nullptr; }
  float* rvar(int l) const { return running ? running + run_off[l] + bn[l].C : nullptr; }
};

}  // namespace ae

using namespace ae;

struct ae_engine {
  ae_engine_config_t cfg;
  int L, NC, Bmax;
  bool simt;
  bool dense_tc = true;               // AE_B200_DENSE_TC=0: eval-mode Linear(4096, L) stays on the CUDA-core kernel
  int nsplit;             // bf16 operand split terms of the tcgen05 path
  Part part[AE_NUM_PARTS];
  MidLayer enc_mid[3];    // conv2..conv4
  MidLayer dec_mid[3];    // convT1..convT3
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0, ws_need = 0;
  // activations
  float *y[4] = {}, *dzy[4] = {};     // encoder raw conv outputs / masked gradients
  float *h = nullptr, *dh = nullptr;  // decoder_input output [B,4,4,256] NHWC and its gradient
  float *t[3] = {}, *dzt[3] = {};     // decoder raw convT outputs / masked gradients
  float *xhat = nullptr;
  // split-bf16 operand planes of the tcgen05 path (see tma_gemm.cu)
  void *ae_pl[3] = {}, *ad_pl[2] = {}, *h_pl = nullptr, *dy_pl = nullptr;
  float *z = nullptr, *dz_dec = nullptr, *dz_head = nullptr, *dz_tot = nullptr;
  float *hid_pre = nullptr, *dhid = nullptr, *logits = nullptr, *dlogits = nullptr;
  // packs
  void* encfc_tc = nullptr;           // tensor-core pack of the encoder's Linear(4096, L) (eval mode, dense_tc.cu)
  float *encfc_fwd = nullptr, *encfc_bwd = nullptr, *decfc_fwd = nullptr, *decfc_bwd = nullptr, *decfc_bias = nullptr;
  float *head_w1t = nullptr;
  // scratch
  float* partial = nullptr;
  size_t partial_bytes = 0;
  double* sse = nullptr;
  unsigned int* head_counter = nullptr;
  double* stats_base = nullptr;
  size_t stats_bytes = 0;
  int fc_split = 64;
  // second stream of the fused step: independent kernels (a layer's weight gradient next to its data gradient, the
  // classifier head next to the decoder, the decoder's gradient allreduce next to the encoder backward) become parallel
  // branches of the captured graph.  Created on first use (the engine can be created without a device for layout queries).
  cudaStream_t side = nullptr, side2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  float* head_partial = nullptr;      // the head's per-CTA partials (it runs beside kernels that use `partial`)
  float* partial_side = nullptr;      // split-K partials of the dense weight gradients when they run on the side stream
  size_t partial_side_bytes = 0;
  // data-parallel step being captured: gradient exchange interleaved with the backward pass
  ae_dp_comm_t* step_comm = nullptr;
  bool dp_single = false;             // data parallel over peer memory: no NCCL exchange inside the backward pass (k_dp_adam does it all)
  bool defer_conv1_wgrad = false;     // the captured step runs conv1's weight gradient beside Adam + re-pack of everything else
  bool stats_cleared = false;         // the fused step zeroed every statistic accumulator in one memset: the parts skip theirs
  // pointers remembered between forward and backward
  const float* last_x = nullptr;
  const float* last_z_dec = nullptr;
  const float* last_z_head = nullptr;
  int last_batch = 0;
};

static void build_layouts(ae_engine* e) {
  const int L = e->L, NC = e->NC;
  Part& E = e->part[AE_PART_ENC];
  const int ch[5] = {3, 32, 64, 128, 256};
  int hw = 32;
  for (int i = 0; i < 4; ++i) {
    E.add((int64_t)ch[i + 1] * ch[i] * 9);  // conv weight
    E.add(ch[i + 1]);                       // conv bias
    E.add(ch[i + 1]);                       // bn gamma
    E.add(ch[i + 1]);                       // bn beta
    BN b; b.C = ch[i + 1]; b.gamma = 4 * i + 2; b.beta = 4 * i + 3; b.count_per_image = (int64_t)hw * hw;
    E.bn.push_back(b);
    hw /= 2;
  }
  E.add((int64_t)L * 4096);
  E.add(L);
  Part& D = e->part[AE_PART_DEC];
  D.add((int64_t)4096 * L);
  D.add(4096);
  const int dch[5] = {256, 128, 64, 32, 3};
  hw = 8;
  for (int i = 0; i < 4; ++i) {
    D.add((int64_t)dch[i] * dch[i + 1] * 9);
    D.add(dch[i + 1]);
    if (i < 3) {
      D.add(dch[i + 1]);
      D.add(dch[i + 1]);
      BN b; b.C = dch[i + 1]; b.gamma = 2 + 4 * i + 2; b.beta = 2 + 4 * i + 3; b.count_per_image = (int64_t)hw * hw;
      D.bn.push_back(b);
      hw *= 2;
    }
  }
  Part& H = e->part[AE_PART_HEAD];
  H.add((int64_t)128 * L);
  H.add(128);
  H.add((int64_t)NC * 128);
  H.add(NC);
  for (int p = 0; p < AE_NUM_PARTS; ++p) {
    int64_t off = 0;
    for (auto& b : e->part[p].bn) { e->part[p].run_off.push_back(off); off += 2 * b.C; }
  }
  // mid layers: encoder conv2..4 (big -> small), decoder convT1..3 (small -> big)
  const int ehs[3] = {16, 8, 4};
  for (int i = 0; i < 3; ++i) {
    MidLayer& m = e->enc_mid[i];
    m.g.B = 0; m.g.Hs = ehs[i]; m.g.Ws = ehs[i]; m.g.Cb = ch[i + 1]; m.g.Cs = ch[i + 2];
    m.g.lHs = ilog2(ehs[i]); m.g.lWs = m.g.lHs;
    m.w = 4 * (i + 1); m.b = 4 * (i + 1) + 1;
  }
  const int dhs[3] = {4, 8, 16};
  for (int i = 0; i < 3; ++i) {
    MidLayer& m = e->dec_mid[i];
    m.g.B = 0; m.g.Hs = dhs[i]; m.g.Ws = dhs[i]; m.g.Cs = dch[i]; m.g.Cb = dch[i + 1];
    m.g.lHs = ilog2(dhs[i]); m.g.lWs = m.g.lHs;
    m.w = 2 + 4 * i; m.b = 2 + 4 * i + 1;
  }
}

// Carve the workspace.  Called with base == nullptr to measure.
static size_t carve(ae_engine* e, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    off = (off + 255) & ~(size_t)255;
    char* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t B = (size_t)e->Bmax;
  const int L = e->L, NC = e->NC;
  const size_t ysz[4] = {B * 32 * 32 * 32, B * 16 * 16 * 64, B * 8 * 8 * 128, B * 4 * 4 * 256};
  for (int i = 0; i < 4; ++i) e->y[i] = (float*)take(ysz[i] * 4);
  for (int i = 0; i < 4; ++i) e->dzy[i] = (float*)take(ysz[i] * 4);
  e->h = (float*)take(B * 4096 * 4);
  e->dh = (float*)take(B * 4096 * 4);
  const size_t tsz[3] = {B * 8 * 8 * 128, B * 16 * 16 * 64, B * 32 * 32 * 32};
  for (int i = 0; i < 3; ++i) e->t[i] = (float*)take(tsz[i] * 4);
  for (int i = 0; i < 3; ++i) e->dzt[i] = (float*)take(tsz[i] * 4);
  e->xhat = (float*)take(B * 3 * 64 * 64 * 4);
  if (!e->simt) {
    const size_t pb = (size_t)2 * e->nsplit;   // bytes per element of a split-bf16 tensor
    for (int i = 0; i < 3; ++i) e->ae_pl[i] = take(ysz[i] * pb);
    for (int i = 0; i < 2; ++i) e->ad_pl[i] = take(tsz[i] * pb);
    e->h_pl = take(B * 4096 * pb);
    e->dy_pl = take(tsz[2] * pb);
  }
  e->z = (float*)take(B * L * 4);
  e->dz_dec = (float*)take(B * L * 4);
  e->dz_head = (float*)take(B * L * 4);
  e->dz_tot = (float*)take(B * L * 4);
  e->hid_pre = (float*)take(B * 128 * 4);
  e->dhid = (float*)take(B * 128 * 4);
  e->logits = (float*)take(B * (size_t)pad4(NC) * 4);
  e->dlogits = (float*)take(B * (size_t)pad4(NC) * 4);
  // BN coefficient blocks and statistics
  char* stats0 = nullptr;
  size_t stats_begin = 0;
  for (int p = 0; p < 2; ++p)
    for (auto& b : e->part[p].bn) b.bnc = (float*)take((size_t)AE_BNC_ROWS * b.C * 4);
  off = (off + 255) & ~(size_t)255;
  stats_begin = off;
  stats0 = base ? base + off : nullptr;
  for (int p = 0; p < 2; ++p)
    for (auto& b : e->part[p].bn) {
      b.stats_f = (double*)(base ? base + off : nullptr); off += (size_t)2 * b.C * 8;
      b.stats_b = (double*)(base ? base + off : nullptr); off += (size_t)2 * b.C * 8;
    }
  e->sse = (double*)(base ? base + off : nullptr); off += 16;
  e->stats_base = (double*)stats0;
  e->stats_bytes = off - stats_begin;
  // packed weights
  for (int i = 0; i < 3; ++i) {
    for (MidLayer* m : {&e->enc_mid[i], &e->dec_mid[i]}) {
      const size_t bytes = e->simt ? (size_t)9 * m->g.Cb * m->g.Cs * 4 : tma_packed_bytes(m->g.Cs, m->g.Cb, e->nsplit);
      m->pk_bytes = bytes;
      m->pk_fwd = take(bytes);
      m->pk_dgrad = take(bytes);
    }
  }
  e->encfc_fwd = (float*)take((size_t)4096 * L * 4);
  e->encfc_bwd = (float*)take((size_t)4096 * L * 4);
  e->decfc_fwd = (float*)take((size_t)4096 * L * 4);
  e->decfc_bwd = (float*)take((size_t)4096 * L * 4);
  e->decfc_bias = (float*)take(4096 * 4);
  e->head_w1t = (float*)take((size_t)128 * L * 4);
  e->encfc_tc = (!e->simt && e->dense_tc && dense_tc_supported(L, 4096)) ? take(dense_tc_pack_bytes(L, 4096, e->nsplit)) : nullptr;
  // split-K / weight-gradient partial buffer
  size_t pb = 0;
  auto upd = [&](size_t v) { if (v > pb) pb = v; };
  for (int i = 0; i < 3; ++i)
    for (MidLayer* m : {&e->enc_mid[i], &e->dec_mid[i]}) {
      const int I = 9 * m->g.Cb, J = m->g.Cs;
      const int Mrows = (int)(B * m->g.Hs * m->g.Ws);
      upd((size_t)colgemm_default_split(Mrows, I, J) * I * J * 4);
      if (!e->simt) {
        // the slice count grows with the batch up to one slice per SM-wave: bound it over all batches <= Bmax
        upd((size_t)148 * I * J * 4);
      }
    }
  upd((size_t)colgemm_default_split((int)B, 4096, L) * 4096 * L * 4);
  upd((size_t)colgemm_default_split((int)B, 128, L) * 128 * L * 4);
  upd((size_t)e->fc_split * B * L * 4);
  upd(thin_wgrad_workspace_bytes((int)B));
  upd(head_fused_workspace_floats((int)B, L, NC) * 4);
  e->partial_bytes = pb;
  e->partial = (float*)take(pb);
  e->head_partial = (float*)take(head_fused_workspace_floats((int)B, L, NC) * 4);
  e->partial_side_bytes = (size_t)colgemm_default_split((int)B, 4096, L) * 4096 * L * 4;
  e->partial_side = (float*)take(e->partial_side_bytes);
  e->head_counter = (unsigned int*)take(256);      // [0..31]: head; [40]: last-CTA counter of a row GEMM's BatchNorm tail job
  return off + 256;
}

static int ensure_side(ae_engine* e) {
  if (e->side) return 0;
  AE_CUDA(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  AE_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  AE_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  AE_CUDA(cudaStreamCreateWithFlags(&e->side2, cudaStreamNonBlocking));
  AE_CUDA(cudaEventCreateWithFlags(&e->ev_fork2, cudaEventDisableTiming));
  AE_CUDA(cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming));
  return 0;
}
// side stream continues from everything issued on `st` so far
static int fork_side(ae_engine* e, cudaStream_t st) {
  AE_TRY(ensure_side(e));
  AE_CUDA(cudaEventRecord(e->ev_fork, st));
  AE_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
  return 0;
}
// `st` continues after everything issued on the side stream so far
static int join_side(ae_engine* e, cudaStream_t st) {
  AE_CUDA(cudaEventRecord(e->ev_join, e->side));
  AE_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
  return 0;
}

// split-K of the two 4096-deep dense GEMMs.  Fewer, deeper splits at large batches were tried (8 instead of 64 at batch
// 4096: 89 + 7 us against 65 + 13 us for GEMM + reduce, profiles/r1_v8_every_kernel_full.txt): the CUDA-core kernel is
// faster with many short K loops, so the split stays fixed.
static int fc_split_for(const ae_engine* e, int batch) {
  (void)batch;
  return e->fc_split;
}

static int check_batch(const ae_engine* e, int batch) {
  AE_CHECK(e->ws != nullptr, "engine: workspace not bound (ae_engine_bind_workspace)");
  AE_CHECK(batch >= 1 && batch <= e->Bmax, "engine: batch %d outside [1, max_batch=%d]", batch, e->Bmax);
  return 0;
}
static int check_part(const ae_engine* e, int p, bool need_grads) {
  AE_CHECK(e->part[p].params != nullptr, "engine: part %d has no parameters bound (ae_engine_bind_part)", p);
  AE_CHECK(!need_grads || e->part[p].grads != nullptr, "engine: part %d has no gradient buffer bound", p);
  return 0;
}

static BnJob bn_fwd_job(const Part& P, int l, int batch, int training) {
  const BN& b = P.bn[l];
  BnJob j{};
  j.kind = BN_JOB_FINALIZE; j.stats = b.stats_f; j.count = (double)batch * (double)b.count_per_image;
  j.gamma = P.P(b.gamma); j.beta = P.P(b.beta); j.rmean = P.rmean(l); j.rvar = P.rvar(l); j.bnc = b.bnc;
  j.dgamma = nullptr; j.dbeta = nullptr; j.C = b.C; j.training = training;
  j.nbt = (training && P.steps) ? P.steps + l : nullptr;     // num_batches_tracked += 1 rides on the layer's finalize job
  j.dzero = nullptr;
  return j;
}
// `dzero`: gradient slot of the bias in front of this BatchNorm (exactly zero: the batch mean removes it)
static BnJob bn_bwd_job(const Part& P, int l, int batch, float* dzero) {
  const BN& b = P.bn[l];
  BnJob j{};
  j.kind = BN_JOB_BWD; j.stats = b.stats_b; j.count = (double)batch * (double)b.count_per_image;
  j.gamma = P.P(b.gamma); j.beta = nullptr; j.rmean = nullptr; j.rvar = nullptr; j.bnc = b.bnc;
  j.dgamma = P.G(b.gamma); j.dbeta = P.G(b.beta); j.C = b.C; j.training = 1;
  j.nbt = nullptr; j.dzero = dzero;
  return j;
}

// Row GEMM of a mid layer.  `a` is the fp32 operand description (transform included); on the tcgen05 path its
// split-bf16 planes are `planes`: produced here when `split_now`, else already current (written earlier this step).
// `job`: the BatchNorm coefficient job of the operand's layer; it runs inside the split kernel (tcgen05 path) or as its
// own launch (CUDA-core path).
static int run_rowgemm(ae_engine* e, RowGemm& r, const Operand& a, void* planes, bool split_now, int64_t a_count,
                       const void* pk, const BnJob* job, cudaStream_t st) {
  if (e->simt) {
    if (job) AE_TRY(run_bn_job(*job, st));
    r.A = a;
    return simt_rowgemm(r, st);
  }
  if (split_now) AE_TRY(tma_split_operand(a, a_count, planes, e->nsplit, job, st));
  else if (job) AE_TRY(run_bn_job(*job, st));
  r.A = split_operand(planes, a.C);
  return tma_rowgemm(r, pk, e->nsplit, st);
}

// Weight gradient of a mid layer: big / small are the fp32 operand descriptions; on the tcgen05 path the planes
// must already be current.
static int run_conv_wgrad(ae_engine* e, const Geom& g, const Operand& big, const Operand& small, const void* big_pl,
                          const void* small_pl, float* dw, cudaStream_t st) {
  if (!e->simt)
    return tma_wgrad(g, big_pl, small_pl, dw, e->partial, e->partial_bytes, e->nsplit, st);
  ColGemm c{};
  c.gather = 1; c.g = g; c.M = g.B * g.Hs * g.Ws; c.I = 9 * g.Cb; c.J = g.Cs;
  c.A = big; c.B = small;
  c.out = dw; c.permC = g.Cb; c.permHW = 9; c.transposed = 1;
  c.splitK = colgemm_default_split(c.M, c.I, c.J);
  c.partial = e->partial;
  AE_CHECK((size_t)c.splitK * c.I * c.J * 4 <= e->partial_bytes, "engine: partial buffer too small");
  return simt_colgemm(c, st);
}

static int run_wgrad(ae_engine* e, ColGemm& c, cudaStream_t st, bool side = false) {   // dense layers
  c.splitK = colgemm_default_split(c.M, c.I, c.J);
  c.partial = side ? e->partial_side : e->partial;
  AE_CHECK((size_t)c.splitK * c.I * c.J * 4 <= (side ? e->partial_side_bytes : e->partial_bytes), "engine: partial buffer too small");
  return simt_colgemm(c, st);
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int ae_engine_create(const ae_engine_config_t* cfg, ae_engine_t** out) {
  AE_CHECK(cfg && out, "ae_engine_create: null argument");
  AE_CHECK(cfg->latent_dim >= 16 && cfg->latent_dim % 16 == 0 && cfg->latent_dim <= 1024,
           "ae_engine_create: latent_dim=%d must be a multiple of 16 in [16,1024]", cfg->latent_dim);
  AE_CHECK(cfg->num_classes >= 2 && cfg->num_classes <= 64, "ae_engine_create: num_classes=%d out of range", cfg->num_classes);
  AE_CHECK(cfg->max_batch >= 1, "ae_engine_create: max_batch must be >= 1");
  AE_CHECK(cfg->precision == AE_PREC_FP32 || cfg->precision == AE_PREC_BF16, "ae_engine_create: bad precision");
  AE_CHECK(cfg->backend == AE_BACKEND_TC || cfg->backend == AE_BACKEND_SIMT, "ae_engine_create: bad backend");
  ae_engine* e = new ae_engine();
  e->cfg = *cfg;
  e->L = cfg->latent_dim; e->NC = cfg->num_classes; e->Bmax = cfg->max_batch;
  e->simt = cfg->backend == AE_BACKEND_SIMT;
  { const char* v = getenv("AE_B200_DENSE_TC"); e->dense_tc = !(v && v[0] == '0'); }
  e->nsplit = cfg->precision == AE_PREC_FP32 ? 2 : 1;
  build_layouts(e);
  e->ws_need = carve(e, nullptr);
  *out = e;
  return 0;
}

void ae_engine_destroy(ae_engine_t* e) {
  if (!e) return;
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_fork2) cudaEventDestroy(e->ev_fork2);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->side2) cudaStreamDestroy(e->side2);
  delete e;
}

int ae_engine_param_layout(const ae_engine_t* e, int part, int64_t* offsets, int64_t* sizes, int64_t* flat_len) {
  if (!e || part < 0 || part >= AE_NUM_PARTS) return -1;
  const Part& p = e->part[part];
  for (size_t i = 0; i < p.t.size(); ++i) {
    if (offsets) offsets[i] = p.t[i].off;
    if (sizes) sizes[i] = p.t[i].size;
  }
  if (flat_len) *flat_len = p.flat_len;
  return (int)p.t.size();
}

int ae_engine_bn_layout(const ae_engine_t* e, int part, int* channels) {
  if (!e || part < 0 || part >= AE_NUM_PARTS) return -1;
  const Part& p = e->part[part];
  for (size_t i = 0; i < p.bn.size(); ++i)
    if (channels) channels[i] = p.bn[i].C;
  return (int)p.bn.size();
}

size_t ae_engine_workspace_bytes(const ae_engine_t* e) { return e ? e->ws_need : 0; }

int ae_engine_bind_workspace(ae_engine_t* e, void* workspace, size_t bytes) {
  AE_CHECK(e && workspace, "ae_engine_bind_workspace: null argument");
  AE_CHECK(bytes >= e->ws_need, "ae_engine_bind_workspace: %zu bytes given, %zu needed", bytes, e->ws_need);
  AE_CHECK(((uintptr_t)workspace & 255) == 0, "ae_engine_bind_workspace: workspace must be 256-byte aligned");
  e->ws = workspace; e->ws_bytes = bytes;
  carve(e, static_cast<char*>(workspace));
  AE_CUDA(cudaMemset(e->head_counter, 0, 256));
  return 0;
}

int ae_engine_bind_part(ae_engine_t* e, int part, float* params, float* grads, float* bn_running, int64_t* bn_steps) {
  AE_CHECK(e && part >= 0 && part < AE_NUM_PARTS, "ae_engine_bind_part: bad part");
  AE_CHECK(params != nullptr, "ae_engine_bind_part: params must not be null");
  AE_CHECK((((uintptr_t)params | (uintptr_t)grads) & 15) == 0, "ae_engine_bind_part: buffers must be 16-byte aligned");
  Part& p = e->part[part];
  p.params = params; p.grads = grads; p.running = bn_running; p.steps = bn_steps;
  return 0;
}

int ae_engine_pack_weights(ae_engine_t* e, int part, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(e && e->ws, "ae_engine_pack_weights: workspace not bound");
  AE_TRY(check_part(e, part, false));
  const Part& p = e->part[part];
  const int L = e->L;
  if (part == AE_PART_ENC || part == AE_PART_DEC) {
    MidLayer* mids = part == AE_PART_ENC ? e->enc_mid : e->dec_mid;
    for (int i = 0; i < 3; ++i) {
      MidLayer& m = mids[i];
      if (e->simt) AE_TRY(pack_conv_simt(p.P(m.w), m.g.Cs, m.g.Cb, (float*)m.pk_fwd, (float*)m.pk_dgrad, st));
      else AE_TRY(tma_pack_conv(p.P(m.w), m.g.Cs, m.g.Cb, e->nsplit, m.pk_fwd, m.pk_dgrad, st));
    }
  }
  if (part == AE_PART_ENC) {
    AE_TRY(pack_linear(p.P(16), L, 4096, 256, 16, 0, e->encfc_fwd, st));   // [4096 nhwc][L]
    AE_TRY(pack_linear(p.P(16), L, 4096, 256, 16, 1, e->encfc_bwd, st));   // [L][4096 nhwc]
    if (e->encfc_tc) AE_TRY(dense_tc_pack(p.P(16), L, 4096, 256, 16, e->nsplit, e->encfc_tc, st));
  } else if (part == AE_PART_DEC) {
    AE_TRY(pack_linear(p.P(0), 4096, L, 256, 16, 2, e->decfc_fwd, st));    // [L][4096 nhwc]
    AE_TRY(pack_linear(p.P(0), 4096, L, 256, 16, 3, e->decfc_bwd, st));    // [4096 nhwc][L]
    AE_TRY(permute_vector(p.P(1), 4096, 256, 16, e->decfc_bias, st));
  } else {
    AE_TRY(pack_linear(p.P(0), 128, L, 0, 0, 0, e->head_w1t, st));         // [L][128]
  }
  return 0;
}

// Every weight re-layout of all three parts in ONE launch (tcgen05 path; the captured step uses this after Adam).
static int pack_all_parts(ae_engine* e, cudaStream_t st) {
  for (int p = 0; p < AE_NUM_PARTS; ++p) AE_TRY(check_part(e, p, false));
  const int L = e->L;
  PackJobs J{};
  int n = 0;
  auto conv = [&](const Part& P, MidLayer& m) {
    PackJob& j = J.job[n++];
    j.kind = PACK_CONV; j.src = P.P(m.w); j.dst = m.pk_fwd; j.dst2 = m.pk_dgrad;
    j.a = m.g.Cs; j.b = m.g.Cb; j.c = e->nsplit; j.total = 2 * 9 * m.g.Cs * m.g.Cb / 8;
  };
  auto lin = [&](const float* w, int N, int K, int permC, int permHW, int kind, float* dst) {
    PackJob& j = J.job[n++];
    j.kind = PACK_LINEAR; j.src = w; j.dst = dst; j.dst2 = nullptr;
    j.a = N; j.b = K; j.c = permC; j.d = permHW; j.e = kind; j.total = N * K;
  };
  const Part& E = e->part[AE_PART_ENC];
  const Part& D = e->part[AE_PART_DEC];
  const Part& H = e->part[AE_PART_HEAD];
  for (int i = 0; i < 3; ++i) { conv(E, e->enc_mid[i]); conv(D, e->dec_mid[i]); }
  lin(E.P(16), L, 4096, 256, 16, 0, e->encfc_fwd);
  lin(E.P(16), L, 4096, 256, 16, 1, e->encfc_bwd);
  lin(D.P(0), 4096, L, 256, 16, 2, e->decfc_fwd);
  lin(D.P(0), 4096, L, 256, 16, 3, e->decfc_bwd);
  lin(H.P(0), 128, L, 0, 0, 0, e->head_w1t);
  if (e->encfc_tc) {
    PackJob& j = J.job[n++];
    j.kind = PACK_DENSE_TC; j.src = E.P(16); j.dst = e->encfc_tc; j.dst2 = nullptr;
    j.a = L; j.b = 4096; j.c = 256; j.d = 16; j.e = e->nsplit; j.total = L * 4096 / 8;
  }
  {
    PackJob& j = J.job[n++];
    j.kind = PACK_PERMUTE; j.src = D.P(1); j.dst = e->decfc_bias; j.dst2 = nullptr;
    j.a = 4096; j.b = 256; j.c = 16; j.total = 4096;
  }
  J.n = n;
  return pack_all(J, st);
}

// ---------------------------------------------------------------------------------------------
// Encoder (NB:499-525)
// ---------------------------------------------------------------------------------------------
int ae_encoder_forward(ae_engine_t* e, const float* x, int batch, int training, float* z, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_ENC, false));
  Part& P = e->part[AE_PART_ENC];
  AE_CHECK(!training || P.running, "ae_encoder_forward: training mode needs BatchNorm running buffers");
  if (training && !e->stats_cleared) {
    AE_CUDA(cudaMemsetAsync(P.bn[0].stats_f, 0, (char*)(P.bn[3].stats_b + 2 * P.bn[3].C) - (char*)P.bn[0].stats_f, st));
  }
  // Eval mode on the tcgen05 path: the BatchNorm coefficients come from the running statistics, so they are known before
  // the first kernel runs and every layer's epilogue can emit the next layer's operand directly -- split-bf16 planes of
  // relu(bn(conv + bias)) -- instead of an fp32 tensor that a second pass would normalise and split.
  const bool fused_eval = !training && !e->simt;
  if (fused_eval)
    for (int l = 0; l < 4; ++l) AE_TRY(run_bn_job(bn_fwd_job(P, l, batch, 0), st));
  // conv1 (3 -> 32): x NCHW fp32 -> y1 NHWC
  {
    Epilogue ep = training ? bias_stats_epilogue(P.P(1), P.bn[0].stats_f, 32)
                           : fused_eval ? bnrelu_split_epilogue(P.P(1), P.bn[0].bnc, 32, e->nsplit) : store_epilogue(P.P(1));
    ep.C = 32;
    // fp32 CUDA cores: a tcgen05 variant was measured in round 1: no faster once the CUDA-core kernel was
    // register-tiled, and it puts the 2-term-split error into the very first layer, which BatchNorm's backward amplifies
    AE_TRY(thin_gather_fwd(raw_operand(x), P.P(0), ep, fused_eval ? (float*)e->ae_pl[0] : e->y[0], batch, st));
  }
  for (int i = 0; i < 3; ++i) {
    MidLayer& m = e->enc_mid[i];
    BN& bin = P.bn[i];
    BN& bout = P.bn[i + 1];
    RowGemm r{};
    r.family = FAM_FPROP; r.g = m.g; r.g.B = batch;
    r.M = batch * m.g.Hs * m.g.Ws; r.N = m.g.Cs; r.K = 9 * m.g.Cb;
    r.Bp = (const float*)m.pk_fwd;
    r.epi = training ? bias_stats_epilogue(P.P(m.b), bout.stats_f, bout.C) : store_epilogue(P.P(m.b));
    r.epi.C = bout.C;
    r.out = e->y[i + 1]; r.splitK = 1; r.partial = nullptr;
    if (fused_eval) {
      // the operand planes were written by the previous layer's epilogue; conv2 / conv3 write the next ones; conv4 writes the
      // planes of the tensor-core dense layer (h_pl is free during an encoder pass), or stores fp32 for the CUDA-core dense
      // layer (which applies BatchNorm + ReLU while it loads)
      if (i < 2 || e->encfc_tc) {
        r.epi = bnrelu_split_epilogue(P.P(m.b), bout.bnc, bout.C, e->nsplit);
        r.out = (float*)(i < 2 ? e->ae_pl[i + 1] : e->h_pl);
      }
      r.A = split_operand(e->ae_pl[i], bin.C);
      AE_TRY(tma_rowgemm(r, m.pk_fwd, e->nsplit, st));
      continue;
    }
    const BnJob job = bn_fwd_job(P, i, batch, training);      // BatchNorm of this layer's input, finalised in the split
    // the last convolution's own BatchNorm has no split behind it (the dense layer applies it while loading): its
    // coefficient job rides on this launch (last CTA out) instead of a launch of its own
    const BnJob last = bn_fwd_job(P, 3, batch, training);
    if (i == 2 && !e->simt) { r.tail_job = &last; r.tail_counter = e->head_counter + 40; }
    AE_TRY(run_rowgemm(e, r, bnrelu_operand(e->y[i], bin.bnc, bin.C), e->ae_pl[i], true,
                       (int64_t)batch * bin.count_per_image * bin.C, m.pk_fwd, &job, st));
  }
  if (!fused_eval && e->simt) AE_TRY(run_bn_job(bn_fwd_job(P, 3, batch, training), st));
  if (fused_eval && e->encfc_tc) {   // Flatten + Linear(4096, L) on tcgen05 over conv4's planes
    int ks = 1;
    AE_TRY(dense_tc(e->h_pl, e->encfc_tc, batch, e->L, 4096, e->nsplit, e->partial, e->partial_bytes, &ks, st));
    AE_TRY(reduce_partials(e->partial, ks, (int64_t)batch * e->L, P.P(17), e->L, nullptr, e->z, st));
    if (z && z != e->z) AE_CUDA(cudaMemcpyAsync(z, e->z, (size_t)batch * e->L * 4, cudaMemcpyDeviceToDevice, st));
  } else {  // Flatten + Linear(4096, L): split-K partials, fixed-order reduce (+bias)
    RowGemm r{};
    r.family = FAM_DENSE; r.M = batch; r.N = e->L; r.K = 4096;
    r.A = bnrelu_operand(e->y[3], P.bn[3].bnc, 256);
    r.Bp = e->encfc_fwd; r.epi = store_epilogue(); r.out = nullptr;
    r.splitK = fc_split_for(e, batch); r.partial = e->partial;
    AE_TRY(simt_rowgemm(r, st));
    AE_TRY(reduce_partials(e->partial, r.splitK, (int64_t)batch * e->L, P.P(17), e->L, nullptr, e->z, st));
    if (z && z != e->z) AE_CUDA(cudaMemcpyAsync(z, e->z, (size_t)batch * e->L * 4, cudaMemcpyDeviceToDevice, st));
  }
  e->last_x = x;
  e->last_batch = batch;
  return 0;
}

// conv1 weight gradient (no data gradient needed): the last piece of the encoder backward
static int conv1_wgrad(ae_engine* e, int batch, cudaStream_t st) {
  Part& P = e->part[AE_PART_ENC];
  return thin_wgrad(bnbwd_operand(e->dzy[0], e->y[0], P.bn[0].bnc, 32), raw_operand(e->last_x), P.G(0), nullptr, e->partial,
                    e->partial_bytes, batch, st);
}

int ae_encoder_backward(ae_engine_t* e, const float* dz, int batch, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_ENC, true));
  AE_CHECK(e->last_x != nullptr && e->last_batch == batch, "ae_encoder_backward: no matching training forward");
  Part& P = e->part[AE_PART_ENC];
  const int L = e->L;
  {  // Linear(4096, L) backward
    ColGemm c{};
    c.gather = 0; c.M = batch; c.I = 4096; c.J = L;
    c.A = bnrelu_operand(e->y[3], P.bn[3].bnc, 256);
    c.B = raw_operand(dz);
    c.out = P.G(16); c.permC = 256; c.permHW = 16; c.transposed = 1;
    const bool par = !e->simt;                         // weight / bias gradient beside the data gradient
    if (par) AE_TRY(fork_side(e, st));
    AE_TRY(run_wgrad(e, c, par ? e->side : st, par));
    AE_TRY(column_sums(dz, batch, L, 0, 0, P.G(17), par ? e->side : st));
    RowGemm r{};
    r.family = FAM_DENSE; r.M = batch; r.N = 4096; r.K = L;
    r.A = raw_operand(dz); r.Bp = e->encfc_bwd;
    r.epi = relubwd_epilogue(e->y[3], P.bn[3].bnc, P.bn[3].stats_b, 256);
    r.out = e->dzy[3]; r.splitK = 1;
    AE_TRY(simt_rowgemm(r, st));
    if (par) AE_TRY(join_side(e, st));
  }
  for (int i = 2; i >= 0; --i) {
    MidLayer& m = e->enc_mid[i];
    BN& bin = P.bn[i];        // BN of this layer's input (big image)
    BN& bout = P.bn[i + 1];   // BN of this layer's output (small image)
    const int Mrows = batch * m.g.Hs * m.g.Ws;
    Geom g = m.g; g.B = batch;
    const Operand a_big = bnrelu_operand(e->y[i], bin.bnc, bin.C);                          // planes: ae_pl[i] (forward)
    const Operand dy_small = bnbwd_operand(e->dzy[i + 1], e->y[i + 1], bout.bnc, bout.C);   // planes: dy_pl (now)
    const BnJob job = bn_bwd_job(P, i + 1, batch, P.G(m.b));  // backward coefficients of the output BatchNorm (+ zero bias gradient)
    void* dyp = e->dy_pl;
    if (!e->simt) AE_TRY(tma_split_operand(dy_small, (int64_t)Mrows * m.g.Cs, dyp, e->nsplit, &job, st));
    else AE_TRY(run_bn_job(job, st));
    // the weight gradient runs beside the data gradient (both only read the dy planes)
    if (!e->simt) AE_TRY(fork_side(e, st));
    AE_TRY(run_conv_wgrad(e, g, a_big, dy_small, e->ae_pl[i], dyp, P.G(m.w), e->simt ? st : e->side));
    RowGemm r{};
    r.family = FAM_DGRAD; r.g = g; r.M = Mrows; r.N = m.g.Cb; r.K = 0;
    r.Bp = (const float*)m.pk_dgrad;
    r.epi = relubwd_epilogue(e->y[i], bin.bnc, bin.stats_b, bin.C);
    r.out = e->dzy[i]; r.splitK = 1;
    // the first layer's BatchNorm backward job (nothing is split behind it: conv1's weight gradient applies it while loading)
    const BnJob first = bn_bwd_job(P, 0, batch, P.G(1));
    if (i == 0 && !e->simt) { r.tail_job = &first; r.tail_counter = e->head_counter + 40; }
    AE_TRY(run_rowgemm(e, r, dy_small, dyp, false, 0, m.pk_dgrad, nullptr, st));
    if (!e->simt) AE_TRY(join_side(e, st));
  }
  if (e->simt) AE_TRY(run_bn_job(bn_bwd_job(P, 0, batch, P.G(1)), st));
  if (!e->defer_conv1_wgrad) AE_TRY(conv1_wgrad(e, batch, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Decoder (NB:607-635)
// ---------------------------------------------------------------------------------------------
static int decoder_forward_impl(ae_engine_t* e, const float* z, int batch, int training, float* x_hat, const float* x_target,
                                cudaStream_t st) {
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_DEC, false));
  Part& P = e->part[AE_PART_DEC];
  AE_CHECK(!training || P.running, "ae_decoder_forward: training mode needs BatchNorm running buffers");
  if (training) {
    if (!e->stats_cleared) AE_CUDA(cudaMemsetAsync(P.bn[0].stats_f, 0, (char*)(e->sse + 2) - (char*)P.bn[0].stats_f, st));
  } else {
    AE_CUDA(cudaMemsetAsync(e->sse, 0, 16, st));
  }
  const bool fused_eval = !training && !e->simt;        // see ae_encoder_forward
  if (fused_eval)
    for (int l = 0; l < 3; ++l) AE_TRY(run_bn_job(bn_fwd_job(P, l, batch, 0), st));
  {  // decoder_input: Linear(L, 4096) + Unflatten, written NHWC
    RowGemm r{};
    r.family = FAM_DENSE; r.M = batch; r.N = 4096; r.K = e->L;
    r.A = raw_operand(z); r.Bp = e->decfc_fwd; r.epi = store_epilogue(e->decfc_bias);
    r.out = e->h; r.splitK = 1;
    if (!e->simt) {       // the only reader is ConvTranspose2d(256,128) on tensor cores: write its operand planes, not fp32
      r.out = nullptr; r.epi.planes = e->h_pl; r.epi.nsplit = e->nsplit;
    }
    AE_TRY(simt_rowgemm(r, st));
  }
  for (int i = 0; i < 3; ++i) {
    MidLayer& m = e->dec_mid[i];
    BN& bout = P.bn[i];
    RowGemm r{};
    r.family = FAM_DGRAD; r.g = m.g; r.g.B = batch; r.M = batch * m.g.Hs * m.g.Ws; r.N = m.g.Cb; r.K = 0;
    Operand a = i == 0 ? raw_operand(e->h) : bnrelu_operand(e->t[i - 1], P.bn[i - 1].bnc, P.bn[i - 1].C);
    if (i == 0) a.C = m.g.Cs;
    r.Bp = (const float*)m.pk_dgrad;
    r.epi = training ? bias_stats_epilogue(P.P(m.b), bout.stats_f, bout.C) : store_epilogue(P.P(m.b));
    r.epi.C = bout.C;
    r.out = e->t[i]; r.splitK = 1;
    if (fused_eval) {
      // convT1 / convT2 emit the next layer's operand planes; convT3 stores fp32 for the 3-channel scatter kernel
      if (i < 2) { r.epi = bnrelu_split_epilogue(P.P(m.b), bout.bnc, bout.C, e->nsplit); r.out = (float*)e->ad_pl[i]; }
      r.A = split_operand(i == 0 ? e->h_pl : e->ad_pl[i - 1], m.g.Cs);
      AE_TRY(tma_rowgemm(r, m.pk_dgrad, e->nsplit, st));
      continue;
    }
    BnJob job{};
    if (i > 0) job = bn_fwd_job(P, i - 1, batch, training);
    // the last 32..256-channel layer's BatchNorm is applied by the 3-channel scatter kernel while loading: its job rides here
    const BnJob last = bn_fwd_job(P, 2, batch, training);
    if (i == 2 && !e->simt) { r.tail_job = &last; r.tail_counter = e->head_counter + 40; }
    AE_TRY(run_rowgemm(e, r, a, i == 0 ? e->h_pl : e->ad_pl[i - 1], i > 0 || e->simt, (int64_t)r.M * m.g.Cs, m.pk_dgrad,
                       i > 0 ? &job : nullptr, st));
  }
  if (!fused_eval && e->simt) AE_TRY(run_bn_job(bn_fwd_job(P, 2, batch, training), st));
  float* xo = x_hat ? x_hat : e->xhat;
  AE_TRY(thin_scatter_sigmoid_fwd(bnrelu_operand(e->t[2], P.bn[2].bnc, 32), P.P(14), P.P(15), xo, x_target, e->sse, batch, st));
  if (x_hat && training) AE_CUDA(cudaMemcpyAsync(e->xhat, x_hat, (size_t)batch * 12288 * 4, cudaMemcpyDeviceToDevice, st));
  e->last_z_dec = z;
  e->last_batch = batch;
  return 0;
}

int ae_decoder_forward(ae_engine_t* e, const float* z, int batch, int training, float* x_hat, ae_stream_t stream) {
  return decoder_forward_impl(e, z, batch, training, x_hat, nullptr, (cudaStream_t)stream);
}

// thin_up: SIGMOID_BWD operand describing d(loss)/d(pre-sigmoid)
static int decoder_backward_impl(ae_engine_t* e, const Operand& thin_up, int batch, float* dz, const float* dz_addend,
                                 cudaStream_t st) {
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_DEC, true));
  AE_CHECK(e->last_z_dec != nullptr && e->last_batch == batch, "ae_decoder_backward: no matching training forward");
  Part& P = e->part[AE_PART_DEC];
  const int L = e->L;
  // convT4 (32 -> 3)
  {
    const Operand wide = bnrelu_operand(e->t[2], P.bn[2].bnc, 32);
    const Epilogue ep = relubwd_epilogue(e->t[2], P.bn[2].bnc, P.bn[2].stats_b, 32);
    if (e->simt) {
      AE_TRY(thin_bwd_fused(wide, thin_up, P.P(14), ep, e->dzt[2], P.G(14), P.G(15), e->partial, e->partial_bytes, batch, st, 0));
    } else {
      // the reduce of the weight-gradient partials leaves the data-gradient path: it runs on the side branch, ahead of
      // the next layer's weight gradient (same stream, so the shared partial buffer is free again when that one starts)
      AE_TRY(thin_bwd_fused(wide, thin_up, P.P(14), ep, e->dzt[2], P.G(14), P.G(15), e->partial, e->partial_bytes, batch, st, 1));
      AE_TRY(fork_side(e, st));
      AE_TRY(thin_bwd_fused(wide, thin_up, P.P(14), ep, e->dzt[2], P.G(14), P.G(15), e->partial, e->partial_bytes, batch, e->side, 2));
    }
  }
  for (int i = 2; i >= 0; --i) {
    MidLayer& m = e->dec_mid[i];
    BN& bout = P.bn[i];  // BN after this layer's output (big image)
    const int Mrows = batch * m.g.Hs * m.g.Ws;
    Operand small = i == 0 ? raw_operand(e->h) : bnrelu_operand(e->t[i - 1], P.bn[i - 1].bnc, P.bn[i - 1].C);
    if (i == 0) small.C = m.g.Cs;                                                   // planes: h_pl / ad_pl[i-1] (forward)
    const Operand dy_big = bnbwd_operand(e->dzt[i], e->t[i], bout.bnc, bout.C);     // planes: dy_pl (now)
    Geom g = m.g; g.B = batch;
    const BnJob job = bn_bwd_job(P, i, batch, P.G(m.b));
    void* dyp = e->dy_pl;
    if (!e->simt) AE_TRY(tma_split_operand(dy_big, (int64_t)Mrows * 4 * m.g.Cb, dyp, e->nsplit, &job, st));
    else AE_TRY(run_bn_job(job, st));
    if (!e->simt) AE_TRY(fork_side(e, st));
    AE_TRY(run_conv_wgrad(e, g, dy_big, small, dyp, i == 0 ? e->h_pl : e->ad_pl[i - 1], P.G(m.w), e->simt ? st : e->side));
    RowGemm r{};
    r.family = FAM_FPROP; r.g = g; r.M = Mrows; r.N = m.g.Cs; r.K = 9 * m.g.Cb;
    r.Bp = (const float*)m.pk_fwd;
    if (i == 0) { r.epi = store_epilogue(); r.epi.C = m.g.Cs; r.out = e->dh; }
    else { BN& bin = P.bn[i - 1]; r.epi = relubwd_epilogue(e->t[i - 1], bin.bnc, bin.stats_b, bin.C); r.out = e->dzt[i - 1]; }
    r.splitK = 1;
    AE_TRY(run_rowgemm(e, r, dy_big, dyp, false, 0, m.pk_fwd, nullptr, st));
    if (!e->simt) AE_TRY(join_side(e, st));
  }
  {  // decoder_input backward
    ColGemm c{};
    c.gather = 0; c.M = batch; c.I = 4096; c.J = L;
    c.A = raw_operand(e->dh); c.B = raw_operand(e->last_z_dec);
    c.out = P.G(0); c.permC = 256; c.permHW = 16; c.transposed = 0;
    const bool par = !e->simt;
    if (par) AE_TRY(fork_side(e, st));
    AE_TRY(run_wgrad(e, c, par ? e->side : st, par));
    AE_TRY(column_sums(e->dh, batch, 4096, 256, 16, P.G(1), par ? e->side : st));
    RowGemm r{};
    r.family = FAM_DENSE; r.M = batch; r.N = L; r.K = 4096;
    r.A = raw_operand(e->dh); r.Bp = e->decfc_bwd; r.epi = store_epilogue(); r.out = nullptr;
    r.splitK = fc_split_for(e, batch); r.partial = e->partial;
    AE_TRY(simt_rowgemm(r, st));
    AE_TRY(reduce_partials(e->partial, r.splitK, (int64_t)batch * L, nullptr, 0, dz_addend, dz, st));
    if (par) AE_TRY(join_side(e, st));
  }
  return 0;
}

int ae_decoder_backward(ae_engine_t* e, const float* d_xhat, int batch, float* dz, ae_stream_t stream) {
  AE_CHECK(e && d_xhat && dz, "ae_decoder_backward: null argument");
  Operand up;
  up.src = d_xhat; up.src2 = e->xhat; up.bnc = nullptr; up.scalar = 0.f; up.mode = AE_OP_SIGMOID_BWD; up.C = 1;
  return decoder_backward_impl(e, up, batch, dz, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Classifier head (NB:692-696)
// ---------------------------------------------------------------------------------------------
int ae_head_forward(ae_engine_t* e, const float* z, int batch, float* logits, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_HEAD, false));
  Part& P = e->part[AE_PART_HEAD];
  RowGemm r{};
  r.family = FAM_DENSE; r.M = batch; r.N = 128; r.K = e->L;
  r.A = raw_operand(z); r.Bp = e->head_w1t; r.epi = store_epilogue(P.P(1)); r.out = e->hid_pre; r.splitK = 1;
  AE_TRY(simt_rowgemm(r, st));
  AE_TRY(head_out(e->hid_pre, P.P(2), P.P(3), logits ? logits : e->logits, batch, 128, e->NC, st));
  e->last_z_head = z;
  e->last_batch = batch;
  return 0;
}

int ae_head_backward(ae_engine_t* e, const float* d_logits, int batch, float* dz, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_TRY(check_batch(e, batch));
  AE_TRY(check_part(e, AE_PART_HEAD, true));
  AE_CHECK(e->last_z_head != nullptr && e->last_batch == batch, "ae_head_backward: no matching forward");
  Part& P = e->part[AE_PART_HEAD];
  const int L = e->L;
  AE_TRY(head_backward_small(d_logits, P.P(2), e->hid_pre, e->dhid, P.G(2), P.G(3), batch, 128, e->NC, st));
  ColGemm c{};
  c.gather = 0; c.M = batch; c.I = 128; c.J = L;
  c.A = raw_operand(e->dhid); c.B = raw_operand(e->last_z_head);
  c.out = P.G(0); c.permC = 0; c.permHW = 0; c.transposed = 0;
  AE_TRY(run_wgrad(e, c, st));
  AE_TRY(column_sums(e->dhid, batch, 128, 0, 0, P.G(1), st));
  RowGemm r{};
  r.family = FAM_DENSE; r.M = batch; r.N = L; r.K = 128;
  r.A = raw_operand(e->dhid); r.Bp = P.P(0); r.epi = store_epilogue(); r.out = dz; r.splitK = 1;
  AE_TRY(simt_rowgemm(r, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused supervised step (NB:2676-2683 without the optimizer) and evaluation (NB:2694-2714)
// ---------------------------------------------------------------------------------------------
int ae_train_step(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha, float* loss_out,
                  ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(e && x && labels && loss_out, "ae_train_step: null argument");
  AE_TRY(check_batch(e, batch));
  // every BatchNorm statistic accumulator of both parts + the squared-error accumulator: one memset for the whole step
  AE_CUDA(cudaMemsetAsync(e->stats_base, 0, e->stats_bytes, st));
  struct Cleared { bool& f; explicit Cleared(bool& b) : f(b) { f = true; } ~Cleared() { f = false; } };
  {
    Cleared guard(e->stats_cleared);
    AE_TRY(ae_encoder_forward(e, x, batch, 1, nullptr, stream));
  }
  const double numel = (double)batch * 12288.0;
  AE_TRY(check_part(e, AE_PART_HEAD, true));
  Part& H = e->part[AE_PART_HEAD];
  // classifier head: forward + cross-entropy + backward in one launch, beside the decoder forward (both only need z);
  // its reduction also assembles the loss from the decoder's squared error, so it follows the join
  const bool par = !e->simt;
  if (par) AE_TRY(fork_side(e, st));
  AE_TRY(head_fused_step(e->z, labels, H.P(0), H.P(1), H.P(2), H.P(3), e->logits, e->dz_head, H.G(0), H.G(1), H.G(2), H.G(3),
                         loss_out, e->sse, numel, alpha, e->head_partial, e->head_counter, batch, e->L, e->NC,
                         par ? e->side : st, st, 1));
  {
    Cleared guard(e->stats_cleared);
    AE_TRY(decoder_forward_impl(e, e->z, batch, 1, nullptr, x, st));
  }
  // the head's reduction (parameter gradients + the loss, which needs the decoder's squared error) stays on the side
  // branch behind the decoder forward; the decoder backward joins that branch before anything reads dz_head
  if (par) AE_TRY(fork_side(e, st));
  AE_TRY(head_fused_step(e->z, labels, H.P(0), H.P(1), H.P(2), H.P(3), e->logits, e->dz_head, H.G(0), H.G(1), H.G(2), H.G(3),
                         loss_out, e->sse, numel, alpha, e->head_partial, e->head_counter, batch, e->L, e->NC,
                         par ? e->side : st, par ? e->side : st, 2));
  Operand up;
  up.src = x; up.src2 = e->xhat; up.bnc = nullptr; up.scalar = (float)(2.0 * (double)alpha / numel);
  up.mode = AE_OP_SIGMOID_BWD; up.C = 1;
  AE_TRY(decoder_backward_impl(e, up, batch, e->dz_tot, e->dz_head, st));
  if (e->step_comm && !e->dp_single) {
    // data parallel: the decoder's and the head's gradients are complete -- exchange them beside the encoder backward
    AE_TRY(ensure_side(e));
    AE_CUDA(cudaEventRecord(e->ev_fork2, st));
    AE_CUDA(cudaStreamWaitEvent(e->side2, e->ev_fork2, 0));
    Part& D = e->part[AE_PART_DEC];
    AE_TRY(ae_dp_allreduce(e->step_comm, D.grads, D.flat_len, e->side2));
    AE_TRY(ae_dp_allreduce(e->step_comm, H.grads, H.flat_len, e->side2));
  }
  AE_TRY(ae_encoder_backward(e, e->dz_tot, batch, stream));
  if (e->step_comm && !e->dp_single) {
    Part& E = e->part[AE_PART_ENC];
    if (e->defer_conv1_wgrad) {
      // captured step: every encoder gradient except conv1.weight (still to be computed) is final -- exchange them on the
      // side branch, beside conv1's weight gradient; ae_step_graph_capture joins the branch before Adam reads them
      const int64_t c1 = E.t[0].size;
      AE_CUDA(cudaEventRecord(e->ev_fork2, st));
      AE_CUDA(cudaStreamWaitEvent(e->side2, e->ev_fork2, 0));
      AE_TRY(ae_dp_allreduce(e->step_comm, E.grads + c1, E.flat_len - c1, e->side2));
    } else {
      AE_TRY(ae_dp_allreduce(e->step_comm, E.grads, E.flat_len, st));
      AE_CUDA(cudaEventRecord(e->ev_join2, e->side2));
      AE_CUDA(cudaStreamWaitEvent(st, e->ev_join2, 0));
    }
  }
  return 0;
}

int ae_eval_step(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha, float* loss_out,
                 float* x_hat, float* logits, float* z, ae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(e && x, "ae_eval_step: null argument");
  AE_TRY(ae_encoder_forward(e, x, batch, 0, z, stream));
  AE_TRY(decoder_forward_impl(e, e->z, batch, 0, x_hat, x, st));
  AE_TRY(ae_head_forward(e, e->z, batch, logits, stream));
  if (labels && loss_out) {
    const double numel = (double)batch * 12288.0;
    AE_TRY(softmax_ce(logits ? logits : e->logits, labels, batch, e->NC, 1.f, loss_out, nullptr, nullptr, e->sse, numel,
                      alpha, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Whole-step CUDA graph
// ---------------------------------------------------------------------------------------------
struct ae_step_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

int ae_step_graph_capture(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha, float* loss_out,
                          float* flat_params, float* flat_grads, float* adam_m, float* adam_v, int64_t flat_len,
                          const ae_adam_config_t* adam, int* step_dev, ae_dp_comm_t* comm, ae_stream_t stream,
                          ae_step_graph_t** out) {
  cudaStream_t st = (cudaStream_t)stream;
  AE_CHECK(e && out && adam, "ae_step_graph_capture: null argument");
  AE_CHECK(st != nullptr, "ae_step_graph_capture: needs a non-default stream");
  AE_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  // with a communicator the step exchanges its gradients itself: sum-allreduces that together cover the flat gradient
  // buffer, every one of them on a side branch beside compute (decoder + head beside the encoder backward, the encoder
  // beside conv1's weight gradient, conv1.weight beside Adam + re-pack of everything else)
  e->step_comm = comm;
  rowgemm2_set_sm_reserve(comm ? ae_dp_max_ctas(comm) : 0);
  // tcgen05 path: conv1's weight gradient (the last 40 us of the backward pass, and the only gradient still missing) runs
  // beside Adam + weight re-pack of every other parameter (single GPU) / beside the encoder's gradient exchange (data
  // parallel); its own 864 weights are updated after the join.
  const int64_t c1 = 864;   // conv1.weight = the first tensor of the encoder part
  const bool split_tail = !e->simt && e->part[AE_PART_ENC].params == flat_params && flat_len > c1;
  e->defer_conv1_wgrad = split_tail;
  // peers attached for exactly these flat buffers (ae_dp_peers_attach): reduce-scatter + Adam + all-gather as ONE kernel over
  // NVLink peer memory instead of NCCL allreduces + Adam (AE_B200_DP_FUSED=0 keeps the NCCL form)
  const char* fused_env = getenv("AE_B200_DP_FUSED");
  const DpAttachment* peers = comm && split_tail && !(fused_env && atoi(fused_env) == 0)
                                  ? dp_find_attachment(comm, flat_params, flat_grads, flat_len) : nullptr;
  e->dp_single = peers != nullptr;
  int rc = ae_train_step(e, x, labels, batch, alpha, loss_out, stream);
  e->step_comm = nullptr;
  e->defer_conv1_wgrad = false;
  e->dp_single = false;
  rowgemm2_set_sm_reserve(0);
  float gscale = 1.f;
  if (rc == 0 && comm) gscale = 1.f / (float)ae_dp_world(comm);
  auto adam_range = [&](int64_t lo, int64_t n, int bump) {
    return adam_step_flat_range(flat_params + lo, flat_grads + lo, adam_m + lo, adam_v + lo, n, adam->lr, adam->beta1, adam->beta2,
                                adam->eps, adam->weight_decay, gscale, step_dev, bump, st);
  };
  auto link = [&](cudaEvent_t ev, cudaStream_t from, cudaStream_t to) {       // `to` continues after everything on `from`
    if (cudaEventRecord(ev, from) != cudaSuccess || cudaStreamWaitEvent(to, ev, 0) != cudaSuccess) {
      set_error("ae_step_graph_capture: event record / wait failed while capturing the data-parallel tail");
      return 1;
    }
    return 0;
  };
  if (rc == 0 && peers) {
    // ONE exchange round after the last gradient (conv1's weights).  Measured on 2 and 8 GPUs: running conv1's weight gradient
    // beside a first round over everything else and its 864 weights in a round of their own is no faster (0.718 / 0.749 ms
    // against 0.713 / 0.747 ms): what the ranks wait for is each other, not the exchange.
    rc = conv1_wgrad(e, batch, st);
    if (rc == 0)
      rc = dp_adam_fused(comm, peers, adam_m, adam_v, 0, flat_len, adam->lr, adam->beta1, adam->beta2, adam->eps, adam->weight_decay,
                         step_dev, 1, st);
    if (rc == 0) rc = pack_all_parts(e, st);
  } else if (rc == 0 && split_tail && comm) {
    rc = conv1_wgrad(e, batch, st);
    // every exchange issued so far (decoder, head, encoder without conv1.weight) is complete before Adam reads it
    if (rc == 0) rc = link(e->ev_join2, e->side2, st);
    // conv1.weight's gradient travels beside Adam + re-pack of all other parameters
    if (rc == 0) rc = link(e->ev_fork2, st, e->side2);
    if (rc == 0) rc = ae_dp_allreduce(comm, flat_grads, c1, e->side2);
    if (rc == 0) rc = adam_range(c1, flat_len - c1, 0);
    if (rc == 0) rc = pack_all_parts(e, st);
    if (rc == 0) rc = link(e->ev_join2, e->side2, st);
    if (rc == 0) rc = adam_range(0, c1, 1);
  } else if (rc == 0 && split_tail) {
    rc = fork_side(e, st);
    if (rc == 0) rc = conv1_wgrad(e, batch, e->side);
    if (rc == 0) rc = adam_range(c1, flat_len - c1, 0);
    if (rc == 0) rc = pack_all_parts(e, st);
    if (rc == 0) rc = join_side(e, st);
    if (rc == 0) rc = adam_range(0, c1, 1);
  } else {
    if (rc == 0)
      rc = adam_step_flat(flat_params, flat_grads, adam_m, adam_v, flat_len, adam->lr, adam->beta1, adam->beta2, adam->eps,
                          adam->weight_decay, gscale, step_dev, st);
    if (rc == 0) {
      if (e->simt) { for (int p = 0; rc == 0 && p < AE_NUM_PARTS; ++p) rc = ae_engine_pack_weights(e, p, stream); }
      else rc = pack_all_parts(e, st);
    }
  }
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) { set_error("ae_step_graph_capture: end capture failed: %s", cudaGetErrorString(ce)); return 1; }
  ae_step_graph* g = new ae_step_graph();
  g->graph = graph;
  ce = cudaGraphInstantiate(&g->exec, graph, 0);
  if (ce != cudaSuccess) {
    set_error("ae_step_graph_capture: instantiate failed: %s", cudaGetErrorString(ce));
    cudaGraphDestroy(graph);
    delete g;
    return 1;
  }
  *out = g;
  return 0;
}

int ae_step_graph_launch(ae_step_graph_t* g, ae_stream_t stream) {
  AE_CHECK(g && g->exec, "ae_step_graph_launch: null graph");
  AE_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return 0;
}

int ae_step_graph_num_kernels(const ae_step_graph_t* g) {
  if (!g || !g->graph) return -1;
  size_t n = 0;
  if (cudaGraphGetNodes(g->graph, nullptr, &n) != cudaSuccess) return -1;
  std::vector<cudaGraphNode_t> nodes(n);
  if (n && cudaGraphGetNodes(g->graph, nodes.data(), &n) != cudaSuccess) return -1;
  int k = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++k;
  }
  return k;
}

void ae_step_graph_destroy(ae_step_graph_t* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

}  // extern "C"

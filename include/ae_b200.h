/*
 * ae_b200.h -- C ABI of the B200-native supervised-autoencoder + MLP hot path.
 *
 * The reference (a PyTorch notebook) has no FFI of its own: its hot path sits behind the
 * torch nn.Module / torch.optim API.  These entry points are what a
 * host binding for that path calls; each one cites the reference lines it replaces
 * (NB:n = line n of Code/Hybrid_autoencoder-MLP_pipeline_for_satellite_image_classification.ipynb).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates user-visible memory: parameters, gradients, optimizer
 *     state, BatchNorm buffers and the workspace are caller-owned (torch tensors in the
 *     Python host layer); engine-private scratch lives inside the caller's workspace;
 *   - every function takes the cudaStream_t to launch on (as void*), is asynchronous, and
 *     returns 0 on success; on failure it returns non-zero and ae_last_error() describes it;
 *   - one host thread per engine; there is no CPU fallback anywhere in the library.
 *
 * Tensor layouts at the boundary are the reference's: images NCHW fp32, weights in torch
 * layout.  Internally activations are NHWC.  "big"/"small" name the two images of a
 * stride-2 3x3 layer: big = H x W x Cb, small = H/2 x W/2 x Cs.  Both torch.nn.Conv2d
 * (Cb -> Cs) and torch.nn.ConvTranspose2d (Cs -> Cb) store their weight as [Cs, Cb, 3, 3].
 */
#ifndef AE_B200_H
#define AE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AE_ABI_VERSION 1

typedef void* ae_stream_t; /* cudaStream_t */

/* arithmetic mode of the GEMM-shaped kernels */
enum {
  AE_PREC_FP32 = 0, /* operands as 2-term bf16 split (hi + lo planes), 3 tcgen05 MMAs per k-step, fp32 accumulate: rel <= 1e-4 */
  AE_PREC_BF16 = 1  /* operands rounded to bf16 (one plane), fp32 accumulate: rel <= 1e-2 */
};
enum {
  AE_BACKEND_TC = 0,  /* tcgen05 / TMEM kernels */
  AE_BACKEND_SIMT = 1 /* fp32 CUDA-core kernels (bring-up / cross-check path, still GPU) */
};

/* operand transforms applied while a kernel loads an activation operand */
enum {
  AE_OP_RAW = 0,        /* value = src */
  AE_OP_BNRELU = 1,     /* value = relu(src * scale[c] + shift[c])                (BatchNorm + ReLU forward) */
  AE_OP_BNBWD = 2,      /* value = A[c] * src + B[c] * (src2 - mean[c]) + C[c]    (BatchNorm backward apply) */
  AE_OP_SIGMOID_BWD = 3, /* value = up * s * (1 - s), s = src2 (sigmoid output): up = src, or scalar * (s - src) when fused MSE */
  AE_OP_SPLIT_BF16 = 4   /* src = split-bf16 planes written by ae_split_operand (the only operand form the tcgen05 GEMMs take) */
};

/* per-BatchNorm-layer coefficient block: 8 rows of C floats */
enum {
  AE_BNC_SCALE = 0, AE_BNC_SHIFT = 1, AE_BNC_MEAN = 2, AE_BNC_RSTD = 3,
  AE_BNC_A = 4, AE_BNC_B = 5, AE_BNC_C = 6, AE_BNC_ROWS = 8
};

typedef struct ae_operand {
  const float* src;
  const float* src2;
  const float* bnc;  /* coefficient block [AE_BNC_ROWS][C] or NULL */
  float scalar;
  int mode;          /* AE_OP_* */
} ae_operand_t;

/* epilogues of the GEMM-shaped kernels */
enum {
  AE_EPI_STORE = 0,          /* out = acc (+ bias) */
  AE_EPI_BIAS_STATS = 1,     /* out = acc + bias; stats[c] += sum(out), stats[C+c] += sum(out^2)  (feeds BatchNorm) */
  AE_EPI_RELUBWD_STATS = 2,  /* out = acc * (scale*y+shift > 0); stats[c] += sum(out), stats[C+c] += sum(out * xhat) */
  AE_EPI_BNRELU_SPLIT = 3    /* eval-mode fusion: `out` receives split-bf16 planes of relu(scale*(acc+bias)+shift) (bnc = the
                                layer's coefficient block from running statistics): the next GEMM's operand, no fp32 pass */
};

typedef struct ae_epilogue {
  int mode;           /* AE_EPI_* */
  const float* bias;  /* [N] or NULL */
  const float* y;     /* RELUBWD: raw forward output of the layer whose BN+ReLU is being differentiated */
  const float* bnc;   /* RELUBWD: that layer's coefficient block */
  double* stats;      /* [2*C] fp64 accumulators (atomically added to), or NULL */
} ae_epilogue_t;

typedef struct ae_conv_geom {
  int batch;
  int hs, ws; /* small image height / width (big image is 2*hs x 2*ws); powers of two */
  int cb, cs; /* channels of the big / small image; multiples of 8 for the GEMM kernels */
} ae_conv_geom_t;

const char* ae_last_error(void);
int ae_abi_version(void);
/* 1 if the device `ordinal` can run the library (compute capability 10.x), else 0 */
int ae_device_supported(int ordinal);

/* ------------------------------------------------------------------------------------------
 * Weight packing.  Replaces nothing in the reference (torch/cuDNN re-layout weights internally);
 * needed because the GEMM kernels read weights as K-major tiles.
 * `w` is [Cs, Cb, 3, 3] fp32 (Conv2d NB:504-516 or ConvTranspose2d NB:616-628).
 * fwd packing  : B operand of the big->small GEMM, K = (ky,kx,cb), N = cs.
 * dgrad packing: B operand of the four small->big phase GEMMs, K = (tap,cs), N = cb.
 * ---------------------------------------------------------------------------------------- */
size_t ae_packed_weight_bytes(int cs, int cb, int precision, int backend);
int ae_pack_conv_weight(const float* w, int cs, int cb, void* packed_fwd, void* packed_dgrad,
                        int precision, int backend, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ae_split_operand -- materialise an activation operand for the tcgen05 GEMMs (AE_BACKEND_TC).
 *   Reads `count` fp32 NHWC values (channel = index % channels), applies the operand transform (RAW, BNRELU = the
 *   BatchNorm2d + ReLU of NB:505-517 / NB:617-625, BNBWD = their backward) and writes split-bf16 planes
 *   [nsplit][count] bf16: plane 0 = rn_bf16(v), plane 1 (AE_PREC_FP32 only) = rn_bf16(v - plane 0).
 *   The GEMM kernels load these planes with TMA; pass them as an operand with mode AE_OP_SPLIT_BF16.
 * ---------------------------------------------------------------------------------------- */
size_t ae_split_operand_bytes(int64_t count, int precision);
int ae_split_operand(const ae_operand_t* op, int channels, int64_t count, void* planes, int precision, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ae_conv2d_s2_fwd -- big -> small 3x3 stride-2 pad-1 gather GEMM.
 *   Conv2d forward (NB:508,512,516) and ConvTranspose2d data-gradient (autograd of NB:616-624).
 *   out[n,oy,ox,cs] = sum_{ky,kx,cb} big(n,2oy-1+ky,2ox-1+kx,cb) * w[cs,cb,ky,kx]   (NHWC)
 * ae_conv2d_s2_dgrad -- small -> big scatter GEMM (4 output-parity phases).
 *   ConvTranspose2d forward (NB:616,620,624) and Conv2d data-gradient (autograd of NB:508-516).
 *   out[n,2iy-1+ky,2ix-1+kx,cb] += small(n,iy,ix,cs) * w[cs,cb,ky,kx]
 * ae_conv2d_s2_wgrad -- weight gradient of either layer type (autograd of NB:508-516, 616-624).
 *   dw[cs,cb,ky,kx] = sum_{n,oy,ox} small(n,oy,ox,cs) * big(n,2oy-1+ky,2ox-1+kx,cb)
 *   Deterministic: split-K partial tiles go to `partials`, a fixed-order reduce writes dw.
 * ---------------------------------------------------------------------------------------- */
int ae_conv2d_s2_fwd(const ae_conv_geom_t* g, const ae_operand_t* big, const void* packed_fwd,
                     const ae_epilogue_t* epi, float* out_small, int precision, int backend,
                     ae_stream_t stream);
int ae_conv2d_s2_dgrad(const ae_conv_geom_t* g, const ae_operand_t* small, const void* packed_dgrad,
                       const ae_epilogue_t* epi, float* out_big, int precision, int backend,
                       ae_stream_t stream);
size_t ae_conv2d_s2_wgrad_workspace_bytes(const ae_conv_geom_t* g, int precision, int backend);
int ae_conv2d_s2_wgrad(const ae_conv_geom_t* g, const ae_operand_t* big, const ae_operand_t* small,
                       float* dw, void* partials, size_t partials_bytes, int precision, int backend,
                       ae_stream_t stream);

/* torch.optim.Adam hyper-parameters (NB:2654, NB:3461) */
typedef struct ae_adam_config {
  float lr, beta1, beta2, eps, weight_decay;
} ae_adam_config_t;

/* ------------------------------------------------------------------------------------------
 * Thin (3-channel) layers: Conv2d(3,32) NB:504 and ConvTranspose2d(32,3)+Sigmoid NB:628-629.
 * thin = [B,3,64,64] NCHW fp32 (the reference's image layout), wide = [B,32,32,32] NHWC.
 * ---------------------------------------------------------------------------------------- */
/* K = 27 / N = 3 do not fill a tensor-core tile: both backends run the fp32 CUDA-core kernels of thin.cu (a tcgen05 variant
 * was measured in round 1 and was no faster).  `precision` only selects the number of bf16 planes an
 * AE_EPI_BNRELU_SPLIT epilogue writes. */
int ae_thin_gather_fwd(const ae_operand_t* thin, const float* w /*[32,3,3,3]*/, const ae_epilogue_t* epi,
                       float* out_wide, int batch, int precision, int backend, ae_stream_t stream);
/* x_hat = sigmoid(convT(wide) + bias); if x != NULL also accumulates sum((x_hat-x)^2) into *sse (fp64) */
int ae_thin_scatter_sigmoid_fwd(const ae_operand_t* wide, const float* w /*[32,3,3,3]*/, const float* bias /*[3]*/,
                                float* x_hat, const float* x, double* sse, int batch, int precision, int backend,
                                ae_stream_t stream);
int ae_thin_wgrad(const ae_operand_t* wide, const ae_operand_t* thin, float* dw /*[32,3,3,3]*/,
                  float* dbias_thin /*[3] or NULL*/, void* partials, size_t partials_bytes, int batch,
                  int precision, int backend, ae_stream_t stream);
size_t ae_thin_wgrad_workspace_bytes(int batch);
/* ae_thin_gather_fwd (data gradient, `epi` = RELUBWD) and ae_thin_wgrad of ConvTranspose2d(32,3) in one pass over the
 * thin operand (autograd of NB:628-629); same workspace as ae_thin_wgrad */
int ae_thin_bwd_fused(const ae_operand_t* wide, const ae_operand_t* thin, const float* w /*[32,3,3,3]*/,
                      const ae_epilogue_t* epi, float* out_wide, float* dw, float* dbias_thin, void* partials,
                      size_t partials_bytes, int batch, int precision, int backend, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm statistics (NB:505-517, NB:617-625; torch defaults eps 1e-5, momentum 0.1).
 * ae_bn_finalize: from fp64 sums (sum, sum of squares) over `count` values per channel compute
 *   mean, biased var, rstd, scale = gamma*rstd, shift = beta - mean*scale into `bnc`, and in
 *   training update running_mean / running_var (unbiased) in place.  training == 0: scale/shift
 *   from the running statistics (stats may be NULL).
 * ae_bn_bwd_reduce: from fp64 sums (sum dz, sum dz*xhat) write dgamma, dbeta and the backward
 *   apply coefficients A, B, C into `bnc` (dy = A*dz + B*(y - mean) + C).
 * ---------------------------------------------------------------------------------------- */
int ae_bn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float* bnc, int channels, int training,
                   ae_stream_t stream);
int ae_bn_bwd_reduce(const double* stats, int64_t count, const float* gamma, float* bnc,
                     float* dgamma, float* dbeta, int channels, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense layers.  ae_linear_fwd: out[m,n] = sum_k A(m,k) * w[n,k] + bias[n]   (nn.Linear, NB:521, NB:611)
 * ae_linear_bwd: dA[m,k] = sum_n dOut[m,n] * w[n,k];  dw[n,k] = sum_m dOut[m,n] * A(m,k);  db[n] = sum_m dOut[m,n]
 * `perm_hw` != 0: the K (fwd: input, 4096 = C*hw) axis of `w` is in the reference's (C,H,W) flatten
 * order (NB:520 / NB:614) while the activation is NHWC; the kernels permute on the fly.
 * ---------------------------------------------------------------------------------------- */
int ae_linear_fwd(const ae_operand_t* a, int a_channels, const float* w, const float* bias, float* out,
                  int m, int n, int k, int perm_k_hw, int perm_n_hw, void* workspace, size_t workspace_bytes,
                  ae_stream_t stream);
int ae_linear_bwd(const ae_operand_t* a, int a_channels, const float* w, const float* d_out,
                  float* d_a, const ae_epilogue_t* d_a_epi, int d_a_channels,
                  float* dw, float* db, int m, int n, int k, int perm_k_hw, int perm_n_hw,
                  void* workspace, size_t workspace_bytes, ae_stream_t stream);
size_t ae_linear_workspace_bytes(int m, int n, int k);

/* ------------------------------------------------------------------------------------------
 * Losses.  ae_softmax_ce_fwd_bwd: nn.CrossEntropyLoss() (NB:2653, NB:3463) mean over the batch;
 *   writes loss (fp32 scalar) and, if d_logits != NULL, (softmax - onehot) * grad_scale / batch.
 * ae_sigmoid_mse_fwd_bwd: nn.MSELoss() (NB:2652) of an already-sigmoided x_hat against x, plus
 *   d(pre-sigmoid) = scale * 2 (x_hat - x) / numel * x_hat (1 - x_hat).
 * ---------------------------------------------------------------------------------------- */
int ae_softmax_ce_fwd_bwd(const float* logits, const int64_t* labels, int batch, int classes,
                          float grad_scale, float* loss, float* d_logits, int* correct, ae_stream_t stream);
/* `loss` must point to 4 floats (8-byte aligned): [0] = loss, [2..3] = fp64 scratch */
int ae_sigmoid_mse_fwd_bwd(const float* x_hat, const float* x, int64_t numel, float scale, float* loss,
                           float* d_pre, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ae_mlp_fwd_bwd_ce -- the whole 64->128->64->10 classifier (NB:2970-2987) forward, softmax
 * cross-entropy and backward in ONE persistent kernel (one thread-block cluster, weights resident
 * in shared memory, BatchNorm1d batch statistics reduced through distributed shared memory).
 * Parameters are the flat fp32 buffer in torch parameters() order (ae_mlp_param_layout).
 *   training != 0: batch statistics, running-stat update, dropout (keep mask given, or Philox from seed),
 *                  gradients written to `grads` (same flat layout).
 *   training == 0: running statistics, no dropout, no gradients (NB:3499, NB:3702).
 * ---------------------------------------------------------------------------------------- */
int64_t ae_mlp_param_layout(int input_dim, int num_classes, int64_t* offsets /*[10]*/, int64_t* sizes /*[10]*/);
int ae_mlp_fwd_bwd_ce(const float* params, float* grads, float* bn_running /*[mean1,var1,mean2,var2]*/,
                      const float* x, const int64_t* labels, const uint8_t* dropout_keep /*[B,128] or NULL*/,
                      uint64_t dropout_seed, float dropout_p, int batch, int input_dim, int num_classes,
                      int training, float* logits, float* loss /*[1]*/, int* correct /*[1]*/,
                      void* workspace, size_t workspace_bytes, ae_stream_t stream);
size_t ae_mlp_workspace_bytes(int batch, int input_dim, int num_classes);
/* One MLP training step of the reference loop (NB:3476-3482) -- forward + CE + backward launch, then the fused Adam launch
 * (torch.optim.Adam semantics, coupled weight decay) -- with nothing baked in that changes from step to step: the dropout
 * seed of a launch is dropout_seed + seed_dev[0] and the launch advances seed_dev[0]; num_batches_tracked (bn_steps,
 * int64[2]) is bumped on the device.  Meant to be captured in a CUDA graph and replayed per batch. */
int ae_mlp_train_step(float* params, float* grads, float* bn_running, int64_t* bn_steps, const float* x, const int64_t* labels,
                      uint64_t dropout_seed, uint64_t* seed_dev, float dropout_p, int batch, int input_dim, int num_classes,
                      float* logits, float* loss, int* correct, void* workspace, size_t workspace_bytes,
                      const ae_adam_config_t* adam, float* adam_m, float* adam_v, int* step_dev, ae_stream_t stream);
/* The same step with its batch read through an index (NB:3443 DataLoader(shuffle=True) + NB:3476-3482): row r of the batch is
 * row order[cursor[0] + r] of x_all / labels_all; cursor = int64[2] {rows consumed, steps done} advances on the device and
 * (loss, correct) of step i are stored in hist[i][0..1].  An epoch is then one graph replay per batch, nothing else. */
int ae_mlp_train_step_indexed(float* params, float* grads, float* bn_running, int64_t* bn_steps, const float* x_all,
                              const int64_t* labels_all, const int64_t* order, int64_t* cursor, float* hist, uint64_t dropout_seed,
                              uint64_t* seed_dev, float dropout_p, int batch, int input_dim, int num_classes, float* logits,
                              float* loss, int* correct, void* workspace, size_t workspace_bytes, const ae_adam_config_t* adam,
                              float* adam_m, float* adam_v, int* step_dev, ae_stream_t stream);
int ae_mlp_forward_eval(const float* params, const float* bn_running, const float* x, int batch, int input_dim,
                        int num_classes, float* logits, int64_t* argmax /*or NULL*/, ae_stream_t stream);
/* Backward of a preceding training-mode ae_mlp_fwd_bwd_ce call made with labels == NULL (forward only,
 * same workspace), given d(loss)/d(logits) computed by the caller (torch's CrossEntropyLoss backward in
 * the drop-in loop NB:3479-3481). */
int ae_mlp_backward(const float* params, float* grads, const float* x, const float* d_logits, float dropout_p,
                    int batch, int input_dim, int num_classes, void* workspace, size_t workspace_bytes,
                    ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ae_adam_step_flat -- torch.optim.Adam.step() (NB:2654/2684, NB:3461/3482) over one flat buffer
 * with 128-bit loads/stores.  g is multiplied by grad_scale first (1/world for data parallel),
 * weight_decay is the coupled L2 of torch.optim.Adam (g += wd * p).  `step_dev` holds the step
 * count on the device (int32, incremented by this call) so the launch is CUDA-graph replayable.
 * ---------------------------------------------------------------------------------------- */
int ae_adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, float grad_scale, int* step_dev,
                      ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ae_augment_u8 -- the reference's input transforms on the device:
 *   training  NB:386-391  RandomHorizontalFlip -> RandomCrop(64, padding=pad) -> ToTensor -> AddGaussianNoise (NB:361-368)
 *   eval      NB:393-395  ToTensor                      (flip = off_y = off_x = noise = NULL, noise_std = 0)
 * images: [num_images,64,64,3] uint8 HWC resident on the device; index[b] (or b) selects the source of output b
 * (the shuffled batch of the DataLoader, NB:420).  flip[b] != 0 mirrors the image first; (off_y[b], off_x[b]) in
 * [0, 2*pad] are RandomCrop's (i, j) into the zero-padded image (pad, pad = no shift).  out: [batch,3,64,64] fp32 =
 * value/255 + (n*noise_std + noise_mean), n from `noise` ([batch,3,64,64] fp32) if given, else -- when noise_std != 0 --
 * standard normals drawn from Philox4x32-10 keyed by (seed, element index).
 * ---------------------------------------------------------------------------------------- */
int ae_augment_u8(const uint8_t* images, int64_t num_images, const int64_t* index, const uint8_t* flip,
                  const int32_t* off_y, const int32_t* off_x, int pad, const float* noise, uint64_t seed,
                  float noise_mean, float noise_std, float* out, int batch, ae_stream_t stream);

/* layout helpers at the boundary (NCHW fp32 <-> NHWC fp32 / bf16) */
int ae_layout_nchw_f32_to_nhwc_f32(const float* src, float* dst, int n, int c, int h, int w, ae_stream_t stream);
int ae_layout_nhwc_f32_to_nchw_f32(const float* src, float* dst, int n, int c, int h, int w, ae_stream_t stream);
int ae_layout_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int n, int c, int h, int w, ae_stream_t stream);
int ae_layout_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int n, int c, int h, int w, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Engine: the SupervisedAutoencoder (NB:685-702) as three parts -- encoder (NB:499-525),
 * decoder (NB:607-635), classifier head (NB:692-696) -- over caller-owned flat buffers.
 * ---------------------------------------------------------------------------------------- */
typedef struct ae_engine ae_engine_t;
enum { AE_PART_ENC = 0, AE_PART_DEC = 1, AE_PART_HEAD = 2, AE_NUM_PARTS = 3 };

typedef struct ae_engine_config {
  int latent_dim;
  int num_classes;
  int max_batch;
  int precision; /* AE_PREC_* */
  int backend;   /* AE_BACKEND_* */
} ae_engine_config_t;

int ae_engine_create(const ae_engine_config_t* cfg, ae_engine_t** out);
void ae_engine_destroy(ae_engine_t* e);
/* Parameter tensors of `part` in torch parameters() order: returns their count, fills offsets/sizes
 * (fp32 elements, offsets 4-aligned) and *flat_len with the padded flat length of the part. */
int ae_engine_param_layout(const ae_engine_t* e, int part, int64_t* offsets, int64_t* sizes, int64_t* flat_len);
/* BatchNorm layers of `part`: returns their count, fills channels[]; the running buffer of a part is
 * [mean_0, var_0, mean_1, var_1, ...] fp32, the step buffer one int64 (num_batches_tracked) per layer. */
int ae_engine_bn_layout(const ae_engine_t* e, int part, int* channels);
size_t ae_engine_workspace_bytes(const ae_engine_t* e);
int ae_engine_bind_workspace(ae_engine_t* e, void* workspace, size_t bytes);
int ae_engine_bind_part(ae_engine_t* e, int part, float* params, float* grads, float* bn_running,
                        int64_t* bn_steps);
/* re-derive the packed weights after the parameters changed (optimizer step, load_state_dict) */
int ae_engine_pack_weights(ae_engine_t* e, int part, ae_stream_t stream);

/* stage-level calls used by the autograd shells.  Backward calls WRITE the part's gradients. */
int ae_encoder_forward(ae_engine_t* e, const float* x, int batch, int training, float* z, ae_stream_t stream);
int ae_decoder_forward(ae_engine_t* e, const float* z, int batch, int training, float* x_hat, ae_stream_t stream);
int ae_head_forward(ae_engine_t* e, const float* z, int batch, float* logits, ae_stream_t stream);
int ae_decoder_backward(ae_engine_t* e, const float* d_xhat, int batch, float* dz, ae_stream_t stream);
int ae_head_backward(ae_engine_t* e, const float* d_logits, int batch, float* dz, ae_stream_t stream);
int ae_encoder_backward(ae_engine_t* e, const float* dz, int batch, ae_stream_t stream);

/* One supervised step of NB:2676-2683 without the optimizer: forward (training mode), loss =
 * alpha*MSE + CE, backward.  Gradients land in the bound flat gradient buffers, losses in
 * loss_out = {loss, mse, ce} (device).  x: [B,3,64,64] fp32 NCHW, labels int64. */
int ae_train_step(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha,
                  float* loss_out, ae_stream_t stream);
/* NB:2694-2714: eval-mode forward + the same loss; optional outputs may be NULL */
int ae_eval_step(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha,
                 float* loss_out, float* x_hat, float* logits, float* z, ae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole-step CUDA graph: [train step] -> [NCCL allreduce] -> [Adam] -> [weight re-pack]
 * captured once, replayed per step (x / labels / loss are fixed device addresses).
 * ---------------------------------------------------------------------------------------- */
typedef struct ae_step_graph ae_step_graph_t;
typedef struct ae_dp_comm ae_dp_comm_t;
int ae_step_graph_capture(ae_engine_t* e, const float* x, const int64_t* labels, int batch, float alpha,
                          float* loss_out, float* flat_params, float* flat_grads, float* adam_m,
                          float* adam_v, int64_t flat_len, const ae_adam_config_t* adam, int* step_dev,
                          ae_dp_comm_t* comm /*or NULL*/, ae_stream_t stream, ae_step_graph_t** out);
int ae_step_graph_launch(ae_step_graph_t* g, ae_stream_t stream);
/* number of kernel nodes in the captured step (the library's own kernels + NCCL's, if any) */
int ae_step_graph_num_kernels(const ae_step_graph_t* g);
void ae_step_graph_destroy(ae_step_graph_t* g);

/* ------------------------------------------------------------------------------------------
 * Data parallel (new capability; the reference is single-device, NB:277): one NCCL communicator
 * per process, one sum-allreduce of the flat fp32 gradient buffer per step.
 * ---------------------------------------------------------------------------------------- */
#define AE_DP_UNIQUE_ID_BYTES 128
int ae_dp_get_unique_id(uint8_t* id_host /*[AE_DP_UNIQUE_ID_BYTES]*/);
int ae_dp_init(const uint8_t* id_host, int rank, int world, ae_dp_comm_t** out);
int ae_dp_allreduce(ae_dp_comm_t* c, float* buf, int64_t n, ae_stream_t stream);
/* Fused exchange over NVLink peer memory (one process per GPU of ONE box): every rank exports the allocations that hold its
 * flat parameter buffer, flat gradient buffer and a zeroed flag block of AE_DP_FLAG_BYTES, the host all-gathers the handles
 * (torch.distributed) and attaches them.  A step captured with this communicator for exactly these flat buffers then replaces
 * "allreduce + Adam" by ONE kernel: reduce-scatter out of the peers' gradient buffers, Adam on the own 1/world shard (only
 * that shard of the moments is maintained), all-gather of the new parameters into every rank's buffer. */
#define AE_DP_IPC_HANDLE_BYTES 64
#define AE_DP_FLAG_BYTES 128
int ae_dp_ipc_export(const void* dev_ptr, uint8_t* handle /*[AE_DP_IPC_HANDLE_BYTES]*/, int64_t* offset);
int ae_dp_peers_attach(ae_dp_comm_t* c, const uint8_t* handles /*[world][3][AE_DP_IPC_HANDLE_BYTES]: params, grads, flags*/,
                       const int64_t* offsets /*[world][3]*/, float* own_params, float* own_grads, void* own_flags,
                       int64_t flat_len);
int ae_dp_world(const ae_dp_comm_t* c);
/* upper bound of the SMs one collective of this communicator occupies (ncclConfig_t.maxCTAs; AE_B200_NCCL_MAX_CTAS) */
int ae_dp_max_ctas(const ae_dp_comm_t* c);
void ae_dp_destroy(ae_dp_comm_t* c);

#ifdef __cplusplus
}
#endif
#endif /* AE_B200_H */

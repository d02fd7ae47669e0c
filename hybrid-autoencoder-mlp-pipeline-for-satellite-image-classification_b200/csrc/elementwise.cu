// BatchNorm statistic finalisation, losses, fused flat Adam and boundary layout conversions.
#include "common.cuh"

namespace ae {

static constexpr double BN_EPS = 1e-5;
static constexpr double BN_MOMENTUM = 0.1;

// ---------------------------------------------------------------------------------------------
// BatchNorm (NB:505-517, 617-625, 2974-2980): eps 1e-5, momentum 0.1, biased var to normalise,
// unbiased var into running_var.
// ---------------------------------------------------------------------------------------------
__global__ void k_bn_finalize(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                              float* __restrict__ bnc, int C, int training, int64_t* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c == 0 && training && nbt) *nbt += 1;
  double mean, var;
  if (training) {
    mean = stats[c] / count;
    var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (rmean) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      rmean[c] = (float)((1.0 - BN_MOMENTUM) * (double)rmean[c] + BN_MOMENTUM * mean);
      rvar[c] = (float)((1.0 - BN_MOMENTUM) * (double)rvar[c] + BN_MOMENTUM * unb);
    }
  } else {
    mean = (double)rmean[c];
    var = (double)rvar[c];
  }
  const float rstd = (float)(1.0 / sqrt(var + BN_EPS));
  const float scale = gamma[c] * rstd;
  bnc[AE_BNC_SCALE * C + c] = scale;
  bnc[AE_BNC_SHIFT * C + c] = beta[c] - (float)mean * scale;
  bnc[AE_BNC_MEAN * C + c] = (float)mean;
  bnc[AE_BNC_RSTD * C + c] = rstd;
}

int bn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta, float* rmean,
                float* rvar, float* bnc, int C, int training, cudaStream_t st) {
  AE_CHECK(training || (rmean && rvar), "bn_finalize: eval mode needs running statistics");
  AE_CHECK(!training || stats, "bn_finalize: training mode needs batch statistics");
  k_bn_finalize<<<(C + 127) / 128, 128, 0, st>>>(stats, (double)count, gamma, beta, rmean, rvar, bnc, C, training, nullptr);
  AE_LAUNCH_CHECK();
  return 0;
}

int bn_bwd_reduce(const double* stats, int64_t count, const float* gamma, float* bnc, float* dgamma, float* dbeta,
                  int C, cudaStream_t st);

__global__ void k_bn_bwd_reduce(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                float* __restrict__ bnc, float* __restrict__ dgamma, float* __restrict__ dbeta, int C,
                                float* __restrict__ dzero);

int run_bn_job(const BnJob& j, cudaStream_t st) {
  if (j.kind == BN_JOB_FINALIZE) {
    AE_CHECK(j.training || (j.rmean && j.rvar), "bn job: eval mode needs running statistics");
    k_bn_finalize<<<(j.C + 127) / 128, 128, 0, st>>>(j.stats, j.count, j.gamma, j.beta, j.rmean, j.rvar, j.bnc, j.C, j.training, j.nbt);
    AE_LAUNCH_CHECK();
  } else if (j.kind == BN_JOB_BWD) {
    k_bn_bwd_reduce<<<(j.C + 127) / 128, 128, 0, st>>>(j.stats, j.count, j.gamma, j.bnc, j.dgamma, j.dbeta, j.C, j.dzero);
    AE_LAUNCH_CHECK();
  }
  return 0;
}

// dy = A*dz + B*(y - mean) + C with  A = gamma*rstd,  B = -A*rstd*S2/M,  C = -A*S1/M
__global__ void k_bn_bwd_reduce(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                float* __restrict__ bnc, float* __restrict__ dgamma, float* __restrict__ dbeta, int C,
                                float* __restrict__ dzero) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dzero) dzero[c] = 0.f;
  const double s1 = stats[c], s2 = stats[C + c];
  const double rstd = (double)bnc[AE_BNC_RSTD * C + c];
  const double a = (double)gamma[c] * rstd;
  const double b = -a * rstd * s2 / count;
  const double k = -a * s1 / count;
  bnc[AE_BNC_A * C + c] = (float)a;
  bnc[AE_BNC_B * C + c] = (float)b;
  bnc[AE_BNC_C * C + c] = (float)k;
  if (dgamma) dgamma[c] = (float)s2;
  if (dbeta) dbeta[c] = (float)s1;
}

int bn_bwd_reduce(const double* stats, int64_t count, const float* gamma, float* bnc, float* dgamma, float* dbeta,
                  int C, cudaStream_t st) {
  k_bn_bwd_reduce<<<(C + 127) / 128, 128, 0, st>>>(stats, (double)count, gamma, bnc, dgamma, dbeta, C, nullptr);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// nn.CrossEntropyLoss() (NB:2653, NB:3463): mean over the batch of -log softmax(logits)[label].
// Single block (the batch is at most a few thousand rows) so the loss is reduced in a fixed order.
// If sse != NULL also assembles {alpha*mse + ce, mse, ce} (NB:2681).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_softmax_ce(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                    int B, int C, float gscale, float* __restrict__ loss,
                                                    float* __restrict__ dlogits, int* __restrict__ correct,
                                                    const double* __restrict__ sse, double numel, float alpha) {
  __shared__ double red[8];
  __shared__ int redc[8];
  double acc = 0.0;
  int ok = 0;
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    const float* row = logits + (size_t)r * C;
    float mx = row[0];
    int am = 0;
    for (int c = 1; c < C; ++c) { const float v = row[c]; if (v > mx) { mx = v; am = c; } }
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
    const int lab = (int)labels[r];
    const float lse = logf(se) + mx;
    acc += (double)(lse - row[lab]);
    ok += (am == lab);
    if (dlogits) {
      const float inv = gscale / ((float)B * se);
      for (int c = 0; c < C; ++c) {
        const float pmass = expf(row[c] - mx) * inv;
        dlogits[(size_t)r * C + c] = pmass - (c == lab ? gscale / (float)B : 0.f);
      }
    }
  }
  acc = warp_sum_d(acc);
  for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redc[threadIdx.x >> 5] = ok; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    int k = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[w]; k += redc[w]; }
    const float ce = (float)(s / (double)B);
    if (sse) {
      const float mse = (float)(*sse / numel);
      loss[0] = alpha * mse + ce;
      loss[1] = mse;
      loss[2] = ce;
    } else {
      loss[0] = ce;
    }
    if (correct) *correct = k;
  }
}

int softmax_ce(const float* logits, const int64_t* labels, int B, int C, float gscale, float* loss, float* dlogits,
               int* correct, const double* sse, double numel, float alpha, cudaStream_t st) {
  k_softmax_ce<<<1, 256, 0, st>>>(logits, labels, B, C, gscale, loss, dlogits, correct, sse, numel, alpha);
  AE_LAUNCH_CHECK();
  return 0;
}

// nn.MSELoss() (NB:2652) on an already-sigmoided x_hat + gradient w.r.t. the pre-sigmoid activation
__global__ void __launch_bounds__(256) k_sigmoid_mse(const float* __restrict__ x_hat, const float* __restrict__ x,
                                                     int64_t n, float scale, double* __restrict__ sse,
                                                     float* __restrict__ d_pre) {
  __shared__ float red[8];
  float acc = 0.f;
  const float g = scale * 2.f / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = x_hat[i], d = s - x[i];
    acc += d * d;
    if (d_pre) d_pre[i] = g * d * s * (1.f - s);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += (double)red[w];
    atomicAdd(sse, s);
  }
}
__global__ void k_finish_mse(const double* sse, double n, float* loss) { *loss = (float)(*sse / n); }

int sigmoid_mse(const float* x_hat, const float* x, int64_t n, float scale, float* loss, float* d_pre, double* sse_tmp,
                cudaStream_t st) {
  AE_CUDA(cudaMemsetAsync(sse_tmp, 0, sizeof(double), st));
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_sigmoid_mse<<<(int)blocks, 256, 0, st>>>(x_hat, x, n, scale, sse_tmp, d_pre);
  AE_LAUNCH_CHECK();
  k_finish_mse<<<1, 1, 0, st>>>(sse_tmp, (double)n, loss);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// torch.optim.Adam.step() over one flat fp32 buffer (NB:2654/2684, NB:3461/3482): 128-bit loads and
// stores, 28 B of traffic per parameter.  step[0] = step count, step[1] = block-done counter: the
// last block to finish bumps the count, so a captured graph can replay the launch unchanged.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n4, float lr, float b1, float b2, float eps,
                                              float wd, float gscale, int* __restrict__ step, int bump) {
  const int t = step[0] + 1;
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gv.x, gv.y, gv.z, gv.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = ge[k] * gscale;
      if (wd != 0.f) gg = fmaf(wd, pe[k], gg);
      me[k] = me[k] + (gg - me[k]) * omb1;            // exp_avg.lerp_(grad, 1-beta1)
      ve[k] = ve[k] * b2 + (gg * gg) * omb2;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
      const float denom = sqrtf(ve[k]) / bc2_sqrt + eps;
      pe[k] = pe[k] - step_size * (me[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
  }
  __syncthreads();
  if (bump && threadIdx.x == 0) {
    __threadfence();
    const int done = atomicAdd(&step[1], 1);
    if (done == (int)gridDim.x - 1) {
      step[1] = 0;
      step[0] = t;
      __threadfence();
    }
  }
}

// bump == 0: apply step number step[0]+1 without advancing the counter (a flat buffer updated by several launches)
int adam_step_flat_range(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                         float wd, float gscale, int* step_dev, int bump, cudaStream_t st) {
  AE_CHECK(n % 4 == 0, "adam_step_flat: n=%lld must be a multiple of 4 (pad the flat buffer)", (long long)n);
  AE_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step_flat: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  k_adam<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n4, lr, b1, b2, eps, wd, gscale, step_dev, bump);
  AE_LAUNCH_CHECK();
  return 0;
}

int adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                   float wd, float gscale, int* step_dev, cudaStream_t st) {
  AE_CHECK(n % 4 == 0, "adam_step_flat: n=%lld must be a multiple of 4 (pad the flat buffer)", (long long)n);
  AE_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step_flat: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  k_adam<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n4, lr, b1, b2, eps, wd, gscale, step_dev, 1);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// boundary layout conversions
// ---------------------------------------------------------------------------------------------
template <typename TD, bool TO_NHWC>
__global__ void k_layout(const void* __restrict__ src, void* __restrict__ dst, int N, int C, int H, int W) {
  const int64_t total = (int64_t)N * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the destination
    if (TO_NHWC) {
      const int c = (int)(i % C);
      const int64_t p = i / C;
      const int w = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
      const float v = static_cast<const float*>(src)[(((int64_t)n * C + c) * H + h) * W + w];
      static_cast<TD*>(dst)[i] = (TD)v;
    } else {
      const int w = (int)(i % W), h = (int)((i / W) % H);
      const int c = (int)((i / ((int64_t)W * H)) % C), n = (int)(i / ((int64_t)W * H * C));
      const TD v = static_cast<const TD*>(src)[(((int64_t)n * H + h) * W + w) * C + c];
      static_cast<float*>(dst)[i] = (float)v;
    }
  }
}

template <typename TD, bool TO_NHWC>
static int layout_launch(const void* src, void* dst, int N, int C, int H, int W, cudaStream_t st) {
  const int64_t total = (int64_t)N * C * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_layout<TD, TO_NHWC><<<(int)blocks, 256, 0, st>>>(src, dst, N, C, H, W);
  AE_LAUNCH_CHECK();
  return 0;
}

int layout_convert(const void* src, void* dst, int N, int C, int H, int W, bool bf16, bool to_nhwc, cudaStream_t st) {
  if (bf16) return to_nhwc ? layout_launch<__nv_bfloat16, true>(src, dst, N, C, H, W, st)
                           : layout_launch<__nv_bfloat16, false>(src, dst, N, C, H, W, st);
  return to_nhwc ? layout_launch<float, true>(src, dst, N, C, H, W, st) : layout_launch<float, false>(src, dst, N, C, H, W, st);
}

}  // namespace ae

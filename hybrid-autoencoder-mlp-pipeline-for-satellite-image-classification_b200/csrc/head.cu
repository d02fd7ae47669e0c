// Classifier head of the supervised autoencoder (NB:692-696): Linear(L,128) -> ReLU -> Linear(128,C).
// The 128-wide layer runs on the generic GEMM kernels; the C(=10)-wide layer is too narrow for them.
#include "common.cuh"

namespace ae {

// logits[r][c] = b2[c] + sum_j relu(hid_pre[r][j]) * W2[c][j]
__global__ void __launch_bounds__(128) k_head_out(const float* __restrict__ hid_pre, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, float* __restrict__ logits, int B,
                                                  int H, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int r = idx / C, c = idx - r * C;
  const float* h = hid_pre + (size_t)r * H;
  const float* w = w2 + (size_t)c * H;
  float acc = 0.f;
  for (int j = 0; j < H; j += 4) {
    const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + j));
    acc = fmaf(fmaxf(hv.x, 0.f), wv.x, acc); acc = fmaf(fmaxf(hv.y, 0.f), wv.y, acc);
    acc = fmaf(fmaxf(hv.z, 0.f), wv.z, acc); acc = fmaf(fmaxf(hv.w, 0.f), wv.w, acc);
  }
  logits[idx] = acc + __ldg(b2 + c);
}

int head_out(const float* hid_pre, const float* w2, const float* b2, float* logits, int B, int H, int C, cudaStream_t st) {
  k_head_out<<<(B * C + 127) / 128, 128, 0, st>>>(hid_pre, w2, b2, logits, B, H, C);
  AE_LAUNCH_CHECK();
  return 0;
}

// dhid[r][j] = (hid_pre[r][j] > 0) * sum_c dlogits[r][c] * W2[c][j]
__global__ void __launch_bounds__(128) k_head_dhid(const float* __restrict__ dlogits, const float* __restrict__ w2,
                                                   const float* __restrict__ hid_pre, float* __restrict__ dhid, int B,
                                                   int H, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int r = idx / H, j = idx - r * H;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(__ldg(dlogits + (size_t)r * C + c), __ldg(w2 + (size_t)c * H + j), acc);
  dhid[idx] = hid_pre[idx] > 0.f ? acc : 0.f;
}

// block c: dW2[c][j] = sum_r dlogits[r][c] * relu(hid_pre[r][j]);  db2[c] = sum_r dlogits[r][c]
__global__ void __launch_bounds__(128) k_head_w2grad(const float* __restrict__ dlogits, const float* __restrict__ hid_pre,
                                                     float* __restrict__ dw2, float* __restrict__ db2, int B, int H, int C) {
  const int c = blockIdx.x;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < B; ++r) acc = fmaf(__ldg(dlogits + (size_t)r * C + c), fmaxf(__ldg(hid_pre + (size_t)r * H + j), 0.f), acc);
    dw2[(size_t)c * H + j] = acc;
  }
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += 32) s += __ldg(dlogits + (size_t)r * C + c);
    s = warp_sum(s);
    if (threadIdx.x == 0) db2[c] = s;
  }
}

int head_backward_small(const float* dlogits, const float* w2, const float* hid_pre, float* dhid, float* dw2, float* db2,
                        int B, int H, int C, cudaStream_t st) {
  k_head_dhid<<<(B * H + 127) / 128, 128, 0, st>>>(dlogits, w2, hid_pre, dhid, B, H, C);
  AE_LAUNCH_CHECK();
  k_head_w2grad<<<C, 128, 0, st>>>(dlogits, hid_pre, dw2, db2, B, H, C);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused head step (NB:692-696 forward, NB:2680 cross-entropy, their backward) in ONE launch for the fused
// training step: W1 / W2 live in shared memory; each CTA owns HF_ROWS batch rows, writes its logits, its share
// of the loss, dz and a partial of every weight gradient; k_head_reduce sums the partials in a fixed order.
// ---------------------------------------------------------------------------------------------
static constexpr int HF_ROWS = 16, HF_THREADS = 256, HF_H = 128;

struct HeadFused {
  const float *z, *w1, *b1, *w2, *b2;
  const int64_t* labels;
  float *logits, *dz, *loss;
  float *gw1, *gb1, *gw2, *gb2;
  float* partial;           // [ctas][L*128 + 128 + C*128 + C + 1]
  unsigned int* counter;
  const double* sse;
  double numel;
  float alpha;
  int B, L, C;
};

__global__ void __launch_bounds__(HF_THREADS) k_head_fused(HeadFused a) {
  extern __shared__ float sm[];
  const int L = a.L, C = a.C, LP = L + 1;
  float* W1s = sm;                              // [128][L+1]
  float* W2s = W1s + HF_H * LP;                 // [C][128]
  float* zs = W2s + C * HF_H;                   // [R][L]
  float* hid = zs + HF_ROWS * L;                // [R][128] pre-activation
  float* dh = hid + HF_ROWS * HF_H;             // [R][128]
  float* dl = dh + HF_ROWS * HF_H;              // [R][16]
  float* lrow = dl + HF_ROWS * 16;              // [R] per-row loss
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * HF_ROWS;
  const int nr = min(HF_ROWS, a.B - r0);

  {  // W1: L is a multiple of 16, so 128-bit loads; all of a thread's loads are issued before the first store
    const int n4 = HF_H * L / 4;
    for (int i0 = tid; i0 < n4; i0 += 8 * HF_THREADS) {
      float4 buf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * HF_THREADS;
        buf[u] = i < n4 ? __ldg(reinterpret_cast<const float4*>(a.w1) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * HF_THREADS;
        if (i < n4) {
          const int e0 = i * 4, j = e0 / L, k = e0 - j * L;
          float* d = W1s + j * LP + k;
          d[0] = buf[u].x; d[1] = buf[u].y; d[2] = buf[u].z; d[3] = buf[u].w;
        }
      }
    }
  }
  for (int i = tid; i < C * HF_H; i += HF_THREADS) W2s[i] = __ldg(a.w2 + i);
  for (int i = tid; i < HF_ROWS * L; i += HF_THREADS) zs[i] = (i / L) < nr ? __ldg(a.z + (size_t)r0 * L + i) : 0.f;
  __syncthreads();
  // hidden layer
  {
    const int j = tid & 127;
    const float bj = __ldg(a.b1 + j);
    for (int r = tid >> 7; r < HF_ROWS; r += 2) {
      float acc = bj;
      for (int k = 0; k < L; ++k) acc = fmaf(zs[r * L + k], W1s[j * LP + k], acc);
      hid[r * HF_H + j] = acc;
    }
  }
  __syncthreads();
  // logits
  for (int i = tid; i < HF_ROWS * C; i += HF_THREADS) {
    const int r = i / C, c = i - r * C;
    float acc = __ldg(a.b2 + c);
    for (int j = 0; j < HF_H; ++j) acc = fmaf(fmaxf(hid[r * HF_H + j], 0.f), W2s[c * HF_H + j], acc);
    dl[r * 16 + c] = acc;
    if (r < nr) a.logits[(size_t)(r0 + r) * C + c] = acc;
  }
  __syncthreads();
  // softmax cross-entropy per row
  if (tid < HF_ROWS) {
    const int r = tid;
    float loss = 0.f;
    if (r < nr) {
      float mx = dl[r * 16];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, dl[r * 16 + c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(dl[r * 16 + c] - mx);
      const int lab = (int)a.labels[r0 + r];
      loss = logf(se) + mx - dl[r * 16 + lab];
      const float inv = 1.f / ((float)a.B * se);
      for (int c = 0; c < C; ++c) dl[r * 16 + c] = expf(dl[r * 16 + c] - mx) * inv - (c == lab ? 1.f / (float)a.B : 0.f);
    } else {
      for (int c = 0; c < C; ++c) dl[r * 16 + c] = 0.f;
    }
    lrow[r] = loss;
  }
  __syncthreads();
  // d hidden
  {
    const int j = tid & 127;
    for (int r = tid >> 7; r < HF_ROWS; r += 2) {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(dl[r * 16 + c], W2s[c * HF_H + j], acc);
      dh[r * HF_H + j] = hid[r * HF_H + j] > 0.f ? acc : 0.f;
    }
  }
  __syncthreads();
  // dz = dh * W1
  for (int i = tid; i < HF_ROWS * L; i += HF_THREADS) {
    const int r = i / L, k = i - r * L;
    float acc = 0.f;
    for (int j = 0; j < HF_H; ++j) acc = fmaf(dh[r * HF_H + j], W1s[j * LP + k], acc);
    if (r < nr) a.dz[(size_t)(r0 + r) * L + k] = acc;
  }
  // partial gradients of this CTA
  const int psize = HF_H * L + HF_H + C * HF_H + C + 1;
  float* part = a.partial + (size_t)blockIdx.x * psize;
  for (int i = tid; i < HF_H * L; i += HF_THREADS) {       // dW1[j][k] = sum_r dh[r][j] * z[r][k]
    const int j = i / L, k = i - j * L;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < HF_ROWS; ++r) acc = fmaf(dh[r * HF_H + j], zs[r * L + k], acc);
    part[i] = acc;
  }
  for (int j = tid; j < HF_H; j += HF_THREADS) {
    float acc = 0.f;
    for (int r = 0; r < HF_ROWS; ++r) acc += dh[r * HF_H + j];
    part[HF_H * L + j] = acc;
  }
  for (int i = tid; i < C * HF_H; i += HF_THREADS) {       // dW2[c][j] = sum_r dl[r][c] * relu(hid[r][j])
    const int c = i / HF_H, j = i - c * HF_H;
    float acc = 0.f;
    for (int r = 0; r < HF_ROWS; ++r) acc = fmaf(dl[r * 16 + c], fmaxf(hid[r * HF_H + j], 0.f), acc);
    part[HF_H * L + HF_H + i] = acc;
  }
  if (tid < C) {
    float acc = 0.f;
    for (int r = 0; r < HF_ROWS; ++r) acc += dl[r * 16 + tid];
    part[HF_H * L + HF_H + C * HF_H + tid] = acc;
  }
  if (tid == 0) {
    float acc = 0.f;
    for (int r = 0; r < HF_ROWS; ++r) acc += lrow[r];
    part[psize - 1] = acc;
  }
}

// Fixed-order sum of the per-CTA partials (deterministic), then the loss assembly {alpha*mse + ce, mse, ce} (NB:2681).
__global__ void __launch_bounds__(256) k_head_reduce(HeadFused a, int nparts) {
  const int L = a.L, C = a.C;
  const int psize = HF_H * L + HF_H + C * HF_H + C + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= psize) return;
  float acc = 0.f;
  int p = 0;
  for (; p + 8 <= nparts; p += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(a.partial + (size_t)(p + u) * psize + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; p < nparts; ++p) acc += __ldg(a.partial + (size_t)p * psize + i);
  if (i < HF_H * L) a.gw1[i] = acc;
  else if (i < HF_H * L + HF_H) a.gb1[i - HF_H * L] = acc;
  else if (i < HF_H * L + HF_H + C * HF_H) a.gw2[i - HF_H * L - HF_H] = acc;
  else if (i < psize - 1) a.gb2[i - HF_H * L - HF_H - C * HF_H] = acc;
  else {
    const float cem = acc / (float)a.B;
    if (a.sse) {
      const float mse = (float)(*a.sse / a.numel);
      a.loss[0] = a.alpha * mse + cem; a.loss[1] = mse; a.loss[2] = cem;
    } else {
      a.loss[0] = cem;
    }
  }
}

size_t head_fused_workspace_floats(int B, int L, int C) {
  const int ctas = (B + HF_ROWS - 1) / HF_ROWS;
  return (size_t)ctas * (HF_H * L + HF_H + C * HF_H + C + 1);
}

// `st_main` launches the per-row kernel; `st_finish` (may be the same stream) launches the reduction that also reads
// the decoder's squared-error sum -- the caller orders st_finish after st_main and after the decoder forward.
int head_fused_step(const float* z, const int64_t* labels, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* logits, float* dz, float* gw1, float* gb1, float* gw2, float* gb2,
                    float* loss, const double* sse, double numel, float alpha, float* partial, unsigned int* counter,
                    int B, int L, int C, cudaStream_t st_main, cudaStream_t st_finish, int phase) {
  AE_CHECK(C <= 16 && L <= 256, "head_fused_step: num_classes=%d (max 16) / latent_dim=%d (max 256) out of range", C, L);
  HeadFused a;
  a.z = z; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.labels = labels; a.logits = logits; a.dz = dz; a.loss = loss;
  a.gw1 = gw1; a.gb1 = gb1; a.gw2 = gw2; a.gb2 = gb2; a.partial = partial; a.counter = counter; a.sse = sse;
  a.numel = numel; a.alpha = alpha; a.B = B; a.L = L; a.C = C;
  const int ctas = (B + HF_ROWS - 1) / HF_ROWS;
  if (phase == 0 || phase == 1) {
    const size_t smem = sizeof(float) * ((size_t)HF_H * (L + 1) + (size_t)C * HF_H + (size_t)HF_ROWS * L +
                                         2 * (size_t)HF_ROWS * HF_H + HF_ROWS * 16 + HF_ROWS);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
      AE_CUDA(cudaFuncSetAttribute(k_head_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    k_head_fused<<<ctas, HF_THREADS, smem, st_main>>>(a);
    AE_LAUNCH_CHECK();
  }
  if (phase == 0 || phase == 2) {
    const int psize = HF_H * L + HF_H + C * HF_H + C + 1;
    k_head_reduce<<<(psize + 255) / 256, 256, 0, st_finish>>>(a, ctas);
    AE_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace ae

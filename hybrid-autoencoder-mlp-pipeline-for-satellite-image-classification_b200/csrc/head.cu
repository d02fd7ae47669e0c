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

}  // namespace ae

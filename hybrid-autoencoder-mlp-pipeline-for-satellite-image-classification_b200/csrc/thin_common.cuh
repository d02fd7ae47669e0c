// Shared pieces of the 3-channel-layer kernels (thin.cu).
#pragma once
#include "tc_common.cuh"

namespace ae {

static constexpr int TH = 64, TW = 64, WH = 32, WW = 32, WC = 32;
static constexpr int TT_THREADS = 128;
static constexpr int TILE_ROWS = 4;                    // wide rows per tile
static constexpr int TILES_PER_IMAGE = WH / TILE_ROWS; // 8
static constexpr int XS_ROWS = 2 * TILE_ROWS + 1;      // thin rows 8*tr-1 .. 8*tr+7
static constexpr int XS_PITCH = 72;                    // thin column c at index c + 4; index 3 = left zero padding
static constexpr int XS_FLOATS = (3 * XS_ROWS * XS_PITCH + 31) / 32 * 32;   // one TMA box [3][XS_ROWS][XS_PITCH], padded to 128 bytes
static constexpr int TW_PART = 868;                    // 864 weights + 3 thin-bias sums + 1 pad

__device__ __forceinline__ float thin_transform(const Operand& op, float a, float s) {
  if (op.mode == AE_OP_RAW) return a;
  const float up = (op.scalar != 0.f) ? op.scalar * (s - a) : a;   // fused MSE gradient, or a given upstream gradient
  return up * s * (1.f - s);                                       // AE_OP_SIGMOID_BWD
}

// (d0, d1) += a * (b0, b1): one packed fp32 FMA (FFMA2, scalar operand broadcast); each half rounds like fmaf
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
  const float2 r = __ffma2_rn(make_float2(a, a), make_float2(b0, b1), make_float2(d0, d1));
  d0 = r.x; d1 = r.y;
}

// lane l ends with the sum over the warp's 32 lanes of element v[l]  (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? v[k] : v[k + off];
      const float keep = upper ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

static constexpr int XS_BYTES = 3 * XS_ROWS * XS_PITCH * 4;   // bytes one image box delivers
static constexpr int WT_FLOATS = 128 * 32;             // one wide tile
static constexpr int WT_BYTES = WT_FLOATS * 4;

// shared-memory layout of one k_thin stage (floats): [xs][xs2 if the thin operand has two sources][wide][wide2 if BNBWD]
struct ThinStage {
  int xs2, wide, wide2, floats;
};
__host__ __device__ inline ThinStage thin_stage_layout(int thin_mode, int wide_mode, bool wgrad) {
  ThinStage L;
  int off = XS_FLOATS;
  L.xs2 = off; if (thin_mode != AE_OP_RAW) off += XS_FLOATS;
  L.wide = off; if (wgrad) off += WT_FLOATS;
  L.wide2 = off; if (wgrad && wide_mode == AE_OP_BNBWD) off += WT_FLOATS;
  L.floats = off;
  return L;
}


}  // namespace ae

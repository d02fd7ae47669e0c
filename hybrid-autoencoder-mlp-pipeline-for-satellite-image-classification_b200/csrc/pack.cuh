// Element-wise weight re-layout routines shared by the stand-alone pack kernels and the fused k_pack_all.
#pragma once
#include "common.cuh"

namespace ae {

// Data-gradient weight pack ("group pack").  Per (n-tile, K chunk c) a block of 9 tiles, ordered so that the tiles one MMA of
// the second-generation row GEMM stacks along N are one contiguous region -- [hi plane of its tiles][lo plane of its tiles] --
// and arrive with ONE cp.async.bulk:
//   group 0 (shift (0,0)): slots 0, 2      group 1 (shift (0,0)): slots 8, 4      group 2 (shift (0,1)): slots 1, 7
//   group 3 (shift (1,0)): slots 6, 3      group 4 (shift (1,1)): slot 5
// (slot = position in the phase-stacked tap order: 0: phase 0; 1-2: phase 1; 3-4: phase 2; 5-8: phase 3).
// dg_pack_plane_offset: byte offset of plane `pl` of the tile in `slot` inside the 9-tile block; plane_bytes = NT * 128.
__host__ __device__ __forceinline__ int dg_pack_group_of_slot(int slot, int* t) {
  const int g = slot == 0 || slot == 2 ? 0 : slot == 8 || slot == 4 ? 1 : slot == 1 || slot == 7 ? 2 : slot == 6 || slot == 3 ? 3 : 4;
  *t = (slot == 2 || slot == 4 || slot == 7 || slot == 3) ? 1 : 0;
  return g;
}
__host__ __device__ __forceinline__ size_t dg_pack_plane_offset(int slot, int pl, int nsplit, size_t plane_bytes) {
  int t;
  const int g = dg_pack_group_of_slot(slot, &t);
  const int ntl = g == 4 ? 1 : 2;
  return ((size_t)(2 * g) * nsplit + (size_t)pl * ntl + t) * plane_bytes;      // groups 0..3 hold two tiles each
}

// Eight consecutive K elements (one 16-byte chunk of a swizzled row; idx8 in [0, 2*9*Cs*Cb/8)) of the conv weight pack:
// w [Cs][Cb][3][3] fp32 -> swizzled bf16 (hi[, lo]) tiles for both row-GEMM orientations (layout described in tma_gemm.cu).
// One 16-byte store per plane instead of eight 2-byte stores.
__device__ __forceinline__ void pack_conv_chunk(int idx8, const float* __restrict__ w, int Cs, int Cb, int nsplit,
                                                uint8_t* __restrict__ fwd, uint8_t* __restrict__ dgrad) {
  const int KCf = Cb >= 64 ? 64 : 32, NTf = Cs >= 64 ? 64 : 32;
  const int NTd = Cb >= 64 ? 64 : 32;
  const int nf8 = Cs * 9 * Cb / 8;
  float v[8];
  uint8_t* base;
  int r, j, NT, KC;
  size_t tile = 0, dg_off_hi = 0, dg_off_lo = 0;
  bool is_dgrad = false;
  if (idx8 < nf8) {
    const int k = (idx8 % (9 * Cb / 8)) * 8, n = idx8 / (9 * Cb / 8);   // n = cs, k = tap*Cb + cb (8 consecutive cb)
    const int tap = k / Cb, cb = k - tap * Cb;
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = w[((size_t)n * Cb + cb + u) * 9 + tap];
    KC = KCf; NT = NTf;
    const int kc = k / KC; j = k - kc * KC;
    r = n % NT; tile = (size_t)(n / NT) * (9 * Cb / KC) + kc; base = fwd;
  } else {
    const int i2 = idx8 - nf8;
    const int k = (i2 % (9 * Cs / 8)) * 8, n = i2 / (9 * Cs / 8);       // n = cb, k = slot*Cs + cs (8 consecutive cs)
    const int slot = k / Cs, cs = k - slot * Cs;        // slot 0: phase 0; 1-2: phase 1; 3-4: phase 2; 5-8: phase 3
    int ky, kx;
    if (slot == 0) { ky = 1; kx = 1; }
    else if (slot <= 2) { ky = 1; kx = slot == 1 ? 0 : 2; }
    else if (slot <= 4) { ky = slot == 3 ? 0 : 2; kx = 1; }
    else { const int t = slot - 5; ky = (t >> 1) ? 2 : 0; kx = (t & 1) ? 2 : 0; }
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = w[((size_t)(cs + u) * Cb + n) * 9 + ky * 3 + kx];
    KC = 64; NT = NTd;
    const int c = cs / KC; j = cs - c * KC;
    r = n % NT; base = dgrad;
    const int cpt = Cs / KC;
    const size_t plane_bytes = (size_t)NT * 128;
    const size_t blk = ((size_t)(n / NT) * cpt + c) * 9 * nsplit * plane_bytes;
    const int swz_d = r & 7;
    const size_t in_plane = (size_t)r * 128 + (size_t)(((j >> 3) ^ swz_d) << 4);
    dg_off_hi = blk + dg_pack_plane_offset(slot, 0, nsplit, plane_bytes) + in_plane;
    dg_off_lo = blk + dg_pack_plane_offset(slot, 1, nsplit, plane_bytes) + in_plane;
    is_dgrad = true;
  }
  const int rowb = KC * 2;
  const int swz = rowb == 128 ? (r & 7) : ((r >> 1) & 3);
  const size_t tile_bytes = (size_t)nsplit * NT * rowb;
  size_t off = tile * tile_bytes + (size_t)r * rowb + (size_t)(((j >> 3) ^ swz) << 4);
  size_t off_lo = off + (size_t)NT * rowb;
  if (is_dgrad) { off = dg_off_hi; off_lo = dg_off_lo; }
  __nv_bfloat162 h[4];
  float lo[8];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    lo[2 * u] = v[2 * u] - __bfloat162float(h[u].x);
    lo[2 * u + 1] = v[2 * u + 1] - __bfloat162float(h[u].y);
  }
  *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(h);
  if (nsplit == 2) {
    __nv_bfloat162 l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) l[u] = __floats2bfloat162_rn(lo[2 * u], lo[2 * u + 1]);
    *reinterpret_cast<uint4*>(base + off_lo) = *reinterpret_cast<const uint4*>(l);
  }
}

// Linear weight w [N][K] (torch).  perm(k) = (k % permC) * permHW + k / permC maps an NHWC flatten index to the
// reference's (C,H,W) flatten index (NB:520 / NB:614).
//  kind 0: dst[k][n]  = w[n][perm(k)]      (K x N, forward of a layer whose INPUT is NHWC-flattened)
//  kind 1: dst[n][k]  = w[n][perm(k)]      (N x K, its data-gradient operand)
//  kind 2: dst[k][n'] = w[perm(n')][k]     (K x N, forward of a layer whose OUTPUT is NHWC-flattened)
//  kind 3: dst[n'][k] = w[perm(n')][k]     (N x K, its data-gradient operand)
__device__ __forceinline__ void pack_linear_elem(int idx, const float* __restrict__ w, int N, int K, int permC, int permHW,
                                                 int kind, float* __restrict__ dst) {
  const int k = idx % K, n = idx / K;  // destination-side logical (n, k)
  if (kind == 0 || kind == 1) {
    const int kp = permC > 0 ? (k % permC) * permHW + k / permC : k;
    const float v = w[(size_t)n * K + kp];
    if (kind == 0) dst[(size_t)k * N + n] = v; else dst[(size_t)n * K + k] = v;
  } else {
    const int np = permC > 0 ? (n % permC) * permHW + n / permC : n;
    const float v = w[(size_t)np * K + k];
    if (kind == 2) dst[(size_t)k * N + n] = v; else dst[(size_t)n * K + k] = v;
  }
}

// Tensor-core pack of a Linear weight w [N][K] (torch) whose INPUT is NHWC-flattened (dense_tc.cu): per 64-deep K chunk one
// SWIZZLE_128B tile [hi plane][lo plane] of N rows x 128 bytes; K index k of the pack is the reference's perm(k).
// idx8 in [0, N*K/8): 8 consecutive K elements = one 16-byte chunk of a row.
__device__ __forceinline__ void pack_dense_tc_chunk(int idx8, const float* __restrict__ w, int N, int K, int permC, int permHW,
                                                    int nsplit, uint8_t* __restrict__ dst) {
  const int n = idx8 / (K / 8), k = (idx8 - n * (K / 8)) * 8;
  const int chunk = k / 64, j = k - chunk * 64;
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int ku = k + u;
    v[u] = w[(size_t)n * K + (permC > 0 ? (ku % permC) * permHW + ku / permC : ku)];
  }
  const size_t plane = (size_t)N * 128;
  const size_t off = (size_t)chunk * nsplit * plane + (size_t)n * 128 + (size_t)(((j >> 3) ^ (n & 7)) << 4);
  __nv_bfloat162 h[4], l[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    l[u] = __floats2bfloat162_rn(v[2 * u] - __bfloat162float(h[u].x), v[2 * u + 1] - __bfloat162float(h[u].y));
  }
  *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(h);
  if (nsplit == 2) *reinterpret_cast<uint4*>(dst + off + plane) = *reinterpret_cast<const uint4*>(l);
}

__device__ __forceinline__ void permute_elem(int i, const float* __restrict__ src, int permC, int permHW, float* __restrict__ dst) {
  dst[i] = src[(i % permC) * permHW + i / permC];
}

// A batch of re-layout jobs executed by ONE launch (k_pack_all): after every optimizer step the engine re-derives
// the six conv weight packs and the dense-layer packs.
enum { PACK_CONV = 0, PACK_LINEAR = 1, PACK_PERMUTE = 2, PACK_DENSE_TC = 3 };
struct PackJob {
  int kind;
  const float* src;
  void* dst;
  void* dst2;
  int a, b, c, d, e;        // CONV: Cs, Cb, nsplit.  LINEAR: N, K, permC, permHW, lkind.  PERMUTE: n, permC, permHW.  DENSE_TC: N, K, permC, permHW, nsplit
  int total;                // work items: 8-element chunks (CONV) or elements
  int first_block;          // first block of this job in the fused launch
};
static constexpr int PACK_MAX_JOBS = 16;
struct PackJobs {
  PackJob job[PACK_MAX_JOBS];
  int n;
};
int pack_all(PackJobs& jobs, cudaStream_t st);

}  // namespace ae

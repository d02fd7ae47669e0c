// The two 3-channel layers: Conv2d(3,32,k3,s2,p1) (NB:504) and ConvTranspose2d(32,3,k3,s2,p1,op1)+Sigmoid
// (NB:628-629).  K=27 / N=3 do not fill a tensor-core tile; these are bandwidth kernels on CUDA cores.
//   thin = [B,3,64,64] NCHW fp32 (the reference's image layout), wide = [B,32,32,32] NHWC fp32.
//   both layers store their weight as [32][3][3][3] = [c32][c3][ky][kx].
//
// All kernels are persistent and work on tiles of 4 wide rows x 32 wide columns (128 wide pixels = 8 thin rows) of
// one image.  The raw thin rows of the NEXT tile arrive with ONE TMA tensor load per source (box [3 channels][9 rows][72
// columns] of the NCHW image starting at column -4, row 8*tr-1: the zero padding of the convolution is the tensor map's
// out-of-bounds fill) and its raw wide tile with one bulk copy (mbarrier, double buffered) while the current tile is
// computed; one in-place pass applies the operand transform once per element.  (Round 2: the 27 row-by-row bulk copies
// this replaced kept warp 0 issuing for ~2.7 k cycles per tile while the other warps waited at the barrier -- ncu: 35 % of
// all stall samples.)  128 threads; in the gather every thread owns 8 consecutive wide pixels x 4
// output channels (one 16-byte weight load and 5/3 patch loads feed 32 FMAs: the kernel is bound by FMA issue, not by
// the shared-memory pipe).
#include <cstring>

#include "thin_common.cuh"
#include "tma_host.cuh"

namespace ae {

// ---------------------------------------------------------------------------------------------
// k_thin: thin -> wide gather (conv1 forward; convT4 data gradient) and / or the weight gradient of a thin layer
//   GATHER: out[p][c32] = epilogue(sum_k patch(p,k) * w[c32][k]), k = (c3,ky,kx); per-channel statistics
//   WGRAD : partial[cta][c32][k] = sum_p wide(p,c32) * patch(p,k); partial[cta][864+c3] = sum of the thin operand
// Both read the same staged thin rows, so the fused convT4 backward touches x / x_hat once.
// ---------------------------------------------------------------------------------------------
// CTAs per SM: the gather alone fits 4 (registers); the stagings of the weight gradient limit the fused backward to 3 and
// the weight gradient with a BatchNorm-backward operand to 2 -- the register budget follows (128 / 168 / 255)
template <bool GATHER, bool WGRAD>
__global__ void __launch_bounds__(TT_THREADS, WGRAD ? (GATHER ? 3 : 2) : 4) k_thin(Operand thin, Operand wide, const float* __restrict__ w, Epilogue e,
                                                     float* __restrict__ out, float* __restrict__ partial, int batch,
                                                     const __grid_constant__ CUtensorMap xmap,
                                                     const __grid_constant__ CUtensorMap xmap2) {
  extern __shared__ __align__(128) float smem_f[];
  const ThinStage L = thin_stage_layout(thin.mode, wide.mode, WGRAD);
  float* stage0 = smem_f;                               // two stages of L.floats floats
  float* Wsm = smem_f + 2 * L.floats;                   // [27][32]                 (GATHER)
  float* sbn = Wsm + 27 * 32;                           // [4][32] scale/shift/mean/rstd (GATHER, RELUBWD)
  float* wbn = sbn + 4 * 32;                            // [4][32] coefficients of the wide operand (WGRAD)
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ float sStat[2][32];
  __shared__ float sB[TT_THREADS / 32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(&bars[0]);

  const int tiles = batch * TILES_PER_IMAGE;
  // one thread: fetch the raw data of `tile` into stage st
  auto issue = [&](int tile, int st) {
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const uint32_t bar = bar0 + 8u * st;
    float* base = stage0 + st * L.floats;
    const int nsrc = thin.mode != AE_OP_RAW ? 2 : 1;
    uint32_t bytes = (uint32_t)(XS_BYTES * nsrc);           // a box always delivers all its bytes (zeros where out of bounds)
    if (WGRAD) bytes += WT_BYTES * (wide.mode == AE_OP_BNBWD ? 2 : 1);
    mbar_arrive_expect_tx(bar, bytes);
    // thin rows 8*tr-1 .. 8*tr+7, columns -4 .. 67: column c lands at index c + 4 of its XS_PITCH-wide row
    tma_load_4d(smem_u32(base), &xmap, -4, 2 * TILE_ROWS * tr - 1, 0, n, bar);
    if (nsrc == 2) tma_load_4d(smem_u32(base + L.xs2), &xmap2, -4, 2 * TILE_ROWS * tr - 1, 0, n, bar);
    if (WGRAD) {
      const size_t m0 = ((size_t)n * WH + tr * TILE_ROWS) * WW;
      bulk_copy_g2s(smem_u32(base + L.wide), wide.src + m0 * WC, WT_BYTES, bar);
      if (wide.mode == AE_OP_BNBWD) bulk_copy_g2s(smem_u32(base + L.wide2), wide.src2 + m0 * WC, WT_BYTES, bar);
    }
  };

  if (tid == 0) {
    tma_prefetch_desc(&xmap);
    if (thin.mode != AE_OP_RAW) tma_prefetch_desc(&xmap2);
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    // the first tile's data is on its way while the CTA stages the weights and coefficients below
    if ((int)blockIdx.x < tiles) issue(blockIdx.x, 0);
  }
  if (GATHER) {
    for (int i = tid; i < 27 * 32; i += TT_THREADS) {
      const int c32 = i / 27, k = i - c32 * 27;
      Wsm[k * 32 + c32] = __ldg(w + i);
    }
    if (e.mode == AE_EPI_RELUBWD_STATS) {
      const int rows[4] = {AE_BNC_SCALE, AE_BNC_SHIFT, AE_BNC_MEAN, AE_BNC_RSTD};
      sbn[tid] = __ldg(e.bnc + rows[tid >> 5] * WC + lane);
    } else if (e.mode == AE_EPI_BNRELU_SPLIT) {
      if (tid < 64) sbn[tid] = __ldg(e.bnc + (tid >> 5 ? AE_BNC_SHIFT : AE_BNC_SCALE) * WC + lane);
    }
    if (tid < 32) { sStat[0][tid] = 0.f; sStat[1][tid] = 0.f; }
  }
  if (WGRAD && wide.mode != AE_OP_RAW) {
    // BNRELU: scale, shift;  BNBWD: A, B, C, mean
    const int rows_relu[4] = {AE_BNC_SCALE, AE_BNC_SHIFT, AE_BNC_SCALE, AE_BNC_SHIFT};
    const int rows_bwd[4] = {AE_BNC_A, AE_BNC_B, AE_BNC_C, AE_BNC_MEAN};
    wbn[tid] = __ldg(wide.bnc + (wide.mode == AE_OP_BNRELU ? rows_relu[tid >> 5] : rows_bwd[tid >> 5]) * WC + lane);
  }
  // GATHER: this thread's 8 wide pixels (tile row g_r, columns g_x0 .. g_x0+7) and 4 channels (g_c0 .. g_c0+3)
  const int g_c0 = (tid & 7) * 4, g_r = tid >> 5, g_x0 = ((tid >> 3) & 3) * 8;
  float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};   // statistics of the thread's channels
  float4 gbias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (GATHER && e.mode != AE_EPI_RELUBWD_STATS && e.bias) gbias = __ldg(reinterpret_cast<const float4*>(e.bias + g_c0));
  const size_t plane_elems = (size_t)batch * WH * WW * WC;
  float bsum[3] = {0.f, 0.f, 0.f};                        // WGRAD: thin-operand sums
  // WGRAD register tile: channels c4*4..+3 x patch taps kg*7..+6 (tap 27 is padding)
  const int c4 = lane & 7, kg = lane >> 3;
  int koff[7];
  float wacc[4][7];
  // weight gradient alone (two CTAs per SM, 255 registers): a lane owns 4 channels x ALL 27 taps for every 4th pixel of its
  // warp's row -- per pixel one 16-byte operand load and 18 patch loads feed 108 FMAs (the 4 x 7 tile above: 8 loads per 28),
  // which takes the kernel off the shared-memory pipe (ncu: 83 % of its peak with the narrow tile)
  constexpr bool WIDE_W = WGRAD && !GATHER;
  float wfull[WIDE_W ? 4 : 1][WIDE_W ? 27 : 1];
  if (WIDE_W) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 27; ++k) wfull[i][WIDE_W ? k : 0] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    int k = kg * 7 + j;
    if (k > 26) k = 26;
    const int c3 = k / 9, rr = k - c3 * 9, ky = rr / 3, kx = rr - ky * 3;
    koff[j] = (c3 * XS_ROWS + ky) * XS_PITCH + 3 + kx;
#pragma unroll
    for (int i = 0; i < 4; ++i) wacc[i][j] = 0.f;
  }
  __syncthreads();

  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int st = it & 1;
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const size_t m0 = ((size_t)n * WH + tr * TILE_ROWS) * WW;      // first wide pixel of the tile
    float* xs = stage0 + st * L.floats;
    float* as = xs + L.wide;
    if (tid == 0 && tile + (int)gridDim.x < tiles) {
      fence_proxy_async();                                // the other stage was last touched by generic-proxy accesses
      issue(tile + gridDim.x, st ^ 1);
    }
    float4 y4[8];                                         // RELUBWD: raw outputs of this thread's pixels / channels (prefetched)
    const size_t row = (m0 + (size_t)g_r * WW + g_x0) * WC + g_c0;
    if (GATHER && e.mode == AE_EPI_RELUBWD_STATS) {             // in flight while the tile is computed
#pragma unroll
      for (int px = 0; px < 8; ++px) y4[px] = __ldg(reinterpret_cast<const float4*>(e.y + row + px * WC));
    }
    mbar_wait(bar0 + 8u * st, (it >> 1) & 1);
    // ---- transform pass (in place, once per element) ----
    if (thin.mode != AE_OP_RAW || WGRAD) {
      for (int i = tid; i < 3 * XS_ROWS * 16; i += TT_THREADS) {
        const int q = i & 15, r = (i >> 4) % XS_ROWS, c3 = i / (16 * XS_ROWS);
        float* px = xs + (c3 * XS_ROWS + r) * XS_PITCH + 4 + q * 4;
        float4 v = *reinterpret_cast<const float4*>(px);    // thin row -1 (r == 0 of an image's first tile) arrived as zeros
        if (thin.mode != AE_OP_RAW) {
          const float4 sg = *reinterpret_cast<const float4*>(px + L.xs2);
          v.x = thin_transform(thin, v.x, sg.x); v.y = thin_transform(thin, v.y, sg.y);
          v.z = thin_transform(thin, v.z, sg.z); v.w = thin_transform(thin, v.w, sg.w);
          *reinterpret_cast<float4*>(px) = v;
        }
        if (WGRAD && r >= 1) bsum[c3 == 0 ? 0 : (c3 == 1 ? 1 : 2)] += (v.x + v.y) + (v.z + v.w);
      }
    }
    if (WGRAD && wide.mode != AE_OP_RAW) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * TT_THREADS;               // float4 index; channel chunk = i & 7
        const int c = (i & 7) * 4;
        float4 v = *reinterpret_cast<const float4*>(as + i * 4);
        const float4 k0 = *reinterpret_cast<const float4*>(wbn + c), k1 = *reinterpret_cast<const float4*>(wbn + 32 + c);
        if (wide.mode == AE_OP_BNRELU) {
          v.x = fmaxf(fmaf(v.x, k0.x, k1.x), 0.f); v.y = fmaxf(fmaf(v.y, k0.y, k1.y), 0.f);
          v.z = fmaxf(fmaf(v.z, k0.z, k1.z), 0.f); v.w = fmaxf(fmaf(v.w, k0.w, k1.w), 0.f);
        } else {                                          // dy = A*dz + B*(y - mean) + C   (same order as load_operand4)
          const float4 y = *reinterpret_cast<const float4*>(xs + L.wide2 + i * 4);
          const float4 k2 = *reinterpret_cast<const float4*>(wbn + 64 + c), k3 = *reinterpret_cast<const float4*>(wbn + 96 + c);
          v.x = fmaf(k0.x, v.x, fmaf(k1.x, y.x - k3.x, k2.x)); v.y = fmaf(k0.y, v.y, fmaf(k1.y, y.y - k3.y, k2.y));
          v.z = fmaf(k0.z, v.z, fmaf(k1.z, y.z - k3.z, k2.z)); v.w = fmaf(k0.w, v.w, fmaf(k1.w, y.w - k3.w, k2.w));
        }
        *reinterpret_cast<float4*>(as + i * 4) = v;
      }
    }
    __syncthreads();

    if (GATHER) {
      float acc[8][4];
#pragma unroll
      for (int px = 0; px < 8; ++px)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[px][j] = 0.f;
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          // thin columns 2*g_x0-1 .. 2*g_x0+15 of thin row 2*g_r-1+ky sit at p[3 .. 19] (column c is stored at index c + 4)
          const float4* ra = reinterpret_cast<const float4*>(xs + (c3 * XS_ROWS + 2 * g_r + ky) * XS_PITCH + 2 * g_x0);
          float p[20];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const float4 t = ra[q];
            p[4 * q + 0] = t.x; p[4 * q + 1] = t.y; p[4 * q + 2] = t.z; p[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 wv = *reinterpret_cast<const float4*>(Wsm + (c3 * 9 + ky * 3 + kx) * 32 + g_c0);
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              const float xa = p[2 * px + 3 + kx];
              fma2(acc[px][0], acc[px][1], xa, wv.x, wv.y);
              fma2(acc[px][2], acc[px][3], xa, wv.z, wv.w);
            }
          }
        }
      }
      if (e.mode == AE_EPI_RELUBWD_STATS) {
        const float4 ksc = *reinterpret_cast<const float4*>(sbn + g_c0), ksh = *reinterpret_cast<const float4*>(sbn + 32 + g_c0);
        const float4 kmu = *reinterpret_cast<const float4*>(sbn + 64 + g_c0), krs = *reinterpret_cast<const float4*>(sbn + 96 + g_c0);
        const float sc[4] = {ksc.x, ksc.y, ksc.z, ksc.w}, sh[4] = {ksh.x, ksh.y, ksh.z, ksh.w};
        const float mu[4] = {kmu.x, kmu.y, kmu.z, kmu.w}, rs[4] = {krs.x, krs.y, krs.z, krs.w};
#pragma unroll
        for (int px = 0; px < 8; ++px) {
          const float4 yq = y4[px];
          const float yv[4] = {yq.x, yq.y, yq.z, yq.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float z = fmaf(yv[j], sc[j], sh[j]);
            const float d = z > 0.f ? acc[px][j] : 0.f;
            acc[px][j] = d;
            st1[j] += d;
            st2[j] = fmaf(d, (yv[j] - mu[j]) * rs[j], st2[j]);
          }
          *reinterpret_cast<float4*>(out + row + px * WC) = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
        }
      } else if (e.mode == AE_EPI_BNRELU_SPLIT) {
        // eval mode: the next layer's tensor-core operand (split-bf16 planes of BatchNorm + ReLU) straight from the accumulators
        const float4 ksc = *reinterpret_cast<const float4*>(sbn + g_c0), ksh = *reinterpret_cast<const float4*>(sbn + 32 + g_c0);
        const float sc[4] = {ksc.x, ksc.y, ksc.z, ksc.w}, sh[4] = {ksh.x, ksh.y, ksh.z, ksh.w};
        const float bs[4] = {gbias.x, gbias.y, gbias.z, gbias.w};
        __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(out);
#pragma unroll
        for (int px = 0; px < 8; ++px) {
          float v[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = fmaxf(fmaf(acc[px][j] + bs[j], sc[j], sh[j]), 0.f);
            lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
          }
          *reinterpret_cast<uint2*>(pl + row + px * WC) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
          if (e.nsplit == 2)
            *reinterpret_cast<uint2*>(pl + plane_elems + row + px * WC) = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
        }
      } else {
        const float bs[4] = {gbias.x, gbias.y, gbias.z, gbias.w};
#pragma unroll
        for (int px = 0; px < 8; ++px) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d = acc[px][j] + bs[j];
            acc[px][j] = d;
            st1[j] += d;
            st2[j] = fmaf(d, d, st2[j]);
          }
          *reinterpret_cast<float4*>(out + row + px * WC) = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
        }
      }
    }

    if (WIDE_W) {
      const float* arow = as + warp * 32 * 32 + c4 * 4;
      const float* prow = xs + 2 * warp * XS_PITCH + 3;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int pc = kg + 4 * i;                            // kg = lane >> 3: this lane's pixel subset
        const float4 a4 = *reinterpret_cast<const float4*>(arow + pc * 32);
        const float* pb = prow + 2 * pc;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float* tp = pb + (c3 * XS_ROWS + ky) * XS_PITCH;
            const float t0 = tp[0];
            const float2 t12 = *reinterpret_cast<const float2*>(tp + 1);   // thin column 2*pc: even index, 8-byte aligned
            const int k = WIDE_W ? c3 * 9 + ky * 3 : 0;
            fma2(wfull[0][k], wfull[1][k], t0, a4.x, a4.y);             fma2(wfull[2][k], wfull[3][k], t0, a4.z, a4.w);
            fma2(wfull[0][k + 1], wfull[1][k + 1], t12.x, a4.x, a4.y);   fma2(wfull[2][k + 1], wfull[3][k + 1], t12.x, a4.z, a4.w);
            fma2(wfull[0][k + 2], wfull[1][k + 2], t12.y, a4.x, a4.y);   fma2(wfull[2][k + 2], wfull[3][k + 2], t12.y, a4.z, a4.w);
          }
      }
    } else if (WGRAD) {
      // warp q accumulates over wide pixels q*32 .. q*32+31 of the tile (= wide row q)
      const float* arow = as + warp * 32 * 32 + c4 * 4;
      const float* prow = xs + 2 * warp * XS_PITCH;
#pragma unroll 4
      for (int pc = 0; pc < 32; ++pc) {
        const float4 a4 = *reinterpret_cast<const float4*>(arow + pc * 32);
        const float* pb = prow + 2 * pc;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float t = pb[koff[j]];
          fma2(wacc[0][j], wacc[1][j], t, a4.x, a4.y);
          fma2(wacc[2][j], wacc[3][j], t, a4.z, a4.w);
        }
      }
    }
    __syncthreads();                                      // every warp is done with stage st before it is refilled
  }

  if (GATHER && e.mode != AE_EPI_STORE && e.mode != AE_EPI_BNRELU_SPLIT && e.stats) {
    // lanes that share the channel group (lane bits 3, 4 = pixel group) are summed first; lanes 0..7 then hold the warp's sums
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 8);  st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 8);
      st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 16); st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&sStat[0][g_c0 + j], st1[j]); atomicAdd(&sStat[1][g_c0 + j], st2[j]); }
    }
    __syncthreads();
    if (tid < 32) {
      atomicAdd(e.stats + tid, (double)sStat[0][tid]);
      atomicAdd(e.stats + WC + tid, (double)sStat[1][tid]);
    }
  }
  if (WGRAD) {
    float* red = stage0 + L.wide;                       // [4 warps][32 ch][28] (stage 0's wide tile: nothing in flight any more)
    if (WIDE_W) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 27; ++k) {                      // sum over the four pixel subsets (lane bits 3, 4)
          float v = wfull[i][WIDE_W ? k : 0];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (kg == 0) red[(warp * 32 + c4 * 4 + i) * 28 + k] = v;
        }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) red[(warp * 32 + c4 * 4 + i) * 28 + kg * 7 + j] = wacc[i][j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sm = warp_sum(bsum[c]);
      if (lane == 0) sB[warp][c] = sm;
    }
    __syncthreads();
    float* dst = partial + (size_t)blockIdx.x * TW_PART;
    for (int i = tid; i < 864; i += TT_THREADS) {
      const int c = i / 27, k = i - c * 27;
      const int o = c * 28 + k;
      dst[i] = (red[o] + red[32 * 28 + o]) + (red[2 * 32 * 28 + o] + red[3 * 32 * 28 + o]);
    }
    if (tid < 3) dst[864 + tid] = (sB[0][tid] + sB[1][tid]) + (sB[2][tid] + sB[3][tid]);
    if (tid == 3) dst[867] = 0.f;
  }
}

// Fixed-order two-level reduction of the per-CTA partials: block b owns 32 outputs, its 32 warps each sum a
// contiguous share of the partials, then one warp adds the 32 shares (deterministic).
__global__ void __launch_bounds__(1024) k_thin_wgrad_reduce(const float* __restrict__ partial, int nparts,
                                                            float* __restrict__ dw, float* __restrict__ dbias) {
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int per = (nparts + 31) / 32;
  const int p0 = wq * per, p1 = min(nparts, p0 + per);
  float sm = 0.f;
  if (i < 867) {
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
      const float v0 = __ldg(partial + (size_t)p * TW_PART + i), v1 = __ldg(partial + (size_t)(p + 1) * TW_PART + i);
      const float v2 = __ldg(partial + (size_t)(p + 2) * TW_PART + i), v3 = __ldg(partial + (size_t)(p + 3) * TW_PART + i);
      sm += (v0 + v1) + (v2 + v3);
    }
    for (; p < p1; ++p) sm += __ldg(partial + (size_t)p * TW_PART + i);
  }
  red[wq][lane] = sm;
  __syncthreads();
  if (wq == 0 && i < 867) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][lane];
    if (i < 864) dw[i] = t;
    else if (dbias) dbias[i - 864] = t;
  }
}

static int thin_blocks(int batch) {           // upper bound of every thin-kernel grid (sizes the partial buffer)
  const int tiles = batch * TILES_PER_IMAGE;
  return tiles < 4 * 148 ? tiles : 4 * 148;
}
size_t thin_wgrad_workspace_bytes(int batch) { return (size_t)thin_blocks(batch) * TW_PART * sizeof(float); }

// Persistent grid = what is co-resident (registers allow 4 CTAs per SM; the two-source / weight-gradient stagings are
// shared-memory limited to 2-3): a grid larger than that would run its excess CTAs as a second, nearly empty wave.
template <typename Kernel>
static int resident_blocks(Kernel kernel, size_t smem, int batch, int* blocks) {
  int dev = 0, sms = 0, occ = 0;
  AE_CUDA(cudaGetDevice(&dev));
  AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TT_THREADS, smem));
  AE_CHECK(occ >= 1, "thin kernel does not fit on an SM with %zu bytes of shared memory", smem);
  const int cap = thin_blocks(batch), res = occ * sms;
  *blocks = res < cap ? res : cap;
  return 0;
}

// [batch][3][64][64] fp32 image as a 4-d tensor map; box = (box_cols columns, box_rows rows, 3 channels, 1 image)
static int encode_image_map(CUtensorMap* map, const float* img, int batch, int box_cols = XS_PITCH, int box_rows = XS_ROWS) {
  EncodeTiledFn fn = encode_fn();
  AE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)TW, (cuuint64_t)TH, 3, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)TW * 4, (cuuint64_t)TH * TW * 4, (cuuint64_t)3 * TH * TW * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 3, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (image) failed with CUresult %d (batch %d)", (int)r, batch);
  return 0;
}

template <bool GATHER, bool WGRAD>
static int launch_thin(const Operand& thin, const Operand& wide, const float* w, const Epilogue& e, float* out,
                       float* partial, int batch, cudaStream_t st, int* blocks_out = nullptr) {
  const ThinStage L = thin_stage_layout(thin.mode, wide.mode, WGRAD);
  const size_t smem = sizeof(float) * (2 * (size_t)L.floats + 27 * 32 + 8 * 32);
  static bool attr_done = false;
  if (!attr_done) {
    // the largest layout (two-source thin operand, BatchNorm-backward wide operand)
    const ThinStage Lmax = thin_stage_layout(AE_OP_SIGMOID_BWD, AE_OP_BNBWD, WGRAD);
    AE_CUDA(cudaFuncSetAttribute(k_thin<GATHER, WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * (2 * (size_t)Lmax.floats + 27 * 32 + 8 * 32))));
    attr_done = true;
  }
  int blocks = 0;
  AE_TRY(resident_blocks(k_thin<GATHER, WGRAD>, smem, batch, &blocks));
  if (blocks_out) *blocks_out = blocks;
  CUtensorMap xmap, xmap2;
  AE_TRY(encode_image_map(&xmap, thin.src, batch));
  if (thin.mode != AE_OP_RAW) AE_TRY(encode_image_map(&xmap2, thin.src2, batch));
  else xmap2 = xmap;
  k_thin<GATHER, WGRAD><<<blocks, TT_THREADS, smem, st>>>(thin, wide, w, e, out, partial, batch, xmap, xmap2);
  AE_LAUNCH_CHECK();
  return 0;
}

static int check_wide_operand(const Operand& wide) {
  AE_CHECK(wide.mode == AE_OP_RAW || wide.mode == AE_OP_BNRELU || wide.mode == AE_OP_BNBWD, "wide operand: mode %d not supported", wide.mode);
  AE_CHECK(((uintptr_t)wide.src & 15) == 0 && (wide.mode != AE_OP_BNBWD || ((uintptr_t)wide.src2 & 15) == 0),
           "wide operand: tensors must be 16-byte aligned");
  return 0;
}

static int check_thin_operand(const Operand& thin) {
  AE_CHECK(thin.mode == AE_OP_RAW || thin.mode == AE_OP_SIGMOID_BWD, "thin operand: mode %d not supported", thin.mode);
  AE_CHECK(((uintptr_t)thin.src & 15) == 0 && (thin.mode == AE_OP_RAW || ((uintptr_t)thin.src2 & 15) == 0),
           "thin operand: image tensors must be 16-byte aligned");
  return 0;
}

int thin_gather_fwd(const Operand& thin, const float* w, const Epilogue& epi, float* out, int batch, cudaStream_t st) {
  AE_TRY(check_thin_operand(thin));
  return launch_thin<true, false>(thin, raw_operand(nullptr), w, epi, out, nullptr, batch, st);
}

int thin_wgrad(const Operand& wide, const Operand& thin, float* dw, float* dbias, void* partials, size_t bytes,
               int batch, cudaStream_t st) {
  AE_TRY(check_thin_operand(thin));
  AE_TRY(check_wide_operand(wide));
  int blocks = 0;
  AE_CHECK(bytes >= (size_t)thin_blocks(batch) * TW_PART * sizeof(float), "thin_wgrad: workspace too small");
  AE_TRY((launch_thin<false, true>(thin, wide, nullptr, store_epilogue(), nullptr, static_cast<float*>(partials), batch, st, &blocks)));
  k_thin_wgrad_reduce<<<(867 + 31) / 32, 1024, 0, st>>>(static_cast<const float*>(partials), blocks, dw, dbias);
  AE_LAUNCH_CHECK();
  return 0;
}

// Fused backward of the thin transposed convolution (NB:628): weight / bias gradient and the data gradient
// (gather with the ReLU-backward epilogue) from one pass over the thin operand.
// phase 0: both launches on `st`; phase 1: the fused kernel only; phase 2: only the fixed-order reduce of its per-CTA
// partials (the engine runs it on a parallel graph branch: nothing on the data-gradient path reads dw / dbias).
int thin_bwd_fused(const Operand& wide, const Operand& thin, const float* w, const Epilogue& epi, float* out_wide, float* dw,
                   float* dbias, void* partials, size_t bytes, int batch, cudaStream_t st, int phase) {
  AE_TRY(check_thin_operand(thin));
  AE_TRY(check_wide_operand(wide));
  int blocks = 0;
  AE_CHECK(bytes >= (size_t)thin_blocks(batch) * TW_PART * sizeof(float), "thin_bwd_fused: workspace too small");
  if (phase == 0 || phase == 1) {
    AE_TRY((launch_thin<true, true>(thin, wide, w, epi, out_wide, static_cast<float*>(partials), batch, st, &blocks)));
  } else {
    const ThinStage L = thin_stage_layout(thin.mode, wide.mode, true);
    AE_TRY(resident_blocks(k_thin<true, true>, sizeof(float) * (2 * (size_t)L.floats + 27 * 32 + 8 * 32), batch, &blocks));
  }
  if (phase == 0 || phase == 2) {
    k_thin_wgrad_reduce<<<(867 + 31) / 32, 1024, 0, st>>>(static_cast<const float*>(partials), blocks, dw, dbias);
    AE_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// wide -> thin scatter + sigmoid (+ squared error): convT4 forward.  The transformed wide tile (5 rows x 33
// columns with a zero halo) is staged once; four lanes share a strip of 4 wide pixels, each summing a quarter of the
// input channels for all 4 output quads, then a register reduce-scatter hands every lane one pixel's 2x2x3 quad
// (no atomics on the output).
// ---------------------------------------------------------------------------------------------
static constexpr int AS_ROWS = TILE_ROWS + 1, AS_COLS = WW + 1;

__global__ void __launch_bounds__(TT_THREADS, 4) k_thin_scatter_sigmoid(Operand wide, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, float* __restrict__ x_hat,
                                                                     const float* __restrict__ x, double* __restrict__ sse,
                                                                     int batch, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) float smem_f[];
  float* raw = smem_f;                                  // [AS_ROWS][32][32]: the tile's wide rows + the halo row, one bulk copy
  float* as = raw + AS_ROWS * WW * WC;                  // [AS_ROWS][AS_COLS][32] transformed, 16-byte chunk c of pixel p at c ^ (p & 7)
  float* Wsm = as + AS_ROWS * AS_COLS * 32;             // [9 taps][3 co][32 ci]
  float* sbn = Wsm + 27 * 32;                           // [2][32] scale, shift of the wide operand
  float* xt = sbn + 64;                                 // [3][8][64]: the tile's rows of the target image (one TMA box)
  __shared__ __align__(8) uint64_t bar_raw, bar_x;
  __shared__ float red[TT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&bar_raw), barx = smem_u32(&bar_x);
  const int tiles = batch * TILES_PER_IMAGE;
  constexpr int UNITS = AS_ROWS * AS_COLS * 8;
  // one thread: fetch the raw rows of `tile` (the image's last tile has no halo row below it)
  auto issue = [&](int tile) {
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const uint32_t bytes = (uint32_t)((tr == TILES_PER_IMAGE - 1 ? TILE_ROWS : AS_ROWS) * WW * WC * 4);
    mbar_arrive_expect_tx(bar, bytes);
    bulk_copy_g2s(smem_u32(raw), wide.src + ((size_t)n * WH + tr * TILE_ROWS) * WW * WC, bytes, bar);
  };
  if (tid == 0) {
    if (x) tma_prefetch_desc(&tmap);
    mbar_init(bar, 1); mbar_init(barx, 1);
    fence_barrier_init();
    if ((int)blockIdx.x < tiles) issue(blockIdx.x);     // on its way while the weights are staged
  }
  for (int o = tid; o < 27 * 32; o += TT_THREADS) {     // o = (tap*3 + co)*32 + ci  <-  w[ci][co][tap]
    const int ci = o & 31, r = o >> 5, tap = r / 3, co = r - tap * 3;
    Wsm[o] = __ldg(w + ci * 27 + co * 9 + tap);
  }
  if (tid < 64)
    sbn[tid] = wide.mode == AE_OP_BNRELU ? __ldg(wide.bnc + (tid >> 5 ? AE_BNC_SHIFT : AE_BNC_SCALE) * WC + lane) : (tid >> 5 ? 0.f : 1.f);
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  float err = 0.f;
  __syncthreads();
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    mbar_wait(bar, it & 1);
    // transform pass: raw -> BatchNorm + ReLU -> swizzled tile with a zero halo (right column, row below the image)
    for (int i = tid; i < UNITS; i += TT_THREADS) {
      const int q = i & 7, p = i >> 3;
      const int pr = p / AS_COLS, pc = p - pr * AS_COLS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tr * TILE_ROWS + pr < WH && pc < WW) {
        v = *reinterpret_cast<const float4*>(raw + (pr * WW + pc) * WC + q * 4);
        if (wide.mode == AE_OP_BNRELU) {
          const float4 sc = *reinterpret_cast<const float4*>(sbn + q * 4), sh = *reinterpret_cast<const float4*>(sbn + 32 + q * 4);
          v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
          v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
        }
      }
      *reinterpret_cast<float4*>(as + p * 32 + ((q ^ (p & 7)) << 2)) = v;
    }
    __syncthreads();                                    // `as` complete, `raw` consumed
    if (tid == 0) {
      fence_proxy_async();
      if (tile + (int)gridDim.x < tiles) issue(tile + gridDim.x);   // the next tile's rows arrive while this one is computed
      if (x) {                                          // ... and so do this tile's rows of the target image (squared error)
        mbar_arrive_expect_tx(barx, 3 * 2 * TILE_ROWS * TW * 4);
        tma_load_4d(smem_u32(xt), &tmap, 0, 2 * TILE_ROWS * tr, 0, n, barx);
      }
    }
    // Thread (warp = tile row, strip = lane >> 2, ciq = lane & 3): the 2x2x3 output quads of the strip's 4 wide pixels,
    // summed over input channels {ciq*4 .. +3} and {16 + ciq*4 .. +3}; one weight load feeds 16 FMAs.  A two-stage
    // reduce-scatter over the 4 ciq lanes then leaves lane ciq with the complete sums of pixel ciq of the strip.
    const int x0 = (lane >> 2) * 4, ciq = lane & 3;
    float acc[4][12];                                   // [pixel of the strip][(py*2 + px)*3 + co]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 12; ++b) acc[a][b] = 0.f;
    // source pixel (iy+dy, ix+dx) contributes to output parity (py,px) through tap (ky,kx):
    //   dy=0: py=0 -> ky=1 ; py=1 -> ky=2        dy=1: py=1 -> ky=0      (same for x)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int q4 = ciq + 4 * half;
      float4 v[2][5];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int p = (warp + dy) * AS_COLS + x0 + c;
          v[dy][c] = *reinterpret_cast<const float4*>(as + p * 32 + ((q4 ^ (p & 7)) << 2));
        }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int py = dy; py < 2; ++py) {
            const int ky = dy ? 0 : (py ? 2 : 1);
#pragma unroll
            for (int px = dx; px < 2; ++px) {
              const int kx = dx ? 0 : (px ? 2 : 1);
#pragma unroll
              for (int co = 0; co < 3; ++co) {
                const float4 wv = *reinterpret_cast<const float4*>(Wsm + ((ky * 3 + kx) * 3 + co) * 32 + q4 * 4);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                  const float4 vv = v[dy][a + dx];
                  float& t = acc[a][(py * 2 + px) * 3 + co];
                  t = fmaf(vv.x, wv.x, fmaf(vv.y, wv.y, fmaf(vv.z, wv.z, fmaf(vv.w, wv.w, t))));
                }
              }
            }
          }
    }
    float r1[2][12], r2[12];
    {
      const bool hi = (ciq & 2) != 0;                   // stage 1 (partner lane ^ 2): keep pixels {0,1} or {2,3}
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 12; ++b) {
          const float send = hi ? acc[a][b] : acc[a + 2][b];
          const float keep = hi ? acc[a + 2][b] : acc[a][b];
          r1[a][b] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      const bool odd = (ciq & 1) != 0;                  // stage 2 (partner lane ^ 1): keep the first or the second of the pair
#pragma unroll
      for (int b = 0; b < 12; ++b) {
        const float send = odd ? r1[0][b] : r1[1][b];
        const float keep = odd ? r1[1][b] : r1[0][b];
        r2[b] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
    }
    const int ty0 = 2 * (tr * TILE_ROWS + warp);        // first of this thread's two thin rows; its wide column is x0 + ciq
    if (x) mbar_wait(barx, it & 1);
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      const float b = co == 0 ? b0 : (co == 1 ? b1 : b2);
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const size_t o = (((size_t)n * 3 + co) * TH + ty0 + py) * TW + 2 * (x0 + ciq);
        const float s0 = 1.f / (1.f + expf(-(r2[(py * 2 + 0) * 3 + co] + b)));
        const float s1 = 1.f / (1.f + expf(-(r2[(py * 2 + 1) * 3 + co] + b)));
        *reinterpret_cast<float2*>(x_hat + o) = make_float2(s0, s1);
        if (x) {
          const float2 t = *reinterpret_cast<const float2*>(xt + (co * 2 * TILE_ROWS + 2 * warp + py) * TW + 2 * (x0 + ciq));
          err += (s0 - t.x) * (s0 - t.x) + (s1 - t.y) * (s1 - t.y);
        }
      }
    }
    __syncthreads();                                    // every warp is done with `as` before the next tile's transform
  }
  if (x && sse) {
    err = warp_sum(err);
    if (lane == 0) red[warp] = err;
    __syncthreads();
    if (tid == 0) atomicAdd(sse, ((double)red[0] + (double)red[1]) + ((double)red[2] + (double)red[3]));
  }
}

int thin_scatter_sigmoid_fwd(const Operand& wide, const float* w, const float* bias, float* x_hat, const float* x,
                             double* sse, int batch, cudaStream_t st) {
  AE_CHECK(wide.mode == AE_OP_RAW || wide.mode == AE_OP_BNRELU, "thin_scatter: wide operand mode %d not supported", wide.mode);
  AE_CHECK(((uintptr_t)wide.src & 15) == 0, "thin_scatter: the wide tensor must be 16-byte aligned");
  const size_t smem = sizeof(float) * (AS_ROWS * WW * WC + AS_ROWS * AS_COLS * 32 + 27 * 32 + 64 + 3 * 2 * TILE_ROWS * TW);
  static bool attr_done = false;
  if (!attr_done) {
    AE_CUDA(cudaFuncSetAttribute(k_thin_scatter_sigmoid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int blocks = 0;
  AE_TRY(resident_blocks(k_thin_scatter_sigmoid, smem, batch, &blocks));
  CUtensorMap tmap;
  if (x) {
    AE_CHECK(((uintptr_t)x & 15) == 0, "thin_scatter: the target image must be 16-byte aligned");
    AE_TRY(encode_image_map(&tmap, x, batch, TW, 2 * TILE_ROWS));
  } else {
    memset(&tmap, 0, sizeof(tmap));
  }
  k_thin_scatter_sigmoid<<<blocks, TT_THREADS, smem, st>>>(wide, w, bias, x_hat, x, sse, batch, tmap);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

// The two 3-channel layers on the tensor cores (AE_BACKEND_TC): Conv2d(3,32,k3,s2,p1) (NB:504) and
// ConvTranspose2d(32,3,k3,s2,p1,op1)+Sigmoid (NB:628-629), forward and backward.
//
// K = 27 and N = 3 are far too small to feed tcgen05 from TMA-shaped operands, but the CUDA-core versions (thin.cu)
// spend ~1100 instructions per pixel on FMAs and shared-memory weight reads.  Here the 128 threads of a CTA only BUILD
// bf16 (hi / lo) operand tiles in shared memory from the TMA-staged raw rows -- the 27-tap patch matrix P[128 px][32]
// and / or the transformed wide tile Wd[128 px][32] -- and one thread issues the MMAs:
//   gather  (conv1 fwd, convT4 dgrad): D1[128 px][32 c] = P (K-major) * W^T          2 k-steps
//   wgrad   (both layers)            : D2[32 c][32 k] += Wd^T (MN-major) * P (MN-major)   8 k-steps, accumulated in
//                                       tensor memory over every tile of the CTA, read out once at the end
//   scatter (convT4 fwd)             : D[128 px][16 = 4 phases x 3 co (+4 pad)] = [a(y,x) a(y,x+1) a(y+1,x) a(y+1,x+1)] * Wz^T
//                                       where Wz holds the tap weight of every (neighbour, phase) pair that is connected
// All operand tiles are SWIZZLE_64B (rows of 32 bf16).  fp32 mode: hi*hi + hi*lo + lo*hi.
#include "thin_common.cuh"

namespace ae {

// byte offset of 16-byte chunk c (0..3) of row r in a SWIZZLE_64B tile (rows of 64 bytes)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void split4(float a, float b, float c, float d, uint2& hi, uint2& lo) {
  hi.x = pack_bf16x2(a, b); hi.y = pack_bf16x2(c, d);
  lo.x = pack_bf16x2(a - __bfloat162float(__float2bfloat16_rn(a)), b - __bfloat162float(__float2bfloat16_rn(b)));
  lo.y = pack_bf16x2(c - __bfloat162float(__float2bfloat16_rn(c)), d - __bfloat162float(__float2bfloat16_rn(d)));
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

static constexpr int TILE8K = 128 * 64;      // bytes of one [128 rows][64 B] tile plane

// ---------------------------------------------------------------------------------------------
// gather and / or weight gradient
// ---------------------------------------------------------------------------------------------
template <bool GATHER, bool WGRAD, int NSPLIT>
__global__ void __launch_bounds__(TT_THREADS) k_thin_tc(Operand thin, Operand wide, const float* __restrict__ w, Epilogue e,
                                                        float* __restrict__ out, float* __restrict__ partial, int batch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Pt = sm;                                         // [NSPLIT][128][64 B]  patch matrix
  uint8_t* Wd = Pt + NSPLIT * TILE8K;                       // [NSPLIT][128][64 B]  transformed wide tile (WGRAD)
  uint8_t* Wt = Wd + (WGRAD ? NSPLIT * TILE8K : 0);         // [NSPLIT][32][64 B]   weights, rows = c32, K = tap (GATHER)
  float* sbn = reinterpret_cast<float*>(Wt + (GATHER ? NSPLIT * 2048 : 0));   // [4][32] epilogue coefficients
  float* wbn = sbn + 128;                                   // [4][32] coefficients of the wide operand
  float* stage0 = wbn + 128;                                // two raw stages (thin_stage_layout)
  const ThinStage L = thin_stage_layout(thin.mode, wide.mode, WGRAD);
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  __shared__ float sStat[2][32];
  __shared__ float sB[TT_THREADS / 32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t mma_bar = bar0 + 16;

  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  if (GATHER) {
    // weights [c32][k = c3*9 + tap], k padded to 32: row c32, bf16 hi / lo
    for (int i = tid; i < 32 * 4; i += TT_THREADS) {
      const int c32 = i >> 2, ch = i & 3;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int k = ch * 8 + j; v[j] = k < 27 ? __ldg(w + c32 * 27 + k) : 0.f; }
      uint2 h0, l0, h1, l1;
      split4(v[0], v[1], v[2], v[3], h0, l0);
      split4(v[4], v[5], v[6], v[7], h1, l1);
      *reinterpret_cast<uint4*>(Wt + sw64_off(c32, ch)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
      if (NSPLIT == 2) *reinterpret_cast<uint4*>(Wt + 2048 + sw64_off(c32, ch)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
    }
    if (e.mode == AE_EPI_RELUBWD_STATS) {
      const int rows[4] = {AE_BNC_SCALE, AE_BNC_SHIFT, AE_BNC_MEAN, AE_BNC_RSTD};
      sbn[tid] = __ldg(e.bnc + rows[tid >> 5] * WC + lane);
    } else {
      sbn[tid] = (tid < 32 && e.bias) ? __ldg(e.bias + tid) : 0.f;
    }
    if (tid < 32) { sStat[0][tid] = 0.f; sStat[1][tid] = 0.f; }
  }
  if (WGRAD && wide.mode != AE_OP_RAW) {
    const int rows_relu[4] = {AE_BNC_SCALE, AE_BNC_SHIFT, AE_BNC_SCALE, AE_BNC_SHIFT};
    const int rows_bwd[4] = {AE_BNC_A, AE_BNC_B, AE_BNC_C, AE_BNC_MEAN};
    wbn[tid] = __ldg(wide.bnc + (wide.mode == AE_OP_BNRELU ? rows_relu[tid >> 5] : rows_bwd[tid >> 5]) * WC + lane);
  }
  for (int i = tid; i < 2 * 3 * XS_ROWS; i += TT_THREADS) {   // left zero padding of both stages, never overwritten
    const int st = i / (3 * XS_ROWS), r = i - st * 3 * XS_ROWS;
    stage0[st * L.floats + r * XS_PITCH + 3] = 0.f;
  }
  float st1 = 0.f, st2 = 0.f;
  float bsum[3] = {0.f, 0.f, 0.f};
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t d1 = tmem_base, d2 = tmem_base + 32;

  const int tiles = batch * TILES_PER_IMAGE;
  auto issue = [&](int tile, int st) {
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const uint32_t bar = bar0 + 8u * st;
    float* base = stage0 + st * L.floats;
    const int r_first = tr == 0 ? 1 : 0;
    const int nsrc = thin.mode != AE_OP_RAW ? 2 : 1;
    uint32_t bytes = (uint32_t)(3 * (XS_ROWS - r_first) * TW * 4 * nsrc);
    if (WGRAD) bytes += WT_BYTES * (wide.mode == AE_OP_BNBWD ? 2 : 1);
    mbar_arrive_expect_tx(bar, bytes);
    for (int c3 = 0; c3 < 3; ++c3)
      for (int r = r_first; r < XS_ROWS; ++r) {
        const size_t off = (((size_t)n * 3 + c3) * TH + (2 * TILE_ROWS * tr - 1 + r)) * TW;
        const int so = (c3 * XS_ROWS + r) * XS_PITCH + 4;
        bulk_copy_g2s(smem_u32(base + so), thin.src + off, TW * 4, bar);
        if (nsrc == 2) bulk_copy_g2s(smem_u32(base + L.xs2 + so), thin.src2 + off, TW * 4, bar);
      }
    if (WGRAD) {
      const size_t m0 = ((size_t)n * WH + tr * TILE_ROWS) * WW;
      bulk_copy_g2s(smem_u32(base + L.wide), wide.src + m0 * WC, WT_BYTES, bar);
      if (wide.mode == AE_OP_BNBWD) bulk_copy_g2s(smem_u32(base + L.wide2), wide.src2 + m0 * WC, WT_BYTES, bar);
    }
  };

  if (tid == 0 && (int)blockIdx.x < tiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int st = it & 1;
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const size_t m0 = ((size_t)n * WH + tr * TILE_ROWS) * WW;
    float* xs = stage0 + st * L.floats;
    const float* wraw = xs + L.wide;
    if (tid == 0 && tile + (int)gridDim.x < tiles) {
      fence_proxy_async();
      issue(tile + gridDim.x, st ^ 1);
    }
    float4 y4[8];                                         // RELUBWD: raw output row of this thread's pixel (prefetched)
    const size_t row = (m0 + (size_t)tid) * WC;
    if (GATHER && e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) y4[q4] = __ldg(reinterpret_cast<const float4*>(e.y + row) + q4);
    }
    mbar_wait(bar0 + 8u * st, (it >> 1) & 1);
    // ---- thin rows: transform in place (once per element), zero the missing halo row, bias-gradient sums ----
    if (thin.mode != AE_OP_RAW || tr == 0 || WGRAD) {
      for (int i = tid; i < 3 * XS_ROWS * 16; i += TT_THREADS) {
        const int q = i & 15, r = (i >> 4) % XS_ROWS, c3 = i / (16 * XS_ROWS);
        float* px = xs + (c3 * XS_ROWS + r) * XS_PITCH + 4 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r > 0 || tr > 0) {
          v = *reinterpret_cast<const float4*>(px);
          if (thin.mode != AE_OP_RAW) {
            const float4 sg = *reinterpret_cast<const float4*>(px + L.xs2);
            v.x = thin_transform(thin, v.x, sg.x); v.y = thin_transform(thin, v.y, sg.y);
            v.z = thin_transform(thin, v.z, sg.z); v.w = thin_transform(thin, v.w, sg.w);
          }
          if (WGRAD && r >= 1) bsum[c3 == 0 ? 0 : (c3 == 1 ? 1 : 2)] += (v.x + v.y) + (v.z + v.w);
        }
        if (thin.mode != AE_OP_RAW || (r == 0 && tr == 0)) *reinterpret_cast<float4*>(px) = v;
      }
      __syncthreads();
    }
    // ---- patch matrix P: row = this thread's wide pixel, 27 taps (+5 zero) ----
    {
      const int x = lane, r = warp;
      float v[32];
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* ra = xs + (c3 * XS_ROWS + 2 * r + ky) * XS_PITCH + 3 + 2 * x;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) v[c3 * 9 + ky * 3 + kx] = ra[kx];
        }
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint2 h0, l0, h1, l1;
        split4(v[ch * 8 + 0], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], h0, l0);
        split4(v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7], h1, l1);
        *reinterpret_cast<uint4*>(Pt + sw64_off(tid, ch)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        if (NSPLIT == 2) *reinterpret_cast<uint4*>(Pt + TILE8K + sw64_off(tid, ch)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
    }
    // ---- transformed wide tile Wd: 128 pixels x 8 float4 units ----
    if (WGRAD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * TT_THREADS;
        const int p = i >> 3, q = i & 7, c = q * 4;
        float4 v = *reinterpret_cast<const float4*>(wraw + i * 4);
        if (wide.mode != AE_OP_RAW) {
          const float4 k0 = *reinterpret_cast<const float4*>(wbn + c), k1 = *reinterpret_cast<const float4*>(wbn + 32 + c);
          if (wide.mode == AE_OP_BNRELU) {
            v.x = fmaxf(fmaf(v.x, k0.x, k1.x), 0.f); v.y = fmaxf(fmaf(v.y, k0.y, k1.y), 0.f);
            v.z = fmaxf(fmaf(v.z, k0.z, k1.z), 0.f); v.w = fmaxf(fmaf(v.w, k0.w, k1.w), 0.f);
          } else {                                        // dy = A*dz + B*(y - mean) + C   (same order as load_operand4)
            const float4 y = *reinterpret_cast<const float4*>(xs + L.wide2 + i * 4);
            const float4 k2 = *reinterpret_cast<const float4*>(wbn + 64 + c), k3 = *reinterpret_cast<const float4*>(wbn + 96 + c);
            v.x = fmaf(k0.x, v.x, fmaf(k1.x, y.x - k3.x, k2.x)); v.y = fmaf(k0.y, v.y, fmaf(k1.y, y.y - k3.y, k2.y));
            v.z = fmaf(k0.z, v.z, fmaf(k1.z, y.z - k3.z, k2.z)); v.w = fmaf(k0.w, v.w, fmaf(k1.w, y.w - k3.w, k2.w));
          }
        }
        uint2 h, l;
        split4(v.x, v.y, v.z, v.w, h, l);
        const uint32_t o = sw64_off(p, q >> 1) + (uint32_t)(q & 1) * 8u;
        *reinterpret_cast<uint2*>(Wd + o) = h;
        if (NSPLIT == 2) *reinterpret_cast<uint2*>(Wd + TILE8K + o) = l;
      }
    }
    fence_proxy_async();                                  // the tiles were written through the generic proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t p0 = smem_u32(Pt), w0 = smem_u32(Wt), a0 = smem_u32(Wd);
      if (GATHER) {
        constexpr uint32_t idesc = make_idesc(32, 0, 0);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t ah = make_desc(p0 + kk * 32, 16, 512, 64), bh = make_desc(w0 + kk * 32, 16, 512, 64);
          umma_bf16(d1, ah, bh, idesc, kk != 0);
          if (NSPLIT == 2) {
            const uint64_t al = make_desc(p0 + TILE8K + kk * 32, 16, 512, 64), bl = make_desc(w0 + 2048 + kk * 32, 16, 512, 64);
            umma_bf16(d1, ah, bl, idesc, 1);
            umma_bf16(d1, al, bh, idesc, 1);
          }
        }
      }
      if (WGRAD) {
        // D2[m = c32 (rows 32..127 alias rows 0..31: leading byte offset 0)][n = tap] += sum over the 128 pixels
        constexpr uint32_t idesc = make_idesc(32, 1, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ah = make_desc(a0 + kk * 1024, 0, 512, 64), bh = make_desc(p0 + kk * 1024, 0, 512, 64);
          umma_bf16(d2, ah, bh, idesc, (it | kk) != 0);
          if (NSPLIT == 2) {
            const uint64_t al = make_desc(a0 + TILE8K + kk * 1024, 0, 512, 64), bl = make_desc(p0 + TILE8K + kk * 1024, 0, 512, 64);
            umma_bf16(d2, ah, bl, idesc, 1);
            umma_bf16(d2, al, bh, idesc, 1);
          }
        }
      }
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, it & 1);                           // operand tiles are free again; D1 is complete
    tc_fence_after();
    if (GATHER) {
      float acc[32];
      tmem_ld32(d1 + ((uint32_t)(warp * 32) << 16), acc);
      float s2v[32];
      if (e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float yv[4] = {y4[q4].x, y4[q4].y, y4[q4].z, y4[q4].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = q4 * 4 + j;
            const float z = fmaf(yv[j], sbn[c], sbn[32 + c]);
            const float d = z > 0.f ? acc[c] : 0.f;
            acc[c] = d;
            s2v[c] = d * ((yv[j] - sbn[64 + c]) * sbn[96 + c]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float d = acc[c] + sbn[c];
          acc[c] = d;
          s2v[c] = d * d;
        }
      }
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) st_global_v8(out + row + j8 * 8, acc, j8);   // full 32-byte sectors
      if (e.mode != AE_EPI_STORE && e.stats) {
        st1 += warp_colsum32(acc, lane);
        st2 += warp_colsum32(s2v, lane);
      }
      tc_fence_before();
    }
    __syncthreads();                                      // stage st and D1 may be reused
  }

  if (GATHER && e.mode != AE_EPI_STORE && e.stats) {
    atomicAdd(&sStat[0][lane], st1);
    atomicAdd(&sStat[1][lane], st2);
    __syncthreads();
    if (tid < 32) {
      atomicAdd(e.stats + tid, (double)sStat[0][tid]);
      atomicAdd(e.stats + WC + tid, (double)sStat[1][tid]);
    }
  }
  if (WGRAD) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sm_ = warp_sum(bsum[c]);
      if (lane == 0) sB[warp][c] = sm_;
    }
    __syncthreads();
    float* dst = partial + (size_t)blockIdx.x * TW_PART;
    if (warp == 0) {                                      // TMEM lanes 0..31 = c32, columns = tap
      float v[32];
      if (it > 0) { tc_fence_after(); tmem_ld32(d2, v); }
      else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 27; ++k) dst[lane * 27 + k] = v[k];
    }
    if (tid < 3) dst[864 + tid] = (sB[0][tid] + sB[1][tid]) + (sB[2][tid] + sB[3][tid]);
    if (tid == 3) dst[867] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// same two-level fixed-order reduction as thin.cu
__global__ void __launch_bounds__(1024) k_thin_tc_wgrad_reduce(const float* __restrict__ partial, int nparts,
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

static int tc_blocks(int batch) {
  const int tiles = batch * TILES_PER_IMAGE;
  return tiles < 4 * 148 ? tiles : 4 * 148;
}
size_t thin_tc_wgrad_workspace_bytes(int batch) { return (size_t)tc_blocks(batch) * TW_PART * sizeof(float); }

template <bool GATHER, bool WGRAD, int NSPLIT>
static int launch_thin_tc(const Operand& thin, const Operand& wide, const float* w, const Epilogue& e, float* out,
                          float* partial, int batch, cudaStream_t st) {
  auto bytes_for = [](int thin_mode, int wide_mode) {
    const ThinStage L = thin_stage_layout(thin_mode, wide_mode, WGRAD);
    return (size_t)1024 + NSPLIT * TILE8K + (WGRAD ? NSPLIT * TILE8K : 0) + (GATHER ? NSPLIT * 2048 : 0) + 256 * 4 +
           2 * (size_t)L.floats * 4;
  };
  static bool attr_done = false;
  if (!attr_done) {
    AE_CUDA(cudaFuncSetAttribute(k_thin_tc<GATHER, WGRAD, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)bytes_for(AE_OP_SIGMOID_BWD, AE_OP_BNBWD)));
    attr_done = true;
  }
  k_thin_tc<GATHER, WGRAD, NSPLIT><<<tc_blocks(batch), TT_THREADS, bytes_for(thin.mode, wide.mode), st>>>(thin, wide, w, e, out,
                                                                                                            partial, batch);
  AE_LAUNCH_CHECK();
  return 0;
}

static int check_ops(const Operand& thin, const Operand* wide) {
  AE_CHECK(thin.mode == AE_OP_RAW || thin.mode == AE_OP_SIGMOID_BWD, "thin operand: mode %d not supported", thin.mode);
  AE_CHECK(((uintptr_t)thin.src & 15) == 0 && (thin.mode == AE_OP_RAW || ((uintptr_t)thin.src2 & 15) == 0),
           "thin operand: image tensors must be 16-byte aligned");
  if (wide) {
    AE_CHECK(wide->mode == AE_OP_RAW || wide->mode == AE_OP_BNRELU || wide->mode == AE_OP_BNBWD, "wide operand: mode %d not supported", wide->mode);
    AE_CHECK(((uintptr_t)wide->src & 15) == 0 && (wide->mode != AE_OP_BNBWD || ((uintptr_t)wide->src2 & 15) == 0),
             "wide operand: tensors must be 16-byte aligned");
  }
  return 0;
}

int thin_tc_gather_fwd(const Operand& thin, const float* w, const Epilogue& epi, float* out, int batch, int nsplit, cudaStream_t st) {
  AE_TRY(check_ops(thin, nullptr));
  AE_CHECK(epi.mode != AE_EPI_BNRELU_SPLIT, "thin_tc_gather_fwd: the split-bf16 epilogue is implemented by the CUDA-core kernel only");
  const Operand none = raw_operand(nullptr);
  return nsplit == 2 ? launch_thin_tc<true, false, 2>(thin, none, w, epi, out, nullptr, batch, st)
                     : launch_thin_tc<true, false, 1>(thin, none, w, epi, out, nullptr, batch, st);
}

int thin_tc_wgrad(const Operand& wide, const Operand& thin, float* dw, float* dbias, void* partials, size_t bytes, int batch,
                  int nsplit, cudaStream_t st) {
  AE_TRY(check_ops(thin, &wide));
  const int blocks = tc_blocks(batch);
  AE_CHECK(bytes >= (size_t)blocks * TW_PART * sizeof(float), "thin_tc_wgrad: workspace too small");
  float* part = static_cast<float*>(partials);
  AE_TRY(nsplit == 2 ? (launch_thin_tc<false, true, 2>(thin, wide, nullptr, store_epilogue(), nullptr, part, batch, st))
                     : (launch_thin_tc<false, true, 1>(thin, wide, nullptr, store_epilogue(), nullptr, part, batch, st)));
  k_thin_tc_wgrad_reduce<<<(867 + 31) / 32, 1024, 0, st>>>(part, blocks, dw, dbias);
  AE_LAUNCH_CHECK();
  return 0;
}

int thin_tc_bwd_fused(const Operand& wide, const Operand& thin, const float* w, const Epilogue& epi, float* out_wide, float* dw,
                      float* dbias, void* partials, size_t bytes, int batch, int nsplit, cudaStream_t st) {
  AE_TRY(check_ops(thin, &wide));
  const int blocks = tc_blocks(batch);
  AE_CHECK(bytes >= (size_t)blocks * TW_PART * sizeof(float), "thin_tc_bwd_fused: workspace too small");
  float* part = static_cast<float*>(partials);
  AE_TRY(nsplit == 2 ? (launch_thin_tc<true, true, 2>(thin, wide, w, epi, out_wide, part, batch, st))
                     : (launch_thin_tc<true, true, 1>(thin, wide, w, epi, out_wide, part, batch, st)));
  k_thin_tc_wgrad_reduce<<<(867 + 31) / 32, 1024, 0, st>>>(part, blocks, dw, dbias);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// scatter + sigmoid (+ squared error): convT4 forward
// ---------------------------------------------------------------------------------------------
static constexpr int SC_ROWS = TILE_ROWS + 1;             // wide rows per tile incl. the bottom halo row
static constexpr int SC_PIX = SC_ROWS * WW;               // 160
static constexpr int SC_COPY = SC_PIX * 64;               // bytes of one plane of one column-shift copy

template <int NSPLIT>
__global__ void __launch_bounds__(TT_THREADS) k_thin_tc_scatter(Operand wide, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ x_hat,
                                                                const float* __restrict__ x, double* __restrict__ sse, int batch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // A copies: [dx][plane][160 rows][64 B]: row p = wide pixel (pr, pc) of the tile holds a(pr, pc + dx) (zero past the edge)
  uint8_t* Ac = sm;
  uint8_t* Bz = Ac + 2 * NSPLIT * SC_COPY;                  // [plane][4 neighbours][16 rows][64 B]
  float* wbn = reinterpret_cast<float*>(Bz + NSPLIT * 4 * 1024);
  float* stage0 = wbn + 64;                                 // two raw stages of SC_PIX * 32 floats
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  __shared__ float red[TT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t mma_bar = bar0 + 16;
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
  // Bz[(py,px,co)][(dy,dx), ci]: tap weight if output parity (py,px) sees neighbour (dy,dx), else 0
  for (int i = tid; i < 16 * 4 * 4; i += TT_THREADS) {
    const int ch = i & 3, nb = (i >> 2) & 3, nrow = i >> 4;
    const int dy = nb >> 1, dx = nb & 1;
    const int ph = nrow / 3, co = nrow - ph * 3, py = ph >> 1, px = ph & 1;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ci = ch * 8 + j;
      float t = 0.f;
      if (nrow < 12 && py >= dy && px >= dx) {
        const int ky = dy ? 0 : (py ? 2 : 1), kx = dx ? 0 : (px ? 2 : 1);
        t = __ldg(w + ci * 27 + co * 9 + ky * 3 + kx);
      }
      v[j] = t;
    }
    uint2 h0, l0, h1, l1;
    split4(v[0], v[1], v[2], v[3], h0, l0);
    split4(v[4], v[5], v[6], v[7], h1, l1);
    *reinterpret_cast<uint4*>(Bz + nb * 1024 + sw64_off(nrow, ch)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    if (NSPLIT == 2) *reinterpret_cast<uint4*>(Bz + 4 * 1024 + nb * 1024 + sw64_off(nrow, ch)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
  }
  if (tid < 64) wbn[tid] = wide.mode == AE_OP_BNRELU ? __ldg(wide.bnc + (tid >> 5 ? AE_BNC_SHIFT : AE_BNC_SCALE) * WC + lane)
                                                     : (tid >> 5 ? 0.f : 1.f);
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  float err = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int tiles = batch * TILES_PER_IMAGE;
  auto issue = [&](int tile, int st) {
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const uint32_t bar = bar0 + 8u * st;
    const int rows = tr == TILES_PER_IMAGE - 1 ? TILE_ROWS : SC_ROWS;      // the halo row below the image does not exist
    mbar_arrive_expect_tx(bar, (uint32_t)rows * WW * WC * 4);
    const size_t m0 = ((size_t)n * WH + tr * TILE_ROWS) * WW;
    bulk_copy_g2s(smem_u32(stage0 + st * SC_PIX * 32), wide.src + m0 * WC, (uint32_t)rows * WW * WC * 4, bar);
  };
  if (tid == 0 && (int)blockIdx.x < tiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int st = it & 1;
    const int n = tile / TILES_PER_IMAGE, tr = tile - n * TILES_PER_IMAGE;
    const float* raw = stage0 + st * SC_PIX * 32;
    if (tid == 0 && tile + (int)gridDim.x < tiles) {
      fence_proxy_async();
      issue(tile + gridDim.x, st ^ 1);
    }
    mbar_wait(bar0 + 8u * st, (it >> 1) & 1);
    const int valid_rows = tr == TILES_PER_IMAGE - 1 ? TILE_ROWS : SC_ROWS;
    // build the two column-shift copies: 160 pixels x 8 float4 units
#pragma unroll
    for (int j = 0; j < SC_PIX * 8 / TT_THREADS; ++j) {
      const int i = tid + j * TT_THREADS;
      const int p = i >> 3, q = i & 7, c = q * 4;
      const int pr = p >> 5, pc = p & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pr < valid_rows) {
        v = *reinterpret_cast<const float4*>(raw + i * 4);
        const float4 k0 = *reinterpret_cast<const float4*>(wbn + c), k1 = *reinterpret_cast<const float4*>(wbn + 32 + c);
        if (wide.mode == AE_OP_BNRELU) {
          v.x = fmaxf(fmaf(v.x, k0.x, k1.x), 0.f); v.y = fmaxf(fmaf(v.y, k0.y, k1.y), 0.f);
          v.z = fmaxf(fmaf(v.z, k0.z, k1.z), 0.f); v.w = fmaxf(fmaf(v.w, k0.w, k1.w), 0.f);
        }
      }
      uint2 h, l;
      split4(v.x, v.y, v.z, v.w, h, l);
      const uint32_t sub = (uint32_t)(q & 1) * 8u;
      *reinterpret_cast<uint2*>(Ac + sw64_off(p, q >> 1) + sub) = h;
      if (NSPLIT == 2) *reinterpret_cast<uint2*>(Ac + SC_COPY + sw64_off(p, q >> 1) + sub) = l;
      if (pc > 0) {                                         // the dx = 1 copy: row p - 1 sees this pixel
        *reinterpret_cast<uint2*>(Ac + NSPLIT * SC_COPY + sw64_off(p - 1, q >> 1) + sub) = h;
        if (NSPLIT == 2) *reinterpret_cast<uint2*>(Ac + NSPLIT * SC_COPY + SC_COPY + sw64_off(p - 1, q >> 1) + sub) = l;
      }
    }
    if (tid < SC_ROWS * 4 * NSPLIT) {                       // right edge of the dx = 1 copy: zeros
      const int pl = tid / (SC_ROWS * 4), r = (tid / 4) % SC_ROWS, ch = tid & 3;
      *reinterpret_cast<uint4*>(Ac + NSPLIT * SC_COPY + pl * SC_COPY + sw64_off(r * 32 + 31, ch)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(16, 0, 0);
      const uint32_t a_base = smem_u32(Ac), b_base = smem_u32(Bz);
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int dy = nb >> 1, dx = nb & 1;
        const uint32_t a0 = a_base + dx * NSPLIT * SC_COPY + dy * 32 * 64;    // rows shifted down by one wide row
        const uint32_t bb = b_base + nb * 1024;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t ah = make_desc(a0 + kk * 32, 16, 512, 64), bh = make_desc(bb + kk * 32, 16, 512, 64);
          umma_bf16(tmem_base, ah, bh, idesc, (nb | kk) != 0);
          if (NSPLIT == 2) {
            const uint64_t al = make_desc(a0 + SC_COPY + kk * 32, 16, 512, 64), bl = make_desc(bb + 4 * 1024 + kk * 32, 16, 512, 64);
            umma_bf16(tmem_base, ah, bl, idesc, 1);
            umma_bf16(tmem_base, al, bh, idesc, 1);
          }
        }
      }
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, it & 1);
    tc_fence_after();
    float acc[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
    const int ty0 = 2 * (tr * TILE_ROWS + warp);          // first of this thread's two thin rows (pixel (warp, lane))
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      const float b = co == 0 ? b0 : (co == 1 ? b1 : b2);
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const size_t o = (((size_t)n * 3 + co) * TH + ty0 + py) * TW + 2 * lane;
        const float s0 = 1.f / (1.f + expf(-(acc[(py * 2 + 0) * 3 + co] + b)));
        const float s1 = 1.f / (1.f + expf(-(acc[(py * 2 + 1) * 3 + co] + b)));
        *reinterpret_cast<float2*>(x_hat + o) = make_float2(s0, s1);
        if (x) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(x + o));
          err += (s0 - t.x) * (s0 - t.x) + (s1 - t.y) * (s1 - t.y);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (x && sse) {
    err = warp_sum(err);
    if (lane == 0) red[warp] = err;
    __syncthreads();
    if (tid == 0) atomicAdd(sse, ((double)red[0] + (double)red[1]) + ((double)red[2] + (double)red[3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

int thin_tc_scatter_sigmoid_fwd(const Operand& wide, const float* w, const float* bias, float* x_hat, const float* x,
                                double* sse, int batch, int nsplit, cudaStream_t st) {
  AE_CHECK(wide.mode == AE_OP_RAW || wide.mode == AE_OP_BNRELU, "thin_tc_scatter: wide operand mode %d not supported", wide.mode);
  AE_CHECK(((uintptr_t)wide.src & 15) == 0, "thin_tc_scatter: wide tensor must be 16-byte aligned");
  const size_t smem = 1024 + (size_t)2 * nsplit * SC_COPY + (size_t)nsplit * 4 * 1024 + 64 * 4 + 2 * (size_t)SC_PIX * 32 * 4;
  static bool attr_done[2] = {false, false};
  if (!attr_done[nsplit - 1]) {
    if (nsplit == 2) AE_CUDA(cudaFuncSetAttribute(k_thin_tc_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else AE_CUDA(cudaFuncSetAttribute(k_thin_tc_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[nsplit - 1] = true;
  }
  if (nsplit == 2) k_thin_tc_scatter<2><<<tc_blocks(batch), TT_THREADS, smem, st>>>(wide, w, bias, x_hat, x, sse, batch);
  else k_thin_tc_scatter<1><<<tc_blocks(batch), TT_THREADS, smem, st>>>(wide, w, bias, x_hat, x, sse, batch);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

// TMA-fed tcgen05 / TMEM implicit-GEMM kernels for the six 32..256-channel stride-2 layers (sm_100a only).
//
// Activation operands are "split-bf16" NHWC tensors: [nsplit][B][H][W][C] bf16, plane 0 = round-to-nearest bf16
// of the fp32 value, plane 1 (AE_PREC_FP32 only) = bf16 of the remainder.  They are produced once per tensor by
// k_split_operand (which also applies BatchNorm+ReLU / BatchNorm-backward), so the GEMM kernels contain no
// CUDA-core operand math: one thread issues 5-d TMA tile loads (cp.async.bulk.tensor) straight into the
// SWIZZLE_128B / SWIZZLE_64B shared-memory layout tcgen05.mma reads, one thread issues the MMAs, four warps run
// the epilogue out of tensor memory.
//
//   row GEMM   D[128 pixels x NT channels] = A[128 x K] * W[NT x K]^T
//       FAM_FPROP (big -> small): Conv2d forward / ConvTranspose2d data gradient.  K = 9 taps x Cb.  The stride-2
//                  gather is expressed as four tensor maps, one per (row, column) parity of the big image, each a
//                  plain unit-stride lattice; tap (ky,kx) is a box of one parity lattice shifted by -1 or 0 and
//                  the zero padding is TMA's out-of-bounds fill.
//       FAM_DGRAD (small -> big, 4 output-parity phases): ConvTranspose2d forward / Conv2d data gradient.
//                  K = (1,2,2,4) taps x Cs; every tap is a box of the small image shifted by 0 or +1.
//   weight gradient  dW[(tap,cb) x cs] = sum over small pixels: A = gathered big image, B = small image, both
//       MN-major operands (the reduction index is the pixel = the row of the TMA box), split over pixel slices,
//       partial tiles reduced in a fixed order (deterministic).
//
// AE_PREC_FP32 issues hi*hi + hi*lo + lo*hi (3 MMAs per k-step, fp32 accumulate): ~2^-17 relative error per product.
#include <cstdlib>
#include <mutex>

#include "pack.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace ae {

// ---------------------------------------------------------------------------------------------
// k_split_operand: fp32 NHWC (+ operand transform) -> split-bf16 planes
// ---------------------------------------------------------------------------------------------
// Optionally runs a BatchNorm coefficient job first (every block derives the coefficients of all channels into shared
// memory from the fp64 sums; block 0 also publishes them and updates the running statistics / parameter gradients),
// which removes the separate k_bn_finalize / k_bn_bwd_reduce launch in front of it.
template <int NSPLIT>
__global__ void __launch_bounds__(256) k_split_operand(Operand op, int64_t n8, int64_t plane_elems,
                                                      __nv_bfloat16* __restrict__ dst, BnJob job) {
  __shared__ __align__(16) float sc[4][256];             // BNRELU: scale, shift.  BNBWD: A, B, C, mean
  const int C = op.C;
  if (op.mode != AE_OP_RAW) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (job.kind == BN_JOB_FINALIZE) {
        double mean, var;
        if (job.training) {
          mean = job.stats[c] / job.count;
          var = job.stats[C + c] / job.count - mean * mean;
          if (var < 0.0) var = 0.0;
        } else {
          mean = (double)job.rmean[c];
          var = (double)job.rvar[c];
        }
        const float rstd = (float)(1.0 / sqrt(var + 1e-5));
        const float scale = job.gamma[c] * rstd;
        const float shift = job.beta[c] - (float)mean * scale;
        sc[0][c] = scale; sc[1][c] = shift;
        if (blockIdx.x == 0) {
          if (c == 0 && job.training && job.nbt) *job.nbt += 1;
          if (job.training && job.rmean) {
            const double unb = job.count > 1.0 ? var * job.count / (job.count - 1.0) : var;
            job.rmean[c] = (float)(0.9 * (double)job.rmean[c] + 0.1 * mean);
            job.rvar[c] = (float)(0.9 * (double)job.rvar[c] + 0.1 * unb);
          }
          job.bnc[AE_BNC_SCALE * C + c] = scale; job.bnc[AE_BNC_SHIFT * C + c] = shift;
          job.bnc[AE_BNC_MEAN * C + c] = (float)mean; job.bnc[AE_BNC_RSTD * C + c] = rstd;
        }
      } else if (job.kind == BN_JOB_BWD) {
        const double s1 = job.stats[c], s2 = job.stats[C + c];
        const double rstd = (double)job.bnc[AE_BNC_RSTD * C + c];
        const double a = (double)job.gamma[c] * rstd;
        const float fa = (float)a, fb = (float)(-a * rstd * s2 / job.count), fk = (float)(-a * s1 / job.count);
        sc[0][c] = fa; sc[1][c] = fb; sc[2][c] = fk; sc[3][c] = job.bnc[AE_BNC_MEAN * C + c];
        if (blockIdx.x == 0) {
          job.bnc[AE_BNC_A * C + c] = fa; job.bnc[AE_BNC_B * C + c] = fb; job.bnc[AE_BNC_C * C + c] = fk;
          if (job.dgamma) job.dgamma[c] = (float)s2;
          if (job.dbeta) job.dbeta[c] = (float)s1;
          if (job.dzero) job.dzero[c] = 0.f;
        }
      } else if (op.mode == AE_OP_BNRELU) {
        sc[0][c] = op.bnc[AE_BNC_SCALE * C + c]; sc[1][c] = op.bnc[AE_BNC_SHIFT * C + c];
      } else {
        sc[0][c] = op.bnc[AE_BNC_A * C + c]; sc[1][c] = op.bnc[AE_BNC_B * C + c];
        sc[2][c] = op.bnc[AE_BNC_C * C + c]; sc[3][c] = op.bnc[AE_BNC_MEAN * C + c];
      }
    }
    __syncthreads();
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const size_t off = (size_t)i * 8;
    const int c = (int)(off % (size_t)C);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(op.src + off));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(op.src + off) + 1);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (op.mode == AE_OP_BNRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[0][c + j], sc[1][c + j]), 0.f);
    } else if (op.mode == AE_OP_BNBWD) {
      // dy = A*dz + B*(y - mean) + C, (y - mean) formed first (same order as load_operand4)
      const float4 y0 = __ldg(reinterpret_cast<const float4*>(op.src2 + off));
      const float4 y1 = __ldg(reinterpret_cast<const float4*>(op.src2 + off) + 1);
      const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(sc[0][c + j], v[j], fmaf(sc[1][c + j], y[j] - sc[3][c + j], sc[2][c + j]));
    }
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + off) = hi;
    if (NSPLIT == 2) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
      uint4 lo;
      lo.x = pack_bf16x2(r[0], r[1]); lo.y = pack_bf16x2(r[2], r[3]); lo.z = pack_bf16x2(r[4], r[5]); lo.w = pack_bf16x2(r[6], r[7]);
      *reinterpret_cast<uint4*>(dst + plane_elems + off) = lo;
    }
  }
}

int tma_split_operand(const Operand& op, int64_t count, void* planes, int nsplit, const BnJob* job, cudaStream_t st) {
  AE_CHECK(count % 8 == 0 && op.C % 8 == 0, "split_operand: count=%lld and C=%d must be multiples of 8", (long long)count, op.C);
  AE_CHECK(op.mode == AE_OP_RAW || op.mode == AE_OP_BNRELU || op.mode == AE_OP_BNBWD, "split_operand: unsupported operand mode %d", op.mode);
  AE_CHECK(op.mode == AE_OP_RAW || op.C <= 256, "split_operand: at most 256 BatchNorm channels (got %d)", op.C);
  AE_CHECK(((uintptr_t)planes & 15) == 0, "split_operand: destination must be 16-byte aligned");
  BnJob j;
  memset(&j, 0, sizeof(j));
  if (job) {
    j = *job;
    AE_CHECK(j.C == op.C && j.bnc == op.bnc, "split_operand: the BatchNorm job must describe the operand's own layer");
    AE_CHECK((j.kind == BN_JOB_FINALIZE && op.mode == AE_OP_BNRELU) || (j.kind == BN_JOB_BWD && op.mode == AE_OP_BNBWD),
             "split_operand: BatchNorm job kind %d does not match operand mode %d", j.kind, op.mode);
  }
  const int64_t n8 = count / 8;
  int64_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (nsplit == 2) k_split_operand<2><<<(int)blocks, 256, 0, st>>>(op, n8, count, (__nv_bfloat16*)planes, j);
  else k_split_operand<1><<<(int)blocks, 256, 0, st>>>(op, n8, count, (__nv_bfloat16*)planes, j);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row GEMM kernel
// ---------------------------------------------------------------------------------------------
static constexpr int RG_THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..5: epilogue

// Developer build only (python build.py with AE_B200_BUILD_VARIANT=trace): per-CTA phase timestamps of the row GEMM, read
// back by scripts/trace_rowgemm.py.  The product library is compiled without AE_TRACE and contains none of this.
#ifdef AE_TRACE
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define AE_TR(slot) do { if (g_trace) g_trace[(size_t)blockIdx.x * 16 + (slot)] = gtimer(); } while (0)
#else
#define AE_TR(slot) do { } while (0)
#endif

struct alignas(64) TmaRow {
  CUtensorMap amap[4];      // FPROP: parity lattices (py*2+px) of the big image; DGRAD: [0] = small image
  const uint8_t* wtiles;    // packed weight tiles
  Geom g;
  Epilogue epi;
  float* out;
  int M, N;
  int cpt;                  // K chunks per tap (C / KC)
  int wchunks;              // K chunks per n-tile in the weight pack (all taps / all phases)
  int bx, by, bn;           // pixel box of one 128-row tile
  BnJob job;                // kind != BN_JOB_NONE: coefficient job of epi.stats, run by the last CTA to finish
  unsigned int* job_counter;
};

// Persistent: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  A tile is (phase, m-tile, n-tile) with
// the n-tile fastest.  The accumulator is double-buffered in tensor memory, so the epilogue of tile i overlaps the
// TMA loads and MMAs of tile i+1, and barrier / TMEM set-up is paid once per CTA.
template <int FAMILY, int NT, int KC, int NSPLIT, int STAGES>
__global__ void __launch_bounds__(RG_THREADS) k_tma_rowgemm(const __grid_constant__ TmaRow q) {
  constexpr int ROWB = KC * 2;                                // bytes per shared-memory row
  constexpr uint32_t A_PLANE = TILE_M * ROWB;
  constexpr uint32_t B_PLANE = NT * ROWB;
  constexpr uint32_t A_BYTES = NSPLIT * A_PLANE, B_BYTES = NSPLIT * B_PLANE;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t TMEM_COLS = 2 * NT;                      // two accumulators
  // The weight pack is laid out in n-tiles of PK_NT = min(NT, 64) rows, [hi plane][lo plane] per (n-tile, chunk).  A
  // 128-wide tile takes two of them: their planes are copied side by side so that each plane is 128 contiguous rows.
  constexpr int PK_NT = NT < 64 ? NT : 64;
  constexpr int NSUB = NT / PK_NT;
  constexpr uint32_t PK_PLANE = PK_NT * ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float sStat[2][256];                           // per output channel, accumulated over every tile of this CTA
  __shared__ __align__(16) float sCoef[256][4];             // per output channel: {scale, shift, mean, rstd} or {bias, 0, 0, 0}

  const Geom g = q.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) AE_TR(0);
  const int tiles_m = (q.M + TILE_M - 1) / TILE_M, tiles_n = q.N / NT;
  const int tiles_mn = tiles_m * tiles_n;
  const int num_tiles = tiles_mn * (FAMILY == FAM_DGRAD ? 4 : 1);

  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };

  if (tid == 0) {
    // the descriptors live in the kernel parameters: start fetching them while the CTA sets itself up
    if (FAMILY == FAM_FPROP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tma_prefetch_desc(&q.amap[i]);
    } else {
      tma_prefetch_desc(&q.amap[0]);
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) AE_TR(1);

  // tile -> (phase, ntile, m0, number of K chunks, first chunk inside the weight pack)
  auto decode = [&](int tile, int& phase, int& ntile, int& m0, int& nkc, int& kc_off) {
    // n-tile fastest: the CTAs working on one m-tile at the same time share its A boxes through the L2 (with the m-tile
    // fastest, an operand larger than the L2 was re-read from DRAM once per n-tile)
    phase = tile / tiles_mn;
    const int rem = tile - phase * tiles_mn;
    const int mt = rem / tiles_n;
    ntile = rem - mt * tiles_n;
    m0 = mt * TILE_M;
    if (FAMILY == FAM_DGRAD) {
      const int py = phase >> 1, px = phase & 1;
      nkc = (1 + py) * (1 + px) * q.cpt;
      kc_off = (phase == 0 ? 0 : phase == 1 ? 1 : phase == 2 ? 3 : 5) * q.cpt;
    } else {
      nkc = 9 * q.cpt;
      kc_off = 0;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      const int P = g.Hs * g.Ws;
      uint32_t gc = 0;                                        // chunk counter over all tiles of this CTA
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int phase, ntile, m0, nkc, kc_off;
        decode(tile, phase, ntile, m0, nkc, kc_off);
        const int py = phase >> 1, px = phase & 1;
        const int n0 = m0 >> (g.lHs + g.lWs);
        const int y0 = (m0 & (P - 1)) >> g.lWs;
        const uint8_t* wsrc = q.wtiles + ((size_t)ntile * NSUB * q.wchunks + kc_off) * (NSPLIT * PK_PLANE);
        for (int it = 0; it < nkc; ++it, ++gc) {
          const uint32_t s = gc % STAGES, round = gc / STAGES;
          if (round > 0) mbar_wait(empty_bar(s), (round - 1) & 1);
          const uint32_t a_dst = smem_u32(smem + (size_t)s * STAGE_BYTES);
          mbar_arrive_expect_tx(full_bar(s), STAGE_BYTES);
          const int tap = it / q.cpt, c0 = (it - tap * q.cpt) * KC;
          if (FAMILY == FAM_FPROP) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const int pmap = ((ky + 1) & 1) * 2 + ((kx + 1) & 1);     // parity lattice of source row 2*oy-1+ky
            tma_load_5d(a_dst, &q.amap[pmap], c0, (kx == 0) ? -1 : 0, y0 + ((ky == 0) ? -1 : 0), n0, 0, full_bar(s));
          } else {
            const int a = tap / (1 + px), b = tap - a * (1 + px);
            const int dy = (py && a == 0) ? 1 : 0, dx = (px && b == 0) ? 1 : 0;   // source = (y + dy, x + dx)
            tma_load_5d(a_dst, &q.amap[0], c0, dx, y0 + dy, n0, 0, full_bar(s));
          }
          if (FAMILY == FAM_DGRAD) {
            // group pack (pack.cuh): the planes of a tile are not adjacent -- one copy per plane and 64-row sub-tile
            const int slot = kc_off / q.cpt + tap, c = it - tap * q.cpt;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const uint8_t* blk = q.wtiles + ((size_t)(ntile * NSUB + sub) * q.cpt + c) * 9 * (NSPLIT * PK_PLANE);
#pragma unroll
              for (int pl = 0; pl < NSPLIT; ++pl)
                bulk_copy_g2s(a_dst + A_BYTES + pl * B_PLANE + sub * PK_PLANE, blk + dg_pack_plane_offset(slot, pl, NSPLIT, PK_PLANE),
                              PK_PLANE, full_bar(s));
            }
          } else if (NSUB == 1) {
            bulk_copy_g2s(a_dst + A_BYTES, wsrc + (size_t)it * B_BYTES, B_BYTES, full_bar(s));
          } else {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
              for (int pl = 0; pl < NSPLIT; ++pl)
                bulk_copy_g2s(a_dst + A_BYTES + pl * B_PLANE + sub * PK_PLANE,
                              wsrc + ((size_t)sub * q.wchunks + it) * (NSPLIT * PK_PLANE) + pl * PK_PLANE, PK_PLANE, full_bar(s));
          }
          if (gc == 0) AE_TR(2);
        }
      }
      AE_TR(3);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(NT, 0, 0);
      uint32_t gc = 0, ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        int phase, ntile, m0, nkc, kc_off;
        decode(tile, phase, ntile, m0, nkc, kc_off);
        const uint32_t ab = ti & 1, ar = ti >> 1;
        if (ar > 0) { mbar_wait(tempty_bar(ab), (ar - 1) & 1); tc_fence_after(); }
        const uint32_t acc = tmem_base + ab * NT;
        for (int it = 0; it < nkc; ++it, ++gc) {
          const uint32_t s = gc % STAGES, round = gc / STAGES;
          mbar_wait(full_bar(s), round & 1);
          tc_fence_after();
          if (gc == 0) AE_TR(4);
          const uint32_t a0 = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t b0 = a0 + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint64_t ah = make_desc(a0 + kk * 32, 16, SBO, ROWB), bh = make_desc(b0 + kk * 32, 16, SBO, ROWB);
            umma_bf16(acc, ah, bh, idesc, (it | kk) != 0);
            if (NSPLIT == 2) {
              const uint64_t al = make_desc(a0 + A_PLANE + kk * 32, 16, SBO, ROWB);
              const uint64_t bl = make_desc(b0 + B_PLANE + kk * 32, 16, SBO, ROWB);
              umma_bf16(acc, ah, bl, idesc, 1);
              umma_bf16(acc, al, bh, idesc, 1);
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(ab));
      }
      AE_TR(5);
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5; warp w owns TMEM lanes 32*(w%4) ..) =====================
    // epilogue-only state (statistics accumulators, per-channel coefficients): loaded behind the CTA-wide barrier so that the
    // producer's first copies do not wait for these global loads; bar 1 = the four epilogue warps
    for (int c = tid - 64; c < 256; c += 128) { sStat[0][c] = 0.f; sStat[1][c] = 0.f; }
    for (int c = tid - 64; c < q.N; c += 128) {              // q.N <= 256 output channels
      float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q.epi.mode == AE_EPI_RELUBWD_STATS) {
        const int ch = c % q.epi.C;
        k = make_float4(__ldg(q.epi.bnc + AE_BNC_SCALE * q.epi.C + ch), __ldg(q.epi.bnc + AE_BNC_SHIFT * q.epi.C + ch),
                        __ldg(q.epi.bnc + AE_BNC_MEAN * q.epi.C + ch), __ldg(q.epi.bnc + AE_BNC_RSTD * q.epi.C + ch));
      } else if (q.epi.mode == AE_EPI_BNRELU_SPLIT) {           // relu(scale*(acc + bias) + shift), rounded exactly as the two-pass form
        const int ch = c % q.epi.C;
        k = make_float4(__ldg(q.epi.bnc + AE_BNC_SCALE * q.epi.C + ch), __ldg(q.epi.bnc + AE_BNC_SHIFT * q.epi.C + ch),
                        q.epi.bias ? __ldg(q.epi.bias + c) : 0.f, 0.f);
      } else if (q.epi.bias) {
        k.x = __ldg(q.epi.bias + c);
      }
      *reinterpret_cast<float4*>(&sCoef[c][0]) = k;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int et = tid - 64;
    const Epilogue e = q.epi;
    const bool do_stats = e.mode != AE_EPI_STORE && e.stats != nullptr;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      int phase, ntile, m0, nkc, kc_off;
      decode(tile, phase, ntile, m0, nkc, kc_off);
      const int py = phase >> 1, px = phase & 1;
      const uint32_t ab = ti & 1;
      const int m = m0 + row;
      const bool row_ok = m < q.M;
      size_t orow = 0;
      if (row_ok) {
        if (FAMILY == FAM_DGRAD) {
          const int x = m & (g.Ws - 1), y = (m >> g.lWs) & (g.Hs - 1), nn = m >> (g.lWs + g.lHs);
          orow = (((size_t)nn * (2 * g.Hs) + 2 * y + py) * (2 * g.Ws) + 2 * x + px) * q.N;
        } else {
          orow = (size_t)m * q.N;
        }
      }
      mbar_wait(tfull_bar(ab), (ti >> 1) & 1);
      tc_fence_after();
      if (et == 0) { if (ti == 0) AE_TR(6); AE_TR(7); }
#pragma unroll 1
      for (int col0 = 0; col0 < NT; col0 += 32) {
        const int n = ntile * NT + col0;                 // global output channel
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + ab * NT + (uint32_t)col0, v);
        if (et == 0 && col0 == 0) AE_TR(11);
        if (col0 + 32 >= NT) {                           // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab));
        }
        float s2[32];
        if (e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok) yv = __ldg(reinterpret_cast<const float4*>(e.y + orow + n) + j4);
            const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 k = *reinterpret_cast<const float4*>(&sCoef[n + j4 * 4 + j][0]);
              const float z = fmaf(ya[j], k.x, k.y);
              const float d = (row_ok && z > 0.f) ? v[j4 * 4 + j] : 0.f;
              v[j4 * 4 + j] = d;
              s2[j4 * 4 + j] = d * ((ya[j] - k.z) * k.w);
            }
          }
        } else if (e.mode == AE_EPI_BNRELU_SPLIT) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 k = *reinterpret_cast<const float4*>(&sCoef[n + j][0]);
            v[j] = fmaxf(fmaf(v[j] + k.z, k.x, k.y), 0.f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float d = v[j] + sCoef[n + j][0];
            d = row_ok ? d : 0.f;
            v[j] = d;
            s2[j] = d * d;
          }
        }
        if (row_ok) {
          if (e.mode == AE_EPI_BNRELU_SPLIT) {
            const size_t plane = (size_t)q.M * q.N * (FAMILY == FAM_DGRAD ? 4 : 1);
            st_global_split32<NSPLIT>(reinterpret_cast<__nv_bfloat16*>(q.out) + orow + n, plane, v);
          } else {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) st_global_v8(q.out + orow + n + j8 * 8, v, j8);
          }
        }
        if (et == 0 && col0 == 0) AE_TR(12);
        if (do_stats) {
          const float a = warp_colsum32_tc(v, lane);
          const float b = warp_colsum32_tc(s2, lane);
          atomicAdd(&sStat[0][n + lane], a);
          atomicAdd(&sStat[1][n + lane], b);
        }
        if (et == 0 && col0 == 0) AE_TR(13);
      }
    }
    if (et == 0) AE_TR(8);
    if (do_stats) {                                         // one flush per CTA: fp64 atomics, one per channel
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = et; c < q.N; c += 128) {
        const float a = sStat[0][c], b = sStat[1][c];
        if (a != 0.f || b != 0.f) {
          const int ch = c % e.C;
          atomicAdd(e.stats + ch, (double)a);
          atomicAdd(e.stats + e.C + ch, (double)b);
        }
      }
    }
    if (q.job.kind != BN_JOB_NONE) {
      // last CTA out: every CTA's statistics are in (its fp64 atomics precede its count), derive the BatchNorm coefficients
      __shared__ unsigned int s_last;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const unsigned int done = atomicAdd(q.job_counter, 1u);
        s_last = done == gridDim.x - 1 ? 1u : 0u;
        if (s_last) *q.job_counter = 0u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (s_last) {
        __threadfence();
        bn_job_run(q.job, et, 128);
      }
    }
    if (et == 0) AE_TR(9);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tid == 0) AE_TR(10);
}

#ifdef AE_TRACE
extern "C" int ae_debug_set_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

// ---------------------------------------------------------------------------------------------
// weight packing: w [Cs][Cb][3][3] fp32 -> swizzled bf16 (hi[, lo]) tiles for both row-GEMM orientations.
// Tile (n_tile, chunk): [plane][NT rows][KC*2 bytes]; element (r, j) of a tile at
//   r*ROWB + ((chunk16 ^ swz(r)) << 4) + (j & 7)*2,  chunk16 = j >> 3,
//   swz(r) = r & 7 for 128-byte rows (SWIZZLE_128B), (r >> 1) & 3 for 64-byte rows (SWIZZLE_64B).
// fwd  : n = cs, K order (tap, cb), KC = min(Cb, 64).   dgrad: n = cb, KC = 64, per (n-tile, K chunk) a block of the 9 tap
//        tiles in the group order of pack.cuh (the tiles one N-stacked MMA reads are contiguous, plane by plane).
// ---------------------------------------------------------------------------------------------
static inline int nt_for(int n) { return n >= 64 ? 64 : 32; }
static inline int kc_fwd(int Cb) { return Cb >= 64 ? 64 : 32; }

__global__ void k_tma_pack_conv(const float* __restrict__ w, int Cs, int Cb, int nsplit, uint8_t* __restrict__ fwd,
                                uint8_t* __restrict__ dgrad) {
  const int total = 2 * 9 * Cs * Cb / 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    pack_conv_chunk(idx, w, Cs, Cb, nsplit, fwd, dgrad);
}

// Every re-layout job of one optimizer step in a single launch.
__global__ void __launch_bounds__(256) k_pack_all(const __grid_constant__ PackJobs jobs) {
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.job[j + 1].first_block) ++j;
  const PackJob& J = jobs.job[j];
  const int nblk = (j + 1 < jobs.n ? jobs.job[j + 1].first_block : (int)gridDim.x) - J.first_block;
  for (int idx = ((int)blockIdx.x - J.first_block) * blockDim.x + threadIdx.x; idx < J.total; idx += nblk * blockDim.x) {
    if (J.kind == PACK_CONV) pack_conv_chunk(idx, J.src, J.a, J.b, J.c, (uint8_t*)J.dst, (uint8_t*)J.dst2);
    else if (J.kind == PACK_LINEAR) pack_linear_elem(idx, J.src, J.a, J.b, J.c, J.d, J.e, (float*)J.dst);
    else if (J.kind == PACK_DENSE_TC) pack_dense_tc_chunk(idx, J.src, J.a, J.b, J.c, J.d, J.e, (uint8_t*)J.dst);
    else permute_elem(idx, J.src, J.b, J.c, (float*)J.dst);
  }
}

int pack_all(PackJobs& jobs, cudaStream_t st) {
  AE_CHECK(jobs.n >= 1 && jobs.n <= PACK_MAX_JOBS, "pack_all: %d jobs", jobs.n);
  int blocks = 0;
  for (int j = 0; j < jobs.n; ++j) {
    jobs.job[j].first_block = blocks;
    int b = (jobs.job[j].total + 255) / 256;
    if (b > 148) b = 148;
    if (b < 1) b = 1;
    blocks += b;
  }
  k_pack_all<<<blocks, 256, 0, st>>>(jobs);
  AE_LAUNCH_CHECK();
  return 0;
}

size_t tma_packed_bytes(int Cs, int Cb, int nsplit) { return (size_t)9 * Cs * Cb * 2 * nsplit + 1024; }

int tma_pack_conv(const float* w, int Cs, int Cb, int nsplit, void* fwd, void* dgrad, cudaStream_t st) {
  AE_CHECK(Cs % 64 == 0 && Cb % 32 == 0, "tma_pack_conv: Cs=%d must be a multiple of 64 and Cb=%d of 32", Cs, Cb);
  AE_CHECK((((uintptr_t)fwd | (uintptr_t)dgrad) & 15) == 0, "tma_pack_conv: packed buffers must be 16-byte aligned");
  const int total = 2 * 9 * Cs * Cb / 8;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_tma_pack_conv<<<blocks, 256, 0, st>>>(w, Cs, Cb, nsplit, (uint8_t*)fwd, (uint8_t*)dgrad);
  AE_LAUNCH_CHECK();
  return 0;
}

bool tma_rowgemm_supported(const RowGemm& p) {
  if (!is_pow2(p.g.Hs) || !is_pow2(p.g.Ws) || p.g.Ws > 128 || p.g.Hs * p.g.Ws < 8) return false;
  if (p.family == FAM_FPROP) return (p.g.Cb == 32 || p.g.Cb % 64 == 0) && p.N % 32 == 0 && p.splitK <= 1;
  if (p.family == FAM_DGRAD) return p.g.Cs % 64 == 0 && p.N % 32 == 0;
  return false;
}

template <int FAMILY, int NT, int KC, int NSPLIT, int STAGES>
static int launch_row(const TmaRow& q, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * NSPLIT * (TILE_M * KC * 2 + NT * KC * 2) + 1024;
  const int tiles = (int)(grid.x * grid.y * grid.z);
  static bool attr_done = false;
  if (!attr_done) {
    AE_CUDA(cudaFuncSetAttribute(k_tma_rowgemm<FAMILY, NT, KC, NSPLIT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int per_sm = 2 * (smem + 2048) <= 227 * 1024 ? 2 : 1;   // ~97 KB per CTA: two CTAs per SM; the 128-wide tiles: one
  const int ctas = tiles < per_sm * 148 ? tiles : per_sm * 148;
  k_tma_rowgemm<FAMILY, NT, KC, NSPLIT, STAGES><<<ctas, RG_THREADS, smem, st>>>(q);
  AE_LAUNCH_CHECK();
  return 0;
}

// p.A.src must point to the split-bf16 planes of the A image (AE_OP_SPLIT_BF16)
int tma_rowgemm(const RowGemm& p, const void* packed, int nsplit, cudaStream_t st) {
  AE_CHECK(tma_rowgemm_supported(p), "tma_rowgemm: unsupported shape");
  AE_CHECK(p.A.mode == AE_OP_SPLIT_BF16, "tma_rowgemm: the A operand must be split-bf16 planes (ae_split_operand)");
  AE_CHECK(((uintptr_t)packed & 15) == 0, "tma_rowgemm: packed weights must be 16-byte aligned");
  // training epilogues go to the second-generation kernel where it is the faster one.  AE_B200_ROWGEMM=1 / 2: A/B switch
  // for measurements and tests (1 = first generation always, 2 = second generation wherever it supports the problem)
  const char* gen_env = getenv("AE_B200_ROWGEMM");
  const int gen = gen_env ? atoi(gen_env) : 0;
  if (gen != 1 && (gen == 2 ? rowgemm2_supported(p) : rowgemm2_preferred(p))) return tma_rowgemm2(p, packed, nsplit, st);
  TmaRow q;
  memset(&q, 0, sizeof(q));
  q.wtiles = (const uint8_t*)packed;
  q.g = p.g; q.epi = p.epi; q.out = p.out; q.M = p.M; q.N = p.N;
  if (p.tail_job) {
    AE_CHECK(p.tail_counter != nullptr && p.tail_job->stats == p.epi.stats && p.tail_job->C <= 256,
             "tma_rowgemm: the tail job must describe the statistics this launch accumulates");
    q.job = *p.tail_job; q.job_counter = p.tail_counter;
  }
  const Geom& g = p.g;
  pixel_box(g.Hs, g.Ws, TILE_M, &q.bx, &q.by, &q.bn);
  // 64-wide n-tiles keep the most CTAs busy at training batch sizes; with thousands of m-tiles (inference) a 128-wide tile
  // halves the A bytes per output and the shared-memory reads per MMA (a 128x64x16 MMA needs 192 B/clk of operands)
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const char* force_env = getenv("AE_B200_FORCE_NT128");       // tests: 1 = wherever the shape allows, -1 = never
  const int force_nt128 = force_env ? atoi(force_env) : 0;
  // (bf16 mode issues a third of the MMAs and is bound by the operand loads: two narrower CTAs per SM hide them better,
  //  measured 6.41 vs 6.32 M images/s)
  const bool wide = p.N % 128 == 0 && (p.family == FAM_DGRAD || kc_fwd(g.Cb) == 64) &&
                    (force_nt128 > 0 || (force_nt128 == 0 && nsplit == 2 &&
                                         (int64_t)m_tiles * (p.N / 128) * (p.family == FAM_DGRAD ? 4 : 1) >= 4 * 148));
  const int NT = wide ? 128 : nt_for(p.N);
  dim3 grid(m_tiles, p.N / NT, 1);
  if (p.family == FAM_FPROP) {
    const int KC = kc_fwd(g.Cb);
    q.cpt = g.Cb / KC; q.wchunks = 9 * q.cpt;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        AE_TRY(encode_map(&q.amap[py * 2 + px], p.A.src, g.B, 2 * g.Hs, 2 * g.Ws, g.Cb, nsplit, py, px, 2, 2, g.Hs, g.Ws, KC,
                          q.bx, q.by, q.bn));
    if (KC == 64) {
      if (NT == 128) return nsplit == 2 ? launch_row<FAM_FPROP, 128, 64, 2, 3>(q, grid, st) : launch_row<FAM_FPROP, 128, 64, 1, 6>(q, grid, st);
      if (NT == 64) return nsplit == 2 ? launch_row<FAM_FPROP, 64, 64, 2, 2>(q, grid, st) : launch_row<FAM_FPROP, 64, 64, 1, 4>(q, grid, st);
      return nsplit == 2 ? launch_row<FAM_FPROP, 32, 64, 2, 2>(q, grid, st) : launch_row<FAM_FPROP, 32, 64, 1, 4>(q, grid, st);
    }
    if (NT == 64) return nsplit == 2 ? launch_row<FAM_FPROP, 64, 32, 2, 4>(q, grid, st) : launch_row<FAM_FPROP, 64, 32, 1, 6>(q, grid, st);
    return nsplit == 2 ? launch_row<FAM_FPROP, 32, 32, 2, 4>(q, grid, st) : launch_row<FAM_FPROP, 32, 32, 1, 6>(q, grid, st);
  }
  q.cpt = g.Cs / 64; q.wchunks = 9 * q.cpt;
  AE_TRY(encode_map(&q.amap[0], p.A.src, g.B, g.Hs, g.Ws, g.Cs, nsplit, 0, 0, 1, 1, g.Hs, g.Ws, 64, q.bx, q.by, q.bn));
  grid.z = 4;
  if (NT == 128) return nsplit == 2 ? launch_row<FAM_DGRAD, 128, 64, 2, 3>(q, grid, st) : launch_row<FAM_DGRAD, 128, 64, 1, 6>(q, grid, st);
  if (NT == 64) return nsplit == 2 ? launch_row<FAM_DGRAD, 64, 64, 2, 2>(q, grid, st) : launch_row<FAM_DGRAD, 64, 64, 1, 4>(q, grid, st);
  return nsplit == 2 ? launch_row<FAM_DGRAD, 32, 64, 2, 2>(q, grid, st) : launch_row<FAM_DGRAD, 32, 64, 1, 4>(q, grid, st);
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient kernel: D[(tap,cb) tile of 128][cs] += sum over a slice of small pixels
// ---------------------------------------------------------------------------------------------
struct alignas(64) TmaWgrad {
  CUtensorMap bigmap[4];    // parity lattices of the big image, box = 64 pixels x KCA channels
  CUtensorMap smallmap;     // small image, box = 64 pixels x 64 channels
  Geom g;
  float* partial;           // [slices][9*Cb][Cs]
  int chunks;               // 64-pixel chunks in total
  int chunks_per_slice;
  int I;                    // 9*Cb
};

// KCA = channels per A group (64, or 32 when Cb == 32); NB = Cs (64/128/256)
template <int KCA, int NB, int NSPLIT, int STAGES>
__global__ void __launch_bounds__(RG_THREADS) k_tma_wgrad(const __grid_constant__ TmaWgrad q) {
  constexpr int KPIX = 64;                                   // pixels (reduction rows) per stage
  constexpr int ROWA = KCA * 2;                              // bytes per A row
  constexpr int GA = 128 / KCA;                              // A groups per 128-row tile
  constexpr int GB = NB / 64;
  constexpr uint32_t A_GROUP = KPIX * ROWA;                  // one TMA box of one plane
  constexpr uint32_t A_PLANE = GA * A_GROUP;                 // = 16 KB
  constexpr uint32_t B_GROUP = KPIX * 128;
  constexpr uint32_t B_PLANE = GB * B_GROUP;
  constexpr uint32_t A_BYTES = NSPLIT * A_PLANE, B_BYTES = NSPLIT * B_PLANE;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;

  const Geom g = q.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, slice = blockIdx.y;
  const int ch_beg = slice * q.chunks_per_slice;
  const int ch_end = min(q.chunks, ch_beg + q.chunks_per_slice);
  const int nch = max(0, ch_end - ch_beg);
  // groups of this tile: linear group index gi -> (tap, c0)
  const int gpt = g.Cb / KCA;                                // groups per tap
  const int gi0 = mt * GA;
  int ngroups = 9 * gpt - gi0;
  if (ngroups > GA) ngroups = GA;

  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), NB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int P = g.Hs * g.Ws;
      for (int it = 0; it < nch; ++it) {
        const int s = it % STAGES, round = it / STAGES;
        if (round > 0) mbar_wait(empty_bar(s), (round - 1) & 1);
        const uint32_t a_dst = smem_u32(smem + (size_t)s * STAGE_BYTES);
        mbar_arrive_expect_tx(full_bar(s), (uint32_t)ngroups * NSPLIT * A_GROUP + B_BYTES);
        const int p0 = (ch_beg + it) * KPIX;                 // first small pixel of the chunk
        const int n0 = p0 >> (g.lHs + g.lWs);
        const int y0 = (p0 & (P - 1)) >> g.lWs;
        const int x0 = p0 & (g.Ws - 1);                      // non-zero only when Ws > 64 (not used by the model)
        for (int gq = 0; gq < ngroups; ++gq) {
          const int gi = gi0 + gq;
          const int tap = gi / gpt, c0 = (gi - tap * gpt) * KCA;
          const int ky = tap / 3, kx = tap - ky * 3;
          const int pmap = ((ky + 1) & 1) * 2 + ((kx + 1) & 1);
          // one box = [plane][64 pixels][ROWA] = NSPLIT * A_GROUP bytes: group gq lands at a_dst + gq * NSPLIT * A_GROUP
          tma_load_5d(a_dst + gq * NSPLIT * A_GROUP, &q.bigmap[pmap], c0, x0 + ((kx == 0) ? -1 : 0),
                      y0 + ((ky == 0) ? -1 : 0), n0, 0, full_bar(s));
        }
        for (int gq = 0; gq < GB; ++gq)
          tma_load_5d(a_dst + A_BYTES + gq * NSPLIT * B_GROUP, &q.smallmap, gq * 64, x0, y0, n0, 0, full_bar(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(NB, 1, 1);
      for (int it = 0; it < nch; ++it) {
        const int s = it % STAGES, round = it / STAGES;
        mbar_wait(full_bar(s), round & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t b0 = a0 + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < KPIX / 16; ++kk) {
          // 16 pixels = two 8-row groups: advance by 16 rows.  Shared layout per operand: [group][plane][64 rows][row bytes],
          // so consecutive M/N groups of one plane are NSPLIT * GROUP bytes apart (the descriptor's leading byte offset).
          const uint64_t ah = make_desc(a0 + kk * 16 * ROWA, NSPLIT * A_GROUP, 8 * ROWA, ROWA);
          const uint64_t bh = make_desc(b0 + kk * 16 * 128, NSPLIT * B_GROUP, 1024, 128);
          umma_bf16(tmem_base, ah, bh, idesc, (it | kk) != 0);
          if (NSPLIT == 2) {
            const uint64_t al = make_desc(a0 + A_GROUP + kk * 16 * ROWA, NSPLIT * A_GROUP, 8 * ROWA, ROWA);
            const uint64_t bl = make_desc(b0 + B_GROUP + kk * 16 * 128, NSPLIT * B_GROUP, 1024, 128);
            umma_bf16(tmem_base, ah, bl, idesc, 1);
            umma_bf16(tmem_base, al, bh, idesc, 1);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int i = mt * 128 + row;                            // (tap, cb) index
    float* dst = q.partial + ((size_t)slice * q.I + i) * NB;
    if (nch > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int col0 = 0; col0 < NB; col0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)col0, v);
        if (i < q.I) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) st_global_v8(dst + col0 + j8 * 8, v, j8);
        }
      }
    } else if (i < q.I) {
      for (int c = 0; c < NB; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NB);
  }
}

// dw[cs][cb][tap] = sum_s partial[s][tap*Cb + cb][cs]   (fixed order -> deterministic).  A block owns 32 float4
// outputs; its 8 warps each sum every 8th slice (all loads of a thread in flight together), then warp 0 adds the 8
// shares in order.
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ partial, int slices, int I, int J, int Cb,
                                                      float* __restrict__ dw) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int n4 = I * J / 4;
  const int idx = blockIdx.x * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < n4) {
    const size_t stride4 = (size_t)I * J / 4;
    int k = wq;
    for (; k + 24 < slices; k += 32) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)k * stride4 + idx);
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)(k + 8) * stride4 + idx);
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)(k + 16) * stride4 + idx);
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)(k + 24) * stride4 + idx);
      s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; k < slices; k += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)k * stride4 + idx);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[wq][lane] = s;
  __syncthreads();
  if (wq == 0 && idx < n4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) { const float4 v = red[g][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    const int e = idx * 4;
    const int i = e / J, j = e - i * J;
    const int tap = i / Cb, cb = i - tap * Cb;
    const size_t o = ((size_t)j * Cb + cb) * 9 + tap;
    const size_t js = (size_t)Cb * 9;
    dw[o] = t.x; dw[o + js] = t.y; dw[o + 2 * js] = t.z; dw[o + 3 * js] = t.w;
  }
}

bool tma_wgrad_supported(const Geom& g) {
  if (!is_pow2(g.Hs) || !is_pow2(g.Ws) || g.Ws > 64 || g.Hs * g.Ws < 8) return false;
  return (g.Cb == 32 || g.Cb == 64 || g.Cb == 128) && (g.Cs == 64 || g.Cs == 128 || g.Cs == 256);
}

int tma_wgrad_slices(const Geom& g) {
  const int chunks = (g.B * g.Hs * g.Ws + 63) / 64;
  const int mtiles = (9 * g.Cb + 127) / 128;
  int s = (148 + mtiles - 1) / mtiles;
  if (s > chunks) s = chunks;
  if (s < 1) s = 1;
  // every slice non-empty
  const int per = (chunks + s - 1) / s;
  return (chunks + per - 1) / per;
}

size_t tma_wgrad_partial_bytes(const Geom& g) { return (size_t)tma_wgrad_slices(g) * 9 * g.Cb * g.Cs * sizeof(float); }

template <int KCA, int NB, int NSPLIT, int STAGES>
static int launch_wgrad(const TmaWgrad& q, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * NSPLIT * (128 * 64 * 2 + NB * 64 * 2) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    AE_CUDA(cudaFuncSetAttribute(k_tma_wgrad<KCA, NB, NSPLIT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_tma_wgrad<KCA, NB, NSPLIT, STAGES><<<grid, RG_THREADS, smem, st>>>(q);
  AE_LAUNCH_CHECK();
  return 0;
}

// big / small: split-bf16 planes of the (transformed) big and small images; dw: [Cs][Cb][3][3] fp32
int tma_wgrad(const Geom& g, const void* big, const void* small, float* dw, float* partial, size_t partial_bytes, int nsplit,
              cudaStream_t st) {
  AE_CHECK(tma_wgrad_supported(g), "tma_wgrad: unsupported shape (Cb=%d Cs=%d Hs=%d Ws=%d)", g.Cb, g.Cs, g.Hs, g.Ws);
  const int slices = tma_wgrad_slices(g);
  AE_CHECK(partial != nullptr && partial_bytes >= (size_t)slices * 9 * g.Cb * g.Cs * 4, "tma_wgrad: partial buffer too small");
  TmaWgrad q;
  memset(&q, 0, sizeof(q));
  q.g = g; q.partial = partial; q.I = 9 * g.Cb;
  q.chunks = (g.B * g.Hs * g.Ws + 63) / 64;
  q.chunks_per_slice = (q.chunks + slices - 1) / slices;
  int bx, by, bn;
  pixel_box(g.Hs, g.Ws, 64, &bx, &by, &bn);
  const int KCA = g.Cb >= 64 ? 64 : 32;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      AE_TRY(encode_map(&q.bigmap[py * 2 + px], big, g.B, 2 * g.Hs, 2 * g.Ws, g.Cb, nsplit, py, px, 2, 2, g.Hs, g.Ws, KCA, bx, by, bn));
    }
  AE_TRY(encode_map(&q.smallmap, small, g.B, g.Hs, g.Ws, g.Cs, nsplit, 0, 0, 1, 1, g.Hs, g.Ws, 64, bx, by, bn));
  dim3 grid((q.I + 127) / 128, slices, 1);
  int rc;
#define AE_WG(KCA_, NB_)                                                                                            \
  (nsplit == 2 ? launch_wgrad<KCA_, NB_, 2, (NB_ == 256 ? 2 : NB_ == 128 ? 3 : 4)>(q, grid, st)                      \
               : launch_wgrad<KCA_, NB_, 1, (NB_ == 256 ? 4 : NB_ == 128 ? 5 : 6)>(q, grid, st))
  if (KCA == 64) rc = g.Cs == 64 ? AE_WG(64, 64) : g.Cs == 128 ? AE_WG(64, 128) : AE_WG(64, 256);
  else rc = g.Cs == 64 ? AE_WG(32, 64) : g.Cs == 128 ? AE_WG(32, 128) : AE_WG(32, 256);
#undef AE_WG
  AE_TRY(rc);
  const int n4 = q.I * g.Cs / 4;
  k_wgrad_reduce<<<(n4 + 31) / 32, 256, 0, st>>>(partial, slices, q.I, g.Cs, g.Cb, dw);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

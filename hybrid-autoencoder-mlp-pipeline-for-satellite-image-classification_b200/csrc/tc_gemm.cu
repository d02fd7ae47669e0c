// tcgen05 / TMEM implicit-GEMM kernels for the six 32..256-channel stride-2 layers (sm_100a only).
//
//   row GEMM  D[128 pixels x NT channels] (+)= A[128 x K] * B[NT x K]^T      (Conv2d fwd / ConvT dgrad: FAM_FPROP,
//                                                                              ConvT fwd / Conv2d dgrad: FAM_DGRAD)
//     A : gathered from the fp32 NHWC activation by 8 producer warps which apply the operand transform
//         (raw | BatchNorm+ReLU | BatchNorm-backward), round to bf16 (1 or 2 split terms) and store 128-byte
//         K-major rows into shared memory in the SWIZZLE_128B pattern the tensor core reads;
//     B : weights, pre-packed (tc_pack_conv) as ready-to-use swizzled tiles; one TMA bulk copy (cp.async.bulk,
//         mbarrier complete_tx) per pipeline stage;
//     D : fp32 accumulator in tensor memory; one thread issues tcgen05.mma, tcgen05.commit frees the stage;
//     epilogue : tcgen05.ld -> bias / ReLU-mask -> global store + per-channel BatchNorm statistics
//         (warp-shuffle transposed reduction -> shared memory -> one fp64 atomic per channel per CTA).
//
// Precision: AE_PREC_FP32 splits every operand element into hi + lo bf16 terms and issues hi*hi + hi*lo + lo*hi
// (3 MMAs per k-step, fp32 accumulate): ~2^-17 relative error per product, i.e. fp32-class results.
#include "common.cuh"

namespace ae {

static constexpr int TC_THREADS = 288;      // 8 producer/epilogue warps + 1 MMA warp
static constexpr int TC_PRODUCERS = 256;
static constexpr int TILE_M = 128;
static constexpr int KCHUNK = 64;           // bf16 elements per 128-byte smem row

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("ae_b200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_128B, 8-row groups 1024 bytes apart
//   K-major : rows are M/N indices, 128 bytes of K per row              (lbo unused -> 1)
//   MN-major: rows are K indices, 128 bytes (64 elements) of M/N per row (lbo = byte stride between 64-element groups)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset, bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, M=128, N=n; majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// v[8] -> hi (and lo) bf16 octets
template <int NSPLIT>
__device__ __forceinline__ void split_store(const float (&v)[8], uint8_t* plane0, uint32_t plane_stride, uint32_t off) {
  uint4 hi;
  hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(plane0 + off) = hi;
  if (NSPLIT == 2) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
    uint4 lo;
    lo.x = pack_bf16x2(r[0], r[1]); lo.y = pack_bf16x2(r[2], r[3]); lo.z = pack_bf16x2(r[4], r[5]); lo.w = pack_bf16x2(r[6], r[7]);
    *reinterpret_cast<uint4*>(plane0 + plane_stride + off) = lo;
  }
}

// 8 consecutive channels of an activation operand
__device__ __forceinline__ void load_operand8(const Operand& op, size_t off, int c, bool valid, float (&v)[8]) {
  const float4 a = load_operand4(op, off, c, valid);
  const float4 b = load_operand4(op, off + 4, c + 4, valid);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// lane l ends with the sum over the warp's 32 lanes of element v[l]
__device__ __forceinline__ float warp_colsum32_tc(float (&v)[32], int lane) {
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

// ---------------------------------------------------------------------------------------------
// Row GEMM kernel
// ---------------------------------------------------------------------------------------------
struct TcRow {
  RowGemm r;
  const uint8_t* packed;   // weight tiles
  int kc_total;            // 64-wide K chunks per n-tile in the pack (all phases)
};

template <int FAMILY, int NT, int NSPLIT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_rowgemm(TcRow q) {
  constexpr uint32_t A_PLANE = TILE_M * 128;               // bytes
  constexpr uint32_t B_PLANE = NT * 128;
  constexpr uint32_t STAGE_BYTES = NSPLIT * (A_PLANE + B_PLANE);
  constexpr uint32_t B_BYTES = NSPLIT * B_PLANE;
  constexpr int NMMA = NSPLIT == 2 ? 3 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;
  __shared__ float sStat[2][NT];

  const RowGemm& p = q.r;
  const Geom g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TILE_M;
  const int ntile = blockIdx.y;
  int py = 0, px = 0, kc_off = 0, nkc;
  if (FAMILY == FAM_DGRAD) {
    const int phase = blockIdx.z;
    py = phase >> 1; px = phase & 1;
    const int cpk = g.Cs / KCHUNK;                         // chunks per tap
    nkc = (1 + py) * (1 + px) * cpk;
    kc_off = (phase == 0 ? 0 : phase == 1 ? 1 : phase == 2 ? 3 : 5) * cpk;
  } else {
    nkc = (p.K + KCHUNK - 1) / KCHUNK;
  }
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), TC_PRODUCERS); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (tid < NT) { sStat[0][tid] = 0.f; sStat[1][tid] = 0.f; }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_slot), NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 8) {
    // ===================== producers =====================
    const int chunk = tid & 7;
    int rn[4], ry[4], rx[4];
    bool rok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + (tid >> 3) + 32 * i;
      rok[i] = m < p.M;
      rx[i] = m & (g.Ws - 1);
      ry[i] = (m >> g.lWs) & (g.Hs - 1);
      rn[i] = m >> (g.lWs + g.lHs);
    }
    const int C = FAMILY == FAM_FPROP ? g.Cb : g.Cs;
    const uint8_t* bsrc = q.packed + ((size_t)ntile * q.kc_total + kc_off) * B_BYTES;
    for (int it = 0; it < nkc; ++it) {
      const int s = it % STAGES, round = it / STAGES;
      if (round > 0) mbar_wait(empty_bar(s), (round - 1) & 1);
      uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
      if (tid == 0) {
        mbar_expect_tx(full_bar(s), B_BYTES);
        bulk_copy_g2s(smem_u32(stage + NSPLIT * A_PLANE), bsrc + (size_t)it * B_BYTES, B_BYTES, full_bar(s));
      }
      const int k = it * KCHUNK + chunk * 8;
      const int tap = k / C, c = k - tap * C;
      int dy = 0, dx = 0;
      bool tap_ok = true;
      if (FAMILY == FAM_FPROP) {
        tap_ok = tap < 9;
        dy = tap / 3 - 1; dx = tap - (tap / 3) * 3 - 1;           // source = (2*oy + dy, 2*ox + dx)
      } else {
        const int a = tap / (1 + px), b = tap - a * (1 + px);
        dy = (py && a == 0) ? 1 : 0; dx = (px && b == 0) ? 1 : 0;   // source = (j + dy, i + dx)
      }
      float v[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bool valid = rok[i] && tap_ok;
        size_t off;
        if (FAMILY == FAM_FPROP) {
          const int iy = 2 * ry[i] + dy, ix = 2 * rx[i] + dx;
          valid = valid && iy >= 0 && ix >= 0;
          off = (((size_t)rn[i] * (2 * g.Hs) + iy) * (2 * g.Ws) + ix) * g.Cb + c;
        } else {
          const int sy = ry[i] + dy, sx = rx[i] + dx;
          valid = valid && sy < g.Hs && sx < g.Ws;
          off = (((size_t)rn[i] * g.Hs + sy) * g.Ws + sx) * g.Cs + c;
        }
        load_operand8(p.A, off, c, valid, v[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = (tid >> 3) + 32 * i;
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
        split_store<NSPLIT>(v[i], stage, A_PLANE, off);
      }
      fence_proxy_async();
      mbar_arrive(full_bar(s));
    }
    // ===================== epilogue =====================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int qd = warp & 3, half = warp >> 2;
    constexpr int COLS_PER_HALF = NT >= 64 ? NT / 2 : NT;
    const bool active = NT >= 64 || half == 0;
    const int row = qd * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const Epilogue e = p.epi;
    size_t orow = 0;
    if (row_ok) {
      if (FAMILY == FAM_DGRAD) {
        const int x = m & (g.Ws - 1), y = (m >> g.lWs) & (g.Hs - 1), nn = m >> (g.lWs + g.lHs);
        orow = (((size_t)nn * (2 * g.Hs) + 2 * y + py) * (2 * g.Ws) + 2 * x + px) * p.N;
      } else {
        orow = (size_t)m * p.N;
      }
    }
    if (active) {
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_HALF; cc += 32) {
        const int col0 = half * COLS_PER_HALF + cc;      // column inside the tile
        const int n = ntile * NT + col0;                 // global output channel
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)col0, v);
        float s2[32];
        if (e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok) yv = __ldg(reinterpret_cast<const float4*>(e.y + orow + n) + j4);
            const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = (n + j4 * 4 + j) % e.C;
              const float z = fmaf(ya[j], __ldg(e.bnc + AE_BNC_SCALE * e.C + ch), __ldg(e.bnc + AE_BNC_SHIFT * e.C + ch));
              float d = (row_ok && z > 0.f) ? v[j4 * 4 + j] : 0.f;
              v[j4 * 4 + j] = d;
              s2[j4 * 4 + j] = d * ((ya[j] - __ldg(e.bnc + AE_BNC_MEAN * e.C + ch)) * __ldg(e.bnc + AE_BNC_RSTD * e.C + ch));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float d = v[j] + (e.bias ? __ldg(e.bias + n + j) : 0.f);
            d = row_ok ? d : 0.f;
            v[j] = d;
            s2[j] = d * d;
          }
        }
        if (row_ok) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            reinterpret_cast<float4*>(p.out + orow + n)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
        if (e.mode != AE_EPI_STORE && e.stats) {
          const float a = warp_colsum32_tc(v, lane);
          const float b = warp_colsum32_tc(s2, lane);
          atomicAdd(&sStat[0][col0 + lane], a);
          atomicAdd(&sStat[1][col0 + lane], b);
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (e.mode != AE_EPI_STORE && e.stats && tid < NT) {
      const int ch = (ntile * NT + tid) % e.C;
      atomicAdd(e.stats + ch, (double)sStat[0][tid]);
      atomicAdd(e.stats + e.C + ch, (double)sStat[1][tid]);
    }
  } else {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(NT, 0, 0);
      for (int it = 0; it < nkc; ++it) {
        const int s = it % STAGES, round = it / STAGES;
        mbar_wait(full_bar(s), round & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t b0 = a0 + NSPLIT * A_PLANE;
#pragma unroll
        for (int kk = 0; kk < KCHUNK / 16; ++kk) {
          const uint64_t ah = make_desc(a0 + kk * 32, 16), bh = make_desc(b0 + kk * 32, 16);
          umma_bf16(tmem_base, ah, bh, idesc, (it | kk) != 0);
          if (NMMA == 3) {
            const uint64_t al = make_desc(a0 + A_PLANE + kk * 32, 16), bl = make_desc(b0 + B_PLANE + kk * 32, 16);
            umma_bf16(tmem_base, ah, bl, idesc, 1);
            umma_bf16(tmem_base, al, bh, idesc, 1);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NT);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline int nt_for(int n) { return n >= 64 ? 64 : 32; }

static inline int fwd_chunks(int Cb) { return (9 * Cb + KCHUNK - 1) / KCHUNK; }

size_t tc_packed_bytes(int Cs, int Cb, int nsplit) {
  const size_t f = (size_t)Cs * fwd_chunks(Cb) * 128 * nsplit;
  const size_t d = (size_t)Cb * (9 * Cs / KCHUNK) * 128 * nsplit;
  return (f > d ? f : d) + 1024;
}

bool tc_rowgemm_supported(const RowGemm& p) {
  if (p.family == FAM_FPROP) return p.g.Cb % 32 == 0 && p.N % 32 == 0 && p.splitK <= 1 && is_pow2(p.g.Cb);
  if (p.family == FAM_DGRAD) return p.g.Cs % 64 == 0 && p.N % 32 == 0 && is_pow2(p.g.Cs);
  return false;
}

// w [Cs][Cb][3][3] fp32 -> swizzled bf16 (hi[, lo]) tiles for both GEMM orientations.
// Tile (n_tile, kc): [plane][NT rows][128 bytes], element (r, j) at r*128 + (((j>>3) ^ (r&7)) << 4) + (j&7)*2.
__global__ void k_tc_pack_conv(const float* __restrict__ w, int Cs, int Cb, int nsplit, uint8_t* __restrict__ fwd,
                               uint8_t* __restrict__ dgrad) {
  const int KCf = (9 * Cb + KCHUNK - 1) / KCHUNK, NTf = Cs >= 64 ? 64 : 32;
  const int KCd = 9 * Cs / KCHUNK, NTd = Cb >= 64 ? 64 : 32;
  const int nf = Cs * KCf * KCHUNK, nd = Cb * KCd * KCHUNK;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nf + nd; idx += gridDim.x * blockDim.x) {
    float v = 0.f;
    uint8_t* base;
    size_t tile_bytes;
    int r, j;
    size_t tile;
    int NT;
    if (idx < nf) {
      j = idx % KCHUNK;
      const int kc = (idx / KCHUNK) % KCf;
      const int n = idx / (KCHUNK * KCf);             // cs
      const int k = kc * KCHUNK + j;
      if (k < 9 * Cb) { const int tap = k / Cb, cb = k - tap * Cb; v = w[((size_t)n * Cb + cb) * 9 + tap]; }
      NT = NTf; r = n % NT; tile = (size_t)(n / NT) * KCf + kc; base = fwd;
    } else {
      const int i2 = idx - nf;
      j = i2 % KCHUNK;
      const int kc = (i2 / KCHUNK) % KCd;
      const int n = i2 / (KCHUNK * KCd);              // cb
      const int k = kc * KCHUNK + j;                  // global k over the 4 stacked phases: (tap slot, cs)
      const int slot = k / Cs, cs = k - slot * Cs;    // slot 0: phase 0; 1-2: phase 1; 3-4: phase 2; 5-8: phase 3
      int ky, kx;
      if (slot == 0) { ky = 1; kx = 1; }
      else if (slot <= 2) { ky = 1; kx = slot == 1 ? 0 : 2; }
      else if (slot <= 4) { ky = slot == 3 ? 0 : 2; kx = 1; }
      else { const int t = slot - 5; ky = (t >> 1) ? 2 : 0; kx = (t & 1) ? 2 : 0; }
      v = w[((size_t)cs * Cb + n) * 9 + ky * 3 + kx];
      NT = NTd; r = n % NT; tile = (size_t)(n / NT) * KCd + kc; base = dgrad;
    }
    tile_bytes = (size_t)nsplit * NT * 128;
    const size_t off = tile * tile_bytes + (size_t)r * 128 + (size_t)((((j >> 3) ^ (r & 7)) << 4) + (j & 7) * 2);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    if (nsplit == 2) *reinterpret_cast<__nv_bfloat16*>(base + off + (size_t)NT * 128) = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

int tc_pack_conv(const float* w, int Cs, int Cb, int nsplit, void* fwd, void* dgrad, cudaStream_t st) {
  AE_CHECK(Cs % 64 == 0 && Cb % 32 == 0, "tc_pack_conv: Cs=%d must be a multiple of 64 and Cb=%d of 32", Cs, Cb);
  AE_CHECK((((uintptr_t)fwd | (uintptr_t)dgrad) & 15) == 0, "tc_pack_conv: packed buffers must be 16-byte aligned");
  const int total = Cs * fwd_chunks(Cb) * KCHUNK + Cb * (9 * Cs / KCHUNK) * KCHUNK;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_tc_pack_conv<<<blocks, 256, 0, st>>>(w, Cs, Cb, nsplit, (uint8_t*)fwd, (uint8_t*)dgrad);
  AE_LAUNCH_CHECK();
  return 0;
}

template <int FAMILY, int NT, int NSPLIT, int STAGES>
static int launch_row(const TcRow& q, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * NSPLIT * (TILE_M * 128 + NT * 128) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    AE_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<FAMILY, NT, NSPLIT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_tc_rowgemm<FAMILY, NT, NSPLIT, STAGES><<<grid, TC_THREADS, smem, st>>>(q);
  AE_LAUNCH_CHECK();
  return 0;
}

int tc_rowgemm(const RowGemm& p, const void* packed, int nsplit, cudaStream_t st) {
  AE_CHECK(tc_rowgemm_supported(p), "tc_rowgemm: unsupported shape");
  AE_CHECK(((uintptr_t)packed & 15) == 0, "tc_rowgemm: packed weights must be 16-byte aligned");
  TcRow q;
  q.r = p;
  q.packed = (const uint8_t*)packed;
  const int NT = nt_for(p.N);
  dim3 grid((p.M + TILE_M - 1) / TILE_M, p.N / NT, 1);
  if (p.family == FAM_FPROP) {
    q.kc_total = fwd_chunks(p.g.Cb);
    if (NT == 64) return nsplit == 2 ? launch_row<FAM_FPROP, 64, 2, 3>(q, grid, st) : launch_row<FAM_FPROP, 64, 1, 4>(q, grid, st);
    return nsplit == 2 ? launch_row<FAM_FPROP, 32, 2, 3>(q, grid, st) : launch_row<FAM_FPROP, 32, 1, 4>(q, grid, st);
  }
  q.kc_total = 9 * p.g.Cs / KCHUNK;
  grid.z = 4;
  if (NT == 64) return nsplit == 2 ? launch_row<FAM_DGRAD, 64, 2, 3>(q, grid, st) : launch_row<FAM_DGRAD, 64, 1, 4>(q, grid, st);
  return nsplit == 2 ? launch_row<FAM_DGRAD, 32, 2, 3>(q, grid, st) : launch_row<FAM_DGRAD, 32, 1, 4>(q, grid, st);
}

int tc_wgrad(const ColGemm& p, int nsplit, cudaStream_t st) {
  (void)nsplit;
  return simt_colgemm(p, st);   // tcgen05 weight-gradient kernel: see tc_wgrad.cu
}

}  // namespace ae

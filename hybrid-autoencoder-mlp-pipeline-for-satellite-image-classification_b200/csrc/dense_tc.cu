// Flatten + Linear(4096, latent) (NB:520-521) on tcgen05 for the eval-mode encoder: the last convolution's epilogue already
// emitted relu(bn(conv4)) as split-bf16 planes in NHWC flatten order, which IS a K-major [batch][4096] operand.
//   partial[ks][m][n] = sum over the K slice ks of A[m][k] * W[n][k]      (fp32 accumulation in tensor memory)
// One CTA per (128-row tile, K slice); the fixed-order reduce_partials adds the slices and the bias.
//   warp 0: TMA producer (one tensor load for the activation planes + one bulk copy for the weight tile per 64-deep K chunk)
//   warp 1: MMA issuer (hi*hi + hi*lo + lo*hi per 16-deep step in fp32 mode)
//   warps 2-5: accumulator -> partial buffer
// HBM bound by the activation planes (batch * 4096 * 4 bytes in fp32 mode, read once).
#include "tc_common.cuh"
#include "tma_host.cuh"
#include "pack.cuh"

namespace ae {

static constexpr int DT_THREADS = 192;
static constexpr int DT_MAXSTAGES = 6;
static constexpr int DT_KC = 64;                       // K elements per chunk: 128-byte swizzled rows

struct DenseTc {
  CUtensorMap amap;          // [nsplit][M][K] bf16 planes, box (64 k, 128 rows, nsplit)
  const uint8_t* wpack;      // [K/64 chunks][nsplit][N rows][128 bytes], SWIZZLE_128B
  float* partial;            // [ksplit][M][N]
  int M, N, chunks, ksplit, stages;
  uint32_t tmem_cols;
};

template <int NSPLIT>
__global__ void __launch_bounds__(DT_THREADS, 1) k_dense_tc(const __grid_constant__ DenseTc q) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * DT_MAXSTAGES + 1];
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t A_PLANE = TILE_M * 128, A_BYTES = NSPLIT * A_PLANE;
  const uint32_t W_PLANE = (uint32_t)q.N * 128, W_BYTES = NSPLIT * W_PLANE, STAGE = A_BYTES + W_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TILE_M, ks = blockIdx.y;
  // K chunks of this slice: the first (chunks % ksplit) slices hold one more
  const int per = q.chunks / q.ksplit, extra = q.chunks - per * q.ksplit;
  const int c_begin = ks * per + (ks < extra ? ks : extra), c_count = per + (ks < extra ? 1 : 0);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (DT_MAXSTAGES + s); };
  const uint32_t done = bar0 + 8u * (2 * DT_MAXSTAGES);
  if (tid == 0) {
    tma_prefetch_desc(&q.amap);
    for (int s = 0; s < DT_MAXSTAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), q.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t s_u32 = smem_u32(smem);

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0, round = 0;
      for (int i = 0; i < c_count; ++i) {
        if (round > 0) mbar_wait(empty(slot), (round - 1) & 1);
        const uint32_t dst = s_u32 + (uint32_t)slot * STAGE;
        mbar_arrive_expect_tx(full(slot), STAGE);
        tma_load_5d(dst, &q.amap, (c_begin + i) * DT_KC, 0, 0, m0, 0, full(slot));
        bulk_copy_g2s(dst + A_BYTES, q.wpack + (size_t)(c_begin + i) * W_BYTES, W_BYTES, full(slot));
        if (++slot == q.stages) { slot = 0; ++round; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(q.N, 0, 0);
      const uint64_t desc0 = make_desc(s_u32, 16, 1024, 128);
      int slot = 0, round = 0;
      for (int i = 0; i < c_count; ++i) {
        mbar_wait(full(slot), round & 1);
        tc_fence_after();
        const uint64_t da = desc0 + (uint64_t)(((uint32_t)slot * STAGE) >> 4), dw = da + (uint64_t)(A_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < DT_KC / 16; ++kk) {
          const uint64_t ah = da + (uint64_t)(kk * 2), bh = dw + (uint64_t)(kk * 2);
          umma_bf16(tmem_base, ah, bh, idesc, !(i == 0 && kk == 0));
          if (NSPLIT == 2) {
            umma_bf16(tmem_base, ah, bh + (uint64_t)(W_PLANE >> 4), idesc, 1);
            umma_bf16(tmem_base, ah + (uint64_t)(A_PLANE >> 4), bh, idesc, 1);
          }
        }
        umma_commit(empty(slot));
        if (++slot == q.stages) { slot = 0; ++round; }
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;                               // a warp reads the TMEM lanes 32 * (warp % 4) ..
    const int row = m0 + qd * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    float* dst = q.partial + ((size_t)ks * q.M + (size_t)(row < q.M ? row : 0)) * q.N;
    for (int c = 0; c < q.N; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c, v);
      if (row < q.M) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8)
          if (c + j8 * 8 < q.N) st_global_v8(dst + c + j8 * 8, v, j8);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, q.tmem_cols);
  }
}

__global__ void __launch_bounds__(256) k_pack_dense_tc(const float* __restrict__ w, int N, int K, int permC, int permHW, int nsplit,
                                                       uint8_t* __restrict__ dst) {
  const int idx8 = blockIdx.x * 256 + threadIdx.x;
  if (idx8 < N * (K / 8)) pack_dense_tc_chunk(idx8, w, N, K, permC, permHW, nsplit, dst);
}

bool dense_tc_supported(int N, int K) { return N >= 16 && N <= 256 && N % 16 == 0 && K % DT_KC == 0 && K >= DT_KC; }
size_t dense_tc_pack_bytes(int N, int K, int nsplit) { return (size_t)N * K * 2 * nsplit; }

// w: torch Linear weight [N][K]; permC > 0: the layer's input is an NHWC flatten of a (permC, permHW) feature map (pack.cuh)
int dense_tc_pack(const float* w, int N, int K, int permC, int permHW, int nsplit, void* pack, cudaStream_t st) {
  AE_CHECK(dense_tc_supported(N, K), "dense_tc_pack: N=%d K=%d not supported", N, K);
  const int total = N * (K / 8);
  k_pack_dense_tc<<<(total + 255) / 256, 256, 0, st>>>(w, N, K, permC, permHW, nsplit, static_cast<uint8_t*>(pack));
  AE_LAUNCH_CHECK();
  return 0;
}

// split count: enough CTAs to fill the chip, every slice at least 4 chunks deep
int dense_tc_split(int M, int K) {
  const int tiles = (M + TILE_M - 1) / TILE_M, chunks = K / DT_KC;
  int ks = 148 / tiles;
  if (ks > chunks / 4) ks = chunks / 4;
  return ks < 1 ? 1 : ks;
}

// out_partial: [ksplit][M][N] fp32 (ksplit = dense_tc_split(M, K)); the caller reduces it (reduce_partials adds the bias)
int dense_tc(const void* a_planes, const void* pack, int M, int N, int K, int nsplit, float* out_partial, size_t partial_bytes,
             int* ksplit_out, cudaStream_t st) {
  AE_CHECK(dense_tc_supported(N, K), "dense_tc: N=%d K=%d not supported", N, K);
  AE_CHECK(nsplit == 1 || nsplit == 2, "dense_tc: nsplit=%d", nsplit);
  DenseTc q{};
  q.M = M; q.N = N; q.chunks = K / DT_KC; q.ksplit = dense_tc_split(M, K);
  AE_CHECK((size_t)q.ksplit * M * N * 4 <= partial_bytes, "dense_tc: partial buffer too small");
  q.wpack = static_cast<const uint8_t*>(pack); q.partial = out_partial;
  uint32_t cols = 32;
  while (cols < (uint32_t)((N + 31) & ~31)) cols <<= 1;
  q.tmem_cols = cols;
  const size_t stage = (size_t)nsplit * (TILE_M * 128 + (size_t)N * 128);
  int stages = (int)((200 * 1024) / stage);
  if (stages > DT_MAXSTAGES) stages = DT_MAXSTAGES;
  AE_CHECK(stages >= 2, "dense_tc: N=%d does not leave two stages of shared memory", N);
  q.stages = stages;
  AE_TRY(encode_map(&q.amap, a_planes, M, 1, 1, K, nsplit, 0, 0, 1, 1, 1, 1, DT_KC, 1, 1, TILE_M));
  const size_t smem = stage * stages + 1024;
  const dim3 grid((M + TILE_M - 1) / TILE_M, q.ksplit);
  if (nsplit == 2) {
    static bool attr2 = false;
    if (!attr2) { AE_CUDA(cudaFuncSetAttribute(k_dense_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024)); attr2 = true; }
    k_dense_tc<2><<<grid, DT_THREADS, smem, st>>>(q);
  } else {
    static bool attr1 = false;
    if (!attr1) { AE_CUDA(cudaFuncSetAttribute(k_dense_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024)); attr1 = true; }
    k_dense_tc<1><<<grid, DT_THREADS, smem, st>>>(q);
  }
  AE_LAUNCH_CHECK();
  if (ksplit_out) *ksplit_out = q.ksplit;
  return 0;
}

}  // namespace ae

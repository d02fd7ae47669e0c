// Row GEMM, second generation: the training-mode launches of the six 32..256-channel stride-2 layers (sm_100a only).
//
// Same operands, weight packs and arithmetic as k_tma_rowgemm (tma_gemm.cu); what changed is decided from per-CTA phase
// timestamps of the first generation (scripts/trace_rowgemm.py, profiles/r2_trace_*): its launches were bound by (1) the
// tensor pipe reading both operands of a 128x64x16 MMA from shared memory (6 KB per MMA: ~69 clk instead of the 32 clk
// floor), (2) unbalanced tiles (the four output-parity phases of the transposed convolution were tiles of 1, 2, 2 and 4
// taps), (3) an epilogue of 32-byte-per-lane global stores (~1.1 us per 32 columns) and latency-exposed loads.
//
//   * One CTA per SM, the shared memory as two rings: activation boxes and weight-tile groups are loaded and released
//     independently.
//   * FAM_DGRAD (ConvTranspose2d forward / Conv2d data gradient): the four output-parity phases of an m-tile are ONE tile
//     with four accumulators in tensor memory, in column order [phase 0, 1, 3, 2].  They share the four distinct shifted
//     boxes of the small image: 4 boxes per K chunk instead of 9, and every tile costs the same.  Phases that read the
//     same box sit in adjacent accumulator columns, so their weight tiles are stacked along N and issued as ONE MMA of
//     N = 2 x tile width: shift (0,0) -> [ph0,ph1] and [ph3,ph2], shift (0,1) -> [ph1,ph3], shift (1,0) -> [ph3,ph2],
//     shift (1,1) -> [ph3].  96 MMAs per tile and K chunk pair instead of 216, each reading 8 KB for twice the work.
//   * FAM_FPROP with N % 128 == 0 and enough tiles: 128-wide tiles from two stacked 64-row weight tiles (same pack).
//   * Layers whose weight groups of an n-tile fit in the ring keep them resident for the whole launch.
//   * Epilogue: 8 warps (two per TMEM lane quadrant).  A warp owns a 32-row x 32-column block: accumulator out of
//     tensor memory, transform, rows into a 128-byte-swizzled shared-memory block, ONE TMA tensor store per block
//     (the output is a tensor map: a plain NHWC tensor, or one pixel lattice per output phase).  The forward activation
//     y of the ReLU-backward epilogue arrives the same way (TMA load issued one block ahead), and the per-channel
//     BatchNorm sums are read column-wise out of the staged block (lane = column) instead of 62 shuffles per block.
//   * The producer and the MMA issuer are single threads: no integer division, descriptors advanced by addition.
#include <cstdlib>

#include "tc_common.cuh"
#include "tma_host.cuh"

namespace ae {

static constexpr int R2_THREADS = 320;   // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..9: epilogue
static constexpr int R2_MAXRING = 12;
static constexpr int R2_BLOCK_BYTES = 4096;        // one staged 32 x 32 fp32 block; 8 output blocks (+ 2 x 8 y blocks: ReLU backward)

#ifdef AE_TRACE
__device__ unsigned long long* g_trace2 = nullptr;
__device__ __forceinline__ unsigned long long gtimer2() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define AE_TR2(slot) do { if (g_trace2) g_trace2[(size_t)blockIdx.x * 16 + (slot)] = gtimer2(); } while (0)
#define AE_CY(var, expr) do { const long long _c = clock64(); expr; var += clock64() - _c; } while (0)
#else
#define AE_TR2(slot) do { } while (0)
#define AE_CY(var, expr) do { expr; } while (0)
#endif

struct alignas(64) TmaRow2 {
  CUtensorMap amap[4];      // FPROP: parity lattices (py*2+px) of the big image; DGRAD: [0] = small image (bf16 planes)
  CUtensorMap omap[4];      // output, fp32: FPROP [0]; DGRAD: the lattice of output phase py*2+px
  CUtensorMap ymap[4];      // AE_EPI_RELUBWD_STATS: the forward activation y, same geometry as omap
  const uint8_t* wtiles;    // packed weight tiles
  Geom g;
  Epilogue epi;
  int M, N;
  int cpt;                  // K chunks per tap (C / KC)
  int wchunks;              // K chunks per n-tile in the weight pack
  int na, nw;               // ring slots
  int ny;                   // forward-activation blocks per epilogue warp (ReLU backward): 2 when shared memory allows
  int resident;             // the (single) n-tile's weight groups are loaded once and stay in the ring
  int bx, by, bn;           // pixel box of one 128-row tile
  BnJob job;                // kind != BN_JOB_NONE: coefficient job of epi.stats, run by the last CTA to finish
  unsigned int* job_counter;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// DGRAD: the weight-tile groups of shift s = dy*2+dx (source pixel (y+dy, x+dx)).  A group is one MMA: `ntl` tiles stacked
// along N, written to the accumulator column blocks col, col+1.  Column order of the phases: [0, 1, 3, 2].  `s0`, `s1`:
// the tiles' slots in the phase-stacked weight pack (slot 0: phase 0; 1-2: phase 1; 3-4: phase 2; 5-8: phase 3).
//   shift (0,0): [ph0 tap 0 | ph1 tap 1] -> columns 0-1,   [ph3 tap 3 | ph2 tap 1] -> columns 2-3
//   shift (0,1): [ph1 tap 0 | ph3 tap 2] -> columns 1-2
//   shift (1,0): [ph3 tap 1 | ph2 tap 0] -> columns 2-3
//   shift (1,1): [ph3 tap 0]             -> column 2
struct DgGroup { int col, ntl, idx; };            // idx: the group's position in the weight pack's 9-tile block (pack.cuh)
__device__ __forceinline__ int dg_groups(int s) { return s == 0 ? 2 : 1; }
__device__ __forceinline__ DgGroup dg_group(int s, int gi) {
  if (s == 0) return gi == 0 ? DgGroup{0, 2, 0} : DgGroup{2, 2, 1};
  if (s == 1) return DgGroup{1, 2, 2};
  if (s == 2) return DgGroup{2, 2, 3};
  return DgGroup{2, 1, 4};
}
// phase `ph` is complete after shift `ph`; its accumulator column block
__device__ __forceinline__ int dg_col_of_phase(int ph) { return ph == 2 ? 3 : ph == 3 ? 2 : ph; }

// WT: rows of one packed weight tile (the pack's n-tile width).  FPROP: the output tile is NSUB stacked weight tiles wide;
// DGRAD: NSUB = 1, every phase has a WT-wide accumulator.
template <int FAMILY, int WT, int KC, int NSPLIT, int NSUB>
__global__ void __launch_bounds__(R2_THREADS, 1) k_rowgemm2(const __grid_constant__ TmaRow2 q) {
  constexpr int ROWB = KC * 2;
  constexpr int GT = FAMILY == FAM_DGRAD ? 2 : NSUB;        // weight tiles per group (ring slot)
  constexpr uint32_t A_PLANE = TILE_M * ROWB, W_PLANE = WT * ROWB;
  constexpr uint32_t WTILE_BYTES = NSPLIT * W_PLANE;        // one tile in the weight pack: [hi plane][lo plane]
  constexpr uint32_t A_SLOT = NSPLIT * A_PLANE;
  constexpr uint32_t W_SLOT = GT * NSPLIT * W_PLANE;        // [hi: GT tiles][lo: GT tiles]
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int NPH = FAMILY == FAM_DGRAD ? 4 : 1;          // accumulators per tile
  constexpr int NS = FAMILY == FAM_DGRAD ? 4 : 9;           // activation boxes per K chunk
  constexpr int PW = FAMILY == FAM_DGRAD ? WT : WT * NSUB;  // accumulator width = output channels per n-tile
  constexpr int CB = PW / 32;                               // 32-column blocks per accumulator
  constexpr int UNITS = NPH * CB;                           // epilogue blocks per tile and lane quadrant
  constexpr uint32_t ACC_COLS = NPH * PW;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;              // double-buffered: 128 / 256 / 512 columns
  constexpr int EPI_WARPS = UNITS >= 2 ? 8 : 4;
  constexpr int MAXN = FAMILY == FAM_DGRAD ? 128 : 256;     // output channels (rowgemm2_supported)
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[4 * R2_MAXRING + 26];
  __shared__ uint32_t tmem_slot;
  __shared__ float sStat[2][MAXN];
  __shared__ __align__(16) float sCoef[MAXN][4];

  const Geom g = q.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) AE_TR2(0);
  const int na = q.na, nw = q.nw;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (size_t)na * A_SLOT;
  uint8_t* sStage = sW + (size_t)nw * W_SLOT;               // 8 output blocks, then 8 y blocks
  const int tiles_m = (q.M + TILE_M - 1) / TILE_M, tiles_n = q.N / PW;
  const int num_tiles = tiles_m * tiles_n;
  const bool resident = q.resident != 0;

  const uint32_t bar0 = smem_u32(&bars[0]);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (R2_MAXRING + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * R2_MAXRING + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (3 * R2_MAXRING + s); };
  auto tfull = [&](int b, int ph) { return bar0 + 8u * (4 * R2_MAXRING + b * 4 + ph); };
  auto tempty = [&](int b) { return bar0 + 8u * (4 * R2_MAXRING + 8 + b); };
  auto ybar = [&](int w) { return bar0 + 8u * (4 * R2_MAXRING + 10 + w); };   // w in [0, 16): warp + 8 * y slot

  if (tid == 0) {
    if (FAMILY == FAM_FPROP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tma_prefetch_desc(&q.amap[i]);
    } else {
      tma_prefetch_desc(&q.amap[0]);
    }
#pragma unroll
    for (int i = 0; i < NPH; ++i) tma_prefetch_desc(&q.omap[i]);
    for (int s = 0; s < R2_MAXRING; ++s) {
      mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      for (int ph = 0; ph < 4; ++ph) mbar_init(tfull(b, ph), 1);
      mbar_init(tempty(b), EPI_WARPS);
    }
    for (int w = 0; w < 16; ++w) mbar_init(ybar(w), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) AE_TR2(1);

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    // Measured (scripts/trace_rowgemm.py): ~100 clk per cp.async.bulk / tensor load, a 256-row activation
    // box keeps the TMA unit busy ~500 clk.  Neither the number of copies (group pack: one per weight group), the ring depth
    // nor the order of the K chunks changes the kernel's time: this thread is not what paces the main loop.  Spreading the
    // weight-plane copies over the lanes of the warp was tried and is slower.
    if (lane == 0) {
      const int P = g.Hs * g.Ws;
      // ring positions advance incrementally: this is one thread, every integer division would be on the critical path
      uint32_t aslot = 0, around = 0, wslot_r = 0, wround = 0;
      bool first = true, traced = false;
      const uint32_t sA_u32 = smem_u32(sA), sW_u32 = smem_u32(sW);
#ifdef AE_TRACE
      unsigned long long* tr_prod = (g_trace2 && blockIdx.x == 0) ? g_trace2 + 16384 + 256 : nullptr;
      long long pc_wait_a = 0, pc_wait_w = 0, pc_t0 = clock64();
#endif
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, first = false) {
        const int mt = tile / tiles_n, nt = tile - mt * tiles_n;   // n-tile fastest: concurrent CTAs share A boxes through L2
        const int m0 = mt * TILE_M;
        const int n0 = m0 >> (g.lHs + g.lWs);
        const int y0 = (m0 & (P - 1)) >> g.lWs;
        // first weight tile of this n-tile in the pack (FPROP: NSUB consecutive pack n-tiles; DGRAD: cpt blocks of 9 tiles)
        const uint8_t* wsrc = q.wtiles + (size_t)nt * (FAMILY == FAM_DGRAD ? 1 : NSUB) * q.wchunks * WTILE_BYTES;
        uint32_t wt = 0;
        for (int s = 0; s < NS; ++s) {
          int ax, ay, amap_i;
          if (FAMILY == FAM_FPROP) {
            const int ky = s / 3, kx = s - ky * 3;
            amap_i = ((ky + 1) & 1) * 2 + ((kx + 1) & 1);             // parity lattice of source row 2*oy-1+ky
            ax = (kx == 0) ? -1 : 0; ay = y0 + ((ky == 0) ? -1 : 0);
          } else {
            amap_i = 0; ax = s & 1; ay = y0 + (s >> 1);
          }
          const int groups = FAMILY == FAM_DGRAD ? dg_groups(s) : 1;
          for (int c = 0; c < q.cpt; ++c) {
            if (around > 0) AE_CY(pc_wait_a, mbar_wait(a_empty(aslot), (around - 1) & 1));
            mbar_arrive_expect_tx(a_full(aslot), A_SLOT);
            tma_load_5d(sA_u32 + aslot * A_SLOT, &q.amap[amap_i], c * KC, ax, ay, n0, 0, a_full(aslot));
            if (!traced) { AE_TR2(2); traced = true; }
            if (++aslot == (uint32_t)na) { aslot = 0; ++around; }
            for (int gi = 0; gi < groups; ++gi, ++wt) {
              uint32_t wslot;
              if (resident) {
                if (!first) continue;
                wslot = wt;
              } else {
                wslot = wslot_r;
                if (wround > 0) AE_CY(pc_wait_w, mbar_wait(w_empty(wslot), (wround - 1) & 1));
                if (++wslot_r == (uint32_t)nw) { wslot_r = 0; ++wround; }
              }
              const uint32_t w_dst = sW_u32 + wslot * W_SLOT;
              if (FAMILY == FAM_DGRAD) {
                // group pack: [hi of the group's tiles][lo of the group's tiles] is one contiguous region -> ONE copy
                const DgGroup G = dg_group(s, gi);
                const uint32_t bytes = (uint32_t)G.ntl * WTILE_BYTES;
                mbar_arrive_expect_tx(w_full(wslot), bytes);
                bulk_copy_g2s(w_dst, wsrc + ((size_t)c * 9 + 2 * G.idx) * WTILE_BYTES, bytes, w_full(wslot));
              } else {
                const uint8_t* t0 = wsrc + (uint32_t)(s * q.cpt + c) * WTILE_BYTES;
                const uint8_t* t1 = t0 + (size_t)q.wchunks * WTILE_BYTES;
                mbar_arrive_expect_tx(w_full(wslot), (uint32_t)NSUB * WTILE_BYTES);
#pragma unroll
                for (int pl = 0; pl < NSPLIT; ++pl) {
                  bulk_copy_g2s(w_dst + pl * GT * W_PLANE, t0 + pl * W_PLANE, W_PLANE, w_full(wslot));
                  if (NSUB == 2) bulk_copy_g2s(w_dst + pl * GT * W_PLANE + W_PLANE, t1 + pl * W_PLANE, W_PLANE, w_full(wslot));
                }
              }
            }
          }
        }
      }
#ifdef AE_TRACE
      if (tr_prod) { tr_prod[0] = (unsigned long long)pc_wait_a; tr_prod[1] = (unsigned long long)pc_wait_w; tr_prod[2] = (unsigned long long)(clock64() - pc_t0); }
#endif
      AE_TR2(3);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(WT, 0, 0), idesc2 = make_idesc(2 * WT, 0, 0);
      uint32_t aslot = 0, around = 0, wslot_r = 0, wround = 0, ti = 0;
      bool first = true, traced = false;
      // descriptors differ only in their start-address field (bits [0,14), units of 16 bytes)
      const uint64_t desc_a0 = make_desc(smem_u32(sA), 16, SBO, ROWB), desc_w0 = make_desc(smem_u32(sW), 16, SBO, ROWB);
#ifdef AE_TRACE
      unsigned long long* tr_items = (g_trace2 && blockIdx.x == 0) ? g_trace2 + 16384 : nullptr;
      long long cy_wait_a = 0, cy_wait_w = 0, cy_issue = 0, cy_t0 = clock64(), n_items = 0;
      const unsigned long long gt0 = gtimer2();
#endif
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti, first = false) {
        const uint32_t buf = ti & 1, use = ti >> 1;
        if (use > 0) { mbar_wait(tempty(buf), (use - 1) & 1); tc_fence_after(); }
        const uint32_t acc0 = tmem_base + buf * ACC_COLS;
        uint32_t wt = 0;
        for (int s = 0; s < NS; ++s) {
          const int groups = FAMILY == FAM_DGRAD ? dg_groups(s) : 1;
          for (int c = 0; c < q.cpt; ++c) {
            AE_CY(cy_wait_a, mbar_wait(a_full(aslot), around & 1); tc_fence_after());
            if (!traced) { AE_TR2(4); traced = true; }
            const uint64_t da = desc_a0 + (uint64_t)((aslot * A_SLOT) >> 4);
            const bool fresh = s == 0 && c == 0;                      // first K contribution to every accumulator of the tile
            for (int gi = 0; gi < groups; ++gi, ++wt) {
              uint32_t wslot;
              if (resident) {
                wslot = wt;
                if (first) { AE_CY(cy_wait_w, mbar_wait(w_full(wslot), 0); tc_fence_after()); }
              } else {
                wslot = wslot_r;
                AE_CY(cy_wait_w, mbar_wait(w_full(wslot), wround & 1); tc_fence_after());
                if (++wslot_r == (uint32_t)nw) { wslot_r = 0; ++wround; }
              }
#ifdef AE_TRACE
              const long long cy_i0 = clock64();
#endif
              const uint64_t dw = desc_w0 + (uint64_t)((wslot * W_SLOT) >> 4);
              uint32_t acc = acc0, idesc = GT == 2 ? idesc2 : idesc1;
              uint32_t lo_off = (GT * W_PLANE) >> 4;                   // lo plane behind the hi planes of the group's tiles
              if (FAMILY == FAM_DGRAD) {
                const DgGroup G = dg_group(s, gi);
                acc = acc0 + G.col * WT;
                idesc = G.ntl == 2 ? idesc2 : idesc1;
                lo_off = (uint32_t)(G.ntl * W_PLANE) >> 4;
              }
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint64_t ah = da + (uint64_t)(kk * 2), bh = dw + (uint64_t)(kk * 2);
                umma_bf16(acc, ah, bh, idesc, !(fresh && kk == 0));
                if (NSPLIT == 2) {
                  umma_bf16(acc, ah, bh + (uint64_t)lo_off, idesc, 1);
                  umma_bf16(acc, ah + (uint64_t)(A_PLANE >> 4), bh, idesc, 1);
                }
              }
              if (!resident) umma_commit(w_empty(wslot));
#ifdef AE_TRACE
              cy_issue += clock64() - cy_i0;
#endif
            }
            umma_commit(a_empty(aslot));
            if (++aslot == (uint32_t)na) { aslot = 0; ++around; }
            if (FAMILY == FAM_DGRAD && c == q.cpt - 1) umma_commit(tfull(buf, s));   // output phase s is complete
#ifdef AE_TRACE
            ++n_items;
#endif
          }
        }
        if (FAMILY == FAM_FPROP) umma_commit(tfull(buf, 0));
      }
#ifdef AE_TRACE
      if (tr_items) {
        tr_items[0] = (unsigned long long)cy_wait_a; tr_items[1] = (unsigned long long)cy_wait_w; tr_items[2] = (unsigned long long)cy_issue;
        tr_items[3] = (unsigned long long)(clock64() - cy_t0); tr_items[4] = gtimer2() - gt0; tr_items[5] = (unsigned long long)n_items;
      }
#endif
      AE_TR2(5);
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9; warp w reads TMEM lanes 32*(w%4) ..) =====================
    const int ew = warp - 2, qd = warp & 3, h = ew >> 2;
    const int et = tid - 64;
    // per-channel coefficients and the statistics accumulators are epilogue-only state: loaded here, behind the CTA-wide
    // barrier, so the producer's first copies do not wait for these global loads
    for (int c = et; c < MAXN; c += 256) { sStat[0][c] = 0.f; sStat[1][c] = 0.f; }
    for (int c = et; c < q.N; c += 256) {              // q.N <= MAXN output channels
      float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q.epi.mode == AE_EPI_RELUBWD_STATS) {
        const int ch = c % q.epi.C;
        k = make_float4(__ldg(q.epi.bnc + AE_BNC_SCALE * q.epi.C + ch), __ldg(q.epi.bnc + AE_BNC_SHIFT * q.epi.C + ch),
                        __ldg(q.epi.bnc + AE_BNC_MEAN * q.epi.C + ch), __ldg(q.epi.bnc + AE_BNC_RSTD * q.epi.C + ch));
      } else if (q.epi.bias) {
        k.x = __ldg(q.epi.bias + c);
      }
      *reinterpret_cast<float4*>(&sCoef[c][0]) = k;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const Epilogue e = q.epi;
    const bool relubwd = e.mode == AE_EPI_RELUBWD_STATS;
    const bool do_stats = e.mode != AE_EPI_STORE && e.stats != nullptr;
    uint8_t* Oreg = sStage + (size_t)ew * R2_BLOCK_BYTES;
    // forward activations of the ReLU-backward epilogue: two blocks per warp (only carved for that epilogue), so that the
    // load of the warp's next unit is in flight while the current one is processed
    uint8_t* const Yreg0 = sStage + (size_t)(8 + ew) * R2_BLOCK_BYTES;
    const uint32_t o_u32 = smem_u32(Oreg);
    const int P = g.Hs * g.Ws;
    if (h < (UNITS >= 2 ? 2 : 1)) {
      // the warp's 32-row block of tile `tile`: pixel coordinates of its first row
      auto block_origin = [&](int tile, int& nt, int& x0, int& y0, int& n0, int& mrow) {
        const int mt = tile / tiles_n;
        nt = tile - mt * tiles_n;
        mrow = mt * TILE_M + qd * 32;
        x0 = mrow & (g.Ws - 1);
        y0 = (mrow & (P - 1)) >> g.lWs;
        n0 = mrow >> (g.lHs + g.lWs);
      };
      auto issue_y = [&](int tile, int u, uint32_t slot) {  // lane 0
        int nt, x0, y0, n0, mrow;
        block_origin(tile, nt, x0, y0, n0, mrow);
        const int ph = u / CB, cb = u - ph * CB;
        const uint32_t yb = ybar(ew + 8 * slot);
        mbar_arrive_expect_tx(yb, R2_BLOCK_BYTES);
        tma_load_4d(smem_u32(Yreg0) + slot * 8 * R2_BLOCK_BYTES, &q.ymap[ph], nt * PW + cb * 32, x0, y0, n0, yb);
      };
      // the warp's unit sequence: (tile, u) -> (tile, u + 2) ... -> (tile + gridDim.x, h) ...
      auto next_unit = [&](int& tile, int& u) {
        u += 2;
        if (u >= UNITS) { tile += gridDim.x; u = h; }
      };
      if (relubwd && lane == 0) {
#pragma unroll
        for (int i = 0; i < NPH; ++i) tma_prefetch_desc(&q.ymap[i]);
        int t0 = blockIdx.x, u0 = h;
        if (t0 < num_tiles) issue_y(t0, u0, 0);
        next_unit(t0, u0);
        if (q.ny == 2 && t0 < num_tiles) issue_y(t0, u0, 1);
      }
      uint32_t ycount = 0, ti = 0;
      const int swz = lane & 7;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        int nt, x0, y0, n0, mrow;
        block_origin(tile, nt, x0, y0, n0, mrow);
        const bool row_ok = mrow + lane < q.M;
        const uint32_t buf = ti & 1;
        for (int u = h; u < UNITS; u += 2) {
          const int ph = u / CB, cb = u - ph * CB;         // phases in the order they complete
          const int n = nt * PW + cb * 32;                 // first output channel of the block
          const uint32_t col = (FAMILY == FAM_DGRAD ? dg_col_of_phase(ph) * WT : 0) + cb * 32;
          const uint32_t yslot = q.ny == 2 ? (ycount & 1) : 0;
          const uint8_t* Yreg = Yreg0 + (size_t)yslot * 8 * R2_BLOCK_BYTES;
          if (relubwd) mbar_wait(ybar(ew + 8 * yslot), (q.ny == 2 ? (ycount >> 1) : ycount) & 1);
          mbar_wait(tfull(buf, ph), (ti >> 1) & 1);
          tc_fence_after();
          if (et == 0 && u == h) { if (ti == 0) AE_TR2(6); AE_TR2(7); }
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + buf * ACC_COLS + col, v);
          if (u + 2 >= UNITS) {                            // this warp's last read of the accumulator buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
          }
          if (relubwd) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 yv = *reinterpret_cast<const float4*>(Yreg + lane * 128 + ((k ^ swz) << 4));
              const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 kf = *reinterpret_cast<const float4*>(&sCoef[n + k * 4 + j][0]);
                const float z = fmaf(ya[j], kf.x, kf.y);
                v[k * 4 + j] = (row_ok && z > 0.f) ? v[k * 4 + j] : 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = row_ok ? v[j] + sCoef[n + j][0] : 0.f;
          }
          // the previous block's tensor store must have read the staging block before it is overwritten
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(Oreg + lane * 128 + ((k ^ swz) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&q.omap[ph], o_u32, n, x0, y0, n0);
            bulk_commit();
          }
          if (do_stats) {
            // lane = column: 32 conflict-free reads per array walk the block's rows
            const float4 kf = *reinterpret_cast<const float4*>(&sCoef[n + lane][0]);
            const int cc = lane >> 2, cw = (lane & 3) * 4;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int off = r * 128 + ((cc ^ (r & 7)) << 4) + cw;
              const float d = *reinterpret_cast<const float*>(Oreg + off);
              s1 += d;
              if (relubwd) {
                const float yv = *reinterpret_cast<const float*>(Yreg + off);
                s2 = fmaf(d, (yv - kf.z) * kf.w, s2);
              } else {
                s2 = fmaf(d, d, s2);
              }
            }
            atomicAdd(&sStat[0][n + lane], s1);
            atomicAdd(&sStat[1][n + lane], s2);
          }
          ++ycount;
          if (relubwd) {
            fence_proxy_async();
            __syncwarp();                                  // every lane is done with the y block
            if (lane == 0) {                               // this slot takes the unit after next (two slots) or the next one
              int tn = tile, un = u;
              next_unit(tn, un);
              if (q.ny == 2) next_unit(tn, un);
              if (tn < num_tiles) issue_y(tn, un, yslot);
            }
          }
        }
      }
      if (lane == 0) bulk_wait0();
    }
    if (et == 0) AE_TR2(8);
    if (do_stats) {                                         // one flush per CTA: fp64 atomics, one per channel
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = et; c < q.N; c += 256) {
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
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        const unsigned int done = atomicAdd(q.job_counter, 1u);
        s_last = done == gridDim.x - 1 ? 1u : 0u;
        if (s_last) *q.job_counter = 0u;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_last) {
        __threadfence();
        bn_job_run(q.job, et, 256);
      }
    }
    if (et == 0) AE_TR2(9);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tid == 0) AE_TR2(10);
}

#ifdef AE_TRACE
extern "C" int ae_debug_set_trace2(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_trace2, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

// SMs left to concurrently running collectives (set by the engine while it captures a data-parallel step)
static int g_sm_reserve = 0;
void rowgemm2_set_sm_reserve(int sms) { g_sm_reserve = sms > 0 ? sms : 0; }

bool rowgemm2_supported(const RowGemm& p) {
  if (!tma_rowgemm_supported(p)) return false;
  if (p.epi.mode != AE_EPI_STORE && p.epi.mode != AE_EPI_BIAS_STATS && p.epi.mode != AE_EPI_RELUBWD_STATS) return false;
  if (p.epi.planes != nullptr || p.out == nullptr) return false;
  if (p.N % 32 != 0 || p.N > (p.family == FAM_DGRAD ? 128 : 256)) return false;
  if (p.family == FAM_FPROP && p.N % 64 != 0) return false;
  return true;
}

bool rowgemm2_preferred(const RowGemm& p) {
  if (!rowgemm2_supported(p)) return false;
  // Measured against the first generation at batch 256 (scripts/rowgemm_bench.py, profiles/r2_rowgemm_gen1_vs_gen2.txt):
  // one CTA per SM needs ~100 tiles to win (ConvTranspose2d 256->128 at batch 256 has 64 merged tiles), and the 32-channel
  // K chunks of Conv2d(32,64) make 16 KB boxes whose per-box cost dominates.
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  if (p.family == FAM_FPROP) {
    if (p.g.Cb < 64) return false;
    const int tiles = p.N % 128 == 0 && (int64_t)m_tiles * (p.N / 128) >= 111 ? m_tiles * (p.N / 128) : m_tiles * (p.N / 64);
    return tiles >= 96;
  }
  return m_tiles * (p.N / (p.N >= 64 ? 64 : 32)) >= 96;
}

template <int FAMILY, int WT, int KC, int NSPLIT, int NSUB>
static int launch_row2(TmaRow2& q, cudaStream_t st) {
  constexpr int GT = FAMILY == FAM_DGRAD ? 2 : NSUB;
  constexpr int PW = FAMILY == FAM_DGRAD ? WT : WT * NSUB;
  constexpr size_t A_SLOT = (size_t)NSPLIT * TILE_M * KC * 2, W_SLOT = (size_t)GT * NSPLIT * WT * KC * 2;
  int dev = 0, sms = 0, smem_max = 0;
  AE_CUDA(cudaGetDevice(&dev));
  AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  AE_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  cudaFuncAttributes fa;
  AE_CUDA(cudaFuncGetAttributes(&fa, k_rowgemm2<FAMILY, WT, KC, NSPLIT, NSUB>));
  const bool relubwd = q.epi.mode == AE_EPI_RELUBWD_STATS;
  const int tiles_n = q.N / PW;
  const int wn = (FAMILY == FAM_DGRAD ? 5 : 9) * q.cpt;      // weight groups per output tile
  const size_t room = (size_t)smem_max - fa.sharedSizeBytes - 1024 - 8 * R2_BLOCK_BYTES;
  // two y blocks per epilogue warp (the next unit's load in flight) when the rings keep at least 3 + 2 slots beside them
  q.ny = relubwd ? ((wn < 3 ? wn : 3) * W_SLOT + 2 * A_SLOT + 16 * R2_BLOCK_BYTES <= room ? 2 : 1) : 0;
  const size_t stage = (size_t)(8 + 8 * q.ny) * R2_BLOCK_BYTES;
  const size_t budget = room - (size_t)8 * q.ny * R2_BLOCK_BYTES;   // both rings
  // resident weights: one n-tile whose groups fit beside at least two activation slots
  q.resident = tiles_n == 1 && wn <= R2_MAXRING && wn * W_SLOT + 2 * A_SLOT <= budget;
  q.nw = q.resident ? wn : (wn < 4 ? wn : 4);
  while (!q.resident && q.nw > 2 && q.nw * W_SLOT + 3 * A_SLOT > budget) --q.nw;
  if (!q.resident && q.nw * W_SLOT + 2 * A_SLOT <= budget && q.nw < 3 && 3 * W_SLOT + 2 * A_SLOT <= budget) q.nw = 3;
  AE_CHECK(q.nw * W_SLOT + 2 * A_SLOT <= budget, "rowgemm2: shared memory budget too small (%zu bytes for the rings)", budget);
  int na = (int)((budget - q.nw * W_SLOT) / A_SLOT);
  if (na > 6) na = 6;
  q.na = na;
  const size_t smem = q.na * A_SLOT + q.nw * W_SLOT + stage + 1024;
  static size_t attr_set[16] = {};
  if (dev >= 16 || attr_set[dev] < smem) {
    AE_CUDA(cudaFuncSetAttribute(k_rowgemm2<FAMILY, WT, KC, NSPLIT, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < 16) attr_set[dev] = smem;
  }
  const int tiles = ((q.M + TILE_M - 1) / TILE_M) * tiles_n;
  const int avail = sms - g_sm_reserve > 16 ? sms - g_sm_reserve : sms;      // one CTA per SM: a busy SM would cost a second wave
  const int ctas = tiles < avail ? tiles : avail;
  k_rowgemm2<FAMILY, WT, KC, NSPLIT, NSUB><<<ctas, R2_THREADS, smem, st>>>(q);
  AE_LAUNCH_CHECK();
  return 0;
}

// p.A.src must point to the split-bf16 planes of the A image (AE_OP_SPLIT_BF16)
int tma_rowgemm2(const RowGemm& p, const void* packed, int nsplit, cudaStream_t st) {
  AE_CHECK(rowgemm2_supported(p), "rowgemm2: unsupported shape or epilogue");
  AE_CHECK(p.A.mode == AE_OP_SPLIT_BF16, "rowgemm2: the A operand must be split-bf16 planes (ae_split_operand)");
  AE_CHECK(((uintptr_t)packed & 15) == 0, "rowgemm2: packed weights must be 16-byte aligned");
  TmaRow2 q;
  memset(&q, 0, sizeof(q));
  q.wtiles = (const uint8_t*)packed;
  q.g = p.g; q.epi = p.epi; q.M = p.M; q.N = p.N;
  if (p.tail_job) {
    AE_CHECK(p.tail_counter != nullptr && p.tail_job->stats == p.epi.stats && p.tail_job->C <= 256,
             "rowgemm2: the tail job must describe the statistics this launch accumulates");
    q.job = *p.tail_job; q.job_counter = p.tail_counter;
  }
  const Geom& g = p.g;
  pixel_box(g.Hs, g.Ws, TILE_M, &q.bx, &q.by, &q.bn);
  int bx32, by32, bn32;
  pixel_box(g.Hs, g.Ws, 32, &bx32, &by32, &bn32);
  const bool relubwd = p.epi.mode == AE_EPI_RELUBWD_STATS;
  AE_CHECK(!relubwd || p.epi.y != nullptr, "rowgemm2: the ReLU-backward epilogue needs the forward activation");
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  if (p.family == FAM_FPROP) {
    const int KC = g.Cb >= 64 ? 64 : 32;
    q.cpt = g.Cb / KC; q.wchunks = 9 * q.cpt;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        AE_TRY(encode_map(&q.amap[py * 2 + px], p.A.src, g.B, 2 * g.Hs, 2 * g.Ws, g.Cb, nsplit, py, px, 2, 2, g.Hs, g.Ws, KC,
                          q.bx, q.by, q.bn));
    AE_TRY(encode_map_f32(&q.omap[0], p.out, g.B, g.Hs, g.Ws, p.N, 0, 0, 1, 1, g.Hs, g.Ws, bx32, by32, bn32));
    if (relubwd) AE_TRY(encode_map_f32(&q.ymap[0], p.epi.y, g.B, g.Hs, g.Ws, p.N, 0, 0, 1, 1, g.Hs, g.Ws, bx32, by32, bn32));
    // 128-wide tiles (one MMA of N = 128 reads 8 KB of operands for twice the work of N = 64) when they still fill the SMs
    const char* fe = getenv("AE_B200_FORCE_NT128");          // tests: 1 = wherever the shape allows, -1 = never
    const int force = fe ? atoi(fe) : 0;
    const bool wide = KC == 64 && p.N % 128 == 0 && (force > 0 || (force == 0 && (int64_t)m_tiles * (p.N / 128) >= 111));
    if (KC == 64) {
      if (wide) return nsplit == 2 ? launch_row2<FAM_FPROP, 64, 64, 2, 2>(q, st) : launch_row2<FAM_FPROP, 64, 64, 1, 2>(q, st);
      return nsplit == 2 ? launch_row2<FAM_FPROP, 64, 64, 2, 1>(q, st) : launch_row2<FAM_FPROP, 64, 64, 1, 1>(q, st);
    }
    return nsplit == 2 ? launch_row2<FAM_FPROP, 64, 32, 2, 1>(q, st) : launch_row2<FAM_FPROP, 64, 32, 1, 1>(q, st);
  }
  q.cpt = g.Cs / 64; q.wchunks = 9 * q.cpt;
  AE_TRY(encode_map(&q.amap[0], p.A.src, g.B, g.Hs, g.Ws, g.Cs, nsplit, 0, 0, 1, 1, g.Hs, g.Ws, 64, q.bx, q.by, q.bn));
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      AE_TRY(encode_map_f32(&q.omap[py * 2 + px], p.out, g.B, 2 * g.Hs, 2 * g.Ws, p.N, py, px, 2, 2, g.Hs, g.Ws, bx32, by32, bn32));
      if (relubwd)
        AE_TRY(encode_map_f32(&q.ymap[py * 2 + px], p.epi.y, g.B, 2 * g.Hs, 2 * g.Ws, p.N, py, px, 2, 2, g.Hs, g.Ws, bx32, by32, bn32));
    }
  if (p.N >= 64) return nsplit == 2 ? launch_row2<FAM_DGRAD, 64, 64, 2, 1>(q, st) : launch_row2<FAM_DGRAD, 64, 64, 1, 1>(q, st);
  return nsplit == 2 ? launch_row2<FAM_DGRAD, 32, 64, 2, 1>(q, st) : launch_row2<FAM_DGRAD, 32, 64, 1, 1>(q, st);
}

}  // namespace ae

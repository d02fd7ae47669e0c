// fp32 CUDA-core implicit-GEMM kernels: the bring-up / cross-check backend (AE_BACKEND_SIMT) and the
// home of the small dense layers.  Same operand-transform and epilogue semantics as the tcgen05 path.
#include "pack.cuh"

namespace ae {

// ------------------------------------------------------------------------------------------------
// Row GEMM: C[m][n] = sum_k A(m,k) * Bp[k][n].  64x64 tile, 16-deep k-step, 256 threads, 4x4 / thread.
// ------------------------------------------------------------------------------------------------
template <int FAMILY>
__global__ void __launch_bounds__(256) k_rowgemm(RowGemm p) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][64];
  __shared__ float sStat[2][64];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const Geom g = p.g;
  int py = 0, px = 0, split = 0;
  int kbeg = 0, kend = p.K;
  size_t brow0 = 0;
  if (FAMILY == FAM_DGRAD) {
    const int phase = blockIdx.z;
    py = phase >> 1; px = phase & 1;
    kend = (1 + py) * (1 + px) * g.Cs;
    const int offs[4] = {0, 1, 3, 5};
    brow0 = (size_t)offs[phase] * g.Cs;
  } else {
    split = blockIdx.z;
    const int kper = p.K / p.splitK;
    kbeg = split * kper;
    kend = kbeg + kper;
  }

  // this thread's A row
  const int ar = tid >> 2, akq = (tid & 3) * 4;
  const int am = m0 + ar;
  const bool arow_ok = am < p.M;
  int an = 0, ay = 0, ax = 0;
  if (FAMILY != FAM_DENSE) {
    ax = am & (g.Ws - 1);
    ay = (am >> g.lWs) & (g.Hs - 1);
    an = am >> (g.lWs + g.lHs);
  }
  const int br = tid >> 4, bc = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    {  // A tile
      const int k = k0 + akq;
      bool valid = arow_ok;
      size_t off = 0;
      int c = 0;
      if (FAMILY == FAM_FPROP) {
        const int tap = k / g.Cb;
        c = k - tap * g.Cb;
        const int ky = tap / 3, kx = tap - ky * 3;
        const int iy = 2 * ay - 1 + ky, ix = 2 * ax - 1 + kx;
        valid = valid && iy >= 0 && ix >= 0;
        off = (((size_t)an * (2 * g.Hs) + iy) * (2 * g.Ws) + ix) * g.Cb + c;
      } else if (FAMILY == FAM_DGRAD) {
        const int t = k / g.Cs;
        c = k - t * g.Cs;
        const int a = t / (1 + px), b = t - a * (1 + px);
        const int sy = ay + ((py && a == 0) ? 1 : 0), sx = ax + ((px && b == 0) ? 1 : 0);
        valid = valid && sy < g.Hs && sx < g.Ws;
        off = (((size_t)an * g.Hs + sy) * g.Ws + sx) * g.Cs + c;
      } else {
        off = (size_t)am * p.K + k;
        c = k % p.A.C;
      }
      const float4 v = load_operand4(p.A, off, c, valid);
      As[akq + 0][ar] = v.x; As[akq + 1][ar] = v.y; As[akq + 2][ar] = v.z; As[akq + 3][ar] = v.w;
    }
    {  // B tile
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + bc < p.N) v = __ldg(reinterpret_cast<const float4*>(p.Bp + (brow0 + k0 + br) * p.N + n0 + bc));
      *reinterpret_cast<float4*>(&Bs[br][bc]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  const bool col_ok = n < p.N;
  if (p.splitK > 1) {
    if (col_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m < p.M)
          *reinterpret_cast<float4*>(p.partial + ((size_t)split * p.M + m) * p.N + n) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    return;
  }

  const Epilogue e = p.epi;
  if (tid < 64) { sStat[0][tid] = 0.f; sStat[1][tid] = 0.f; }
  __syncthreads();
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (col_ok) {
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n)); bias[0] = b.x; bias[1] = b.y; bias[2] = b.z; bias[3] = b.w; }
    const int ch = n % e.C;
    float sc[4], sh[4], mu[4], rs[4];
    if (e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[j] = __ldg(e.bnc + AE_BNC_SCALE * e.C + ch + j); sh[j] = __ldg(e.bnc + AE_BNC_SHIFT * e.C + ch + j);
        mu[j] = __ldg(e.bnc + AE_BNC_MEAN * e.C + ch + j);  rs[j] = __ldg(e.bnc + AE_BNC_RSTD * e.C + ch + j);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= p.M) continue;
      size_t row;
      if (FAMILY == FAM_DGRAD) {
        const int x = m & (g.Ws - 1), y = (m >> g.lWs) & (g.Hs - 1), nn = m >> (g.lWs + g.lHs);
        row = (((size_t)nn * (2 * g.Hs) + 2 * y + py) * (2 * g.Ws) + 2 * x + px) * p.N;
      } else {
        row = (size_t)m * p.N;
      }
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
      if (e.mode == AE_EPI_BIAS_STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[j] += v[j]; s2[j] += v[j] * v[j]; }
      } else if (e.mode == AE_EPI_RELUBWD_STATS) {
        const float4 yv4 = __ldg(reinterpret_cast<const float4*>(e.y + row + n));
        const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = fmaf(yv[j], sc[j], sh[j]);
          v[j] = z > 0.f ? v[j] : 0.f;
          s1[j] += v[j];
          s2[j] += v[j] * ((yv[j] - mu[j]) * rs[j]);
        }
      }
      if (p.out) *reinterpret_cast<float4*>(p.out + row + n) = make_float4(v[0], v[1], v[2], v[3]);
      if (e.planes) {     // the consumer is a tcgen05 GEMM: hand it its operand planes directly (no separate split pass)
        __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(e.planes);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(pl + row + n) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        if (e.nsplit == 2) {
          float lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
          __nv_bfloat162 l0 = __floats2bfloat162_rn(lo[0], lo[1]), l1 = __floats2bfloat162_rn(lo[2], lo[3]);
          *reinterpret_cast<uint2*>(pl + (size_t)p.M * p.N + row + n) =
              make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
        }
      }
    }
  }
  if (e.mode != AE_EPI_STORE && e.stats) {
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&sStat[0][tx * 4 + j], s1[j]); atomicAdd(&sStat[1][tx * 4 + j], s2[j]); }
    }
    __syncthreads();
    if (tid < 64 && n0 + tid < p.N) {
      const int ch = (n0 + tid) % e.C;
      atomicAdd(e.stats + ch, (double)sStat[0][tid]);
      atomicAdd(e.stats + e.C + ch, (double)sStat[1][tid]);
    }
  }
}

int simt_rowgemm(const RowGemm& p, cudaStream_t st) {
  AE_CHECK(p.N % 4 == 0, "simt_rowgemm: N=%d must be a multiple of 4", p.N);
  AE_CHECK(p.epi.mode != AE_EPI_BNRELU_SPLIT, "simt_rowgemm: the split-bf16 epilogue exists on the tcgen05 path only");
  AE_CHECK(p.epi.planes == nullptr || (p.family == FAM_DENSE && p.epi.mode == AE_EPI_STORE && p.splitK <= 1),
           "simt_rowgemm: operand planes can only be written by a dense GEMM with the plain store epilogue");
  AE_CHECK(p.out != nullptr || p.epi.planes != nullptr || p.splitK > 1, "simt_rowgemm: no output");
  dim3 grid((p.M + 63) / 64, (p.N + 63) / 64, 1);
  if (p.family == FAM_DGRAD) {
    AE_CHECK(p.g.Cs % 16 == 0, "simt_rowgemm: Cs=%d must be a multiple of 16", p.g.Cs);
    grid.z = 4;
    k_rowgemm<FAM_DGRAD><<<grid, 256, 0, st>>>(p);
  } else {
    const int s = p.splitK > 1 ? p.splitK : 1;
    AE_CHECK(p.K % (16 * s) == 0, "simt_rowgemm: K=%d must be a multiple of 16*splitK (%d)", p.K, s);
    grid.z = s;
    RowGemm q = p;
    q.splitK = s;
    if (p.family == FAM_FPROP) k_rowgemm<FAM_FPROP><<<grid, 256, 0, st>>>(q);
    else k_rowgemm<FAM_DENSE><<<grid, 256, 0, st>>>(q);
  }
  AE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Column GEMM (reduction over rows): C[i][j] = sum_m A(m,i) * B(m,j).  Weight gradients.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_colgemm(ColGemm p) {
  __shared__ __align__(16) float As[16][64];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64, split = blockIdx.z;
  const Geom g = p.g;
  const int mper = (((p.M + p.splitK - 1) / p.splitK) + 15) & ~15;
  const int mbeg = split * mper;
  const int mend = min(p.M, mbeg + mper);
  const int r = tid >> 4, q = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;

  // column decode of this thread's A chunk (fixed across the loop)
  const int ai = i0 + q;
  const bool ai_ok = ai < p.I;
  int tap = 0, acb = 0, ky = 0, kx = 0;
  if (p.gather) { tap = ai / g.Cb; acb = ai - tap * g.Cb; ky = tap / 3; kx = tap - ky * 3; }
  const int bj = j0 + q;
  const bool bj_ok = bj < p.J;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int mb = mbeg; mb < mend; mb += 16) {
    const int m = mb + r;
    const bool m_ok = m < mend;
    {
      bool valid = m_ok && ai_ok;
      size_t off;
      int c;
      if (p.gather) {
        const int ox = m & (g.Ws - 1), oy = (m >> g.lWs) & (g.Hs - 1), n = m >> (g.lWs + g.lHs);
        const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
        valid = valid && iy >= 0 && ix >= 0;
        off = (((size_t)n * (2 * g.Hs) + iy) * (2 * g.Ws) + ix) * g.Cb + acb;
        c = acb;
      } else {
        off = (size_t)m * p.I + ai;
        c = ai % p.A.C;
      }
      *reinterpret_cast<float4*>(&As[r][q]) = load_operand4(p.A, off, c, valid);
    }
    {
      const size_t off = (size_t)m * p.J + bj;
      *reinterpret_cast<float4*>(&Bs[r][q]) = load_operand4(p.B, off, bj % p.B.C, m_ok && bj_ok);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* dst = p.splitK > 1 ? p.partial + (size_t)split * p.I * p.J : p.out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= p.I) continue;
    const int ip = p.permC > 0 ? (ii % p.permC) * p.permHW + ii / p.permC : ii;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj >= p.J) continue;
      const size_t idx = p.transposed ? (size_t)jj * p.I + ip : (size_t)ip * p.J + jj;
      dst[idx] = acc[i][j];
    }
  }
}

int colgemm_default_split(int M, int I, int J) {
  const int tiles = ((I + 63) / 64) * ((J + 63) / 64);
  int s = (2 * 148 + tiles - 1) / tiles;
  const int max_by_rows = (M + 63) / 64;
  if (s > max_by_rows) s = max_by_rows;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

int simt_colgemm(const ColGemm& p, cudaStream_t st) {
  AE_CHECK(p.I % 4 == 0 && p.J % 4 == 0, "simt_colgemm: I=%d, J=%d must be multiples of 4", p.I, p.J);
  AE_CHECK(p.splitK >= 1, "simt_colgemm: splitK must be >= 1");
  AE_CHECK(p.splitK == 1 || p.partial != nullptr, "simt_colgemm: split-K needs a partial buffer");
  dim3 grid((p.I + 63) / 64, (p.J + 63) / 64, p.splitK);
  k_colgemm<<<grid, 256, 0, st>>>(p);
  AE_LAUNCH_CHECK();
  if (p.splitK > 1) return reduce_partials(p.partial, p.splitK, (int64_t)p.I * p.J, nullptr, 0, nullptr, p.out, st);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Fixed-order reduction of split partials (deterministic), optional bias / addend.
// ------------------------------------------------------------------------------------------------
__global__ void k_reduce_partials(const float* __restrict__ partial, int splits, int64_t n,
                                  const float* __restrict__ bias, int bias_n,
                                  const float* __restrict__ addend, float* __restrict__ out) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = __ldg(reinterpret_cast<const float4*>(partial) + i);
    for (int k = 1; k < splits; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (size_t)k * n) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (bias) {
      const int b = (int)((i * 4) % bias_n);
      s.x += __ldg(bias + b); s.y += __ldg(bias + b + 1); s.z += __ldg(bias + b + 2); s.w += __ldg(bias + b + 3);
    }
    if (addend) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(addend) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
  }
}

// Many splits, few outputs (the latent-sized results of the two 4096-wide dense layers): a block owns 32 float4 outputs,
// its 8 warps each sum every 8th split with four loads in flight, warp 0 adds the 8 shares in a fixed order.
__global__ void __launch_bounds__(256) k_reduce_partials_wide(const float* __restrict__ partial, int splits, int64_t n,
                                                              const float* __restrict__ bias, int bias_n,
                                                              const float* __restrict__ addend, float* __restrict__ out) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t n4 = n >> 2, i = (int64_t)blockIdx.x * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    const float4* src = reinterpret_cast<const float4*>(partial) + i;
    int k = wq;
    for (; k + 24 < splits; k += 32) {
      const float4 v0 = __ldg(src + (size_t)k * n4), v1 = __ldg(src + (size_t)(k + 8) * n4);
      const float4 v2 = __ldg(src + (size_t)(k + 16) * n4), v3 = __ldg(src + (size_t)(k + 24) * n4);
      s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; k < splits; k += 8) {
      const float4 v = __ldg(src + (size_t)k * n4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[wq][lane] = s;
  __syncthreads();
  if (wq == 0 && i < n4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) { const float4 v = red[g][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    if (bias) {
      const int b = (int)((i * 4) % bias_n);
      t.x += __ldg(bias + b); t.y += __ldg(bias + b + 1); t.z += __ldg(bias + b + 2); t.w += __ldg(bias + b + 3);
    }
    if (addend) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(addend) + i);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = t;
  }
}

int reduce_partials(const float* partial, int splits, int64_t n, const float* bias, int bias_n,
                    const float* addend, float* out, cudaStream_t st) {
  AE_CHECK(n % 4 == 0 && (bias_n == 0 || bias_n % 4 == 0), "reduce_partials: sizes must be multiples of 4");
  if (splits >= 16 && n / 4 <= 32 * 148 * 16) {
    k_reduce_partials_wide<<<(int)((n / 4 + 31) / 32), 256, 0, st>>>(partial, splits, n, bias, bias_n, addend, out);
    AE_LAUNCH_CHECK();
    return 0;
  }
  const int threads = 256;
  int64_t blocks = (n / 4 + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_reduce_partials<<<(int)blocks, threads, 0, st>>>(partial, splits, n, bias, bias_n, addend, out);
  AE_LAUNCH_CHECK();
  return 0;
}

// db[perm(n)] = sum_m a[m][n]; one warp per 32 columns, rows strided over the block's warps
__global__ void __launch_bounds__(256) k_column_sums(const float* __restrict__ a, int M, int N, int permC,
                                                     int permHW, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N)
    for (int m = w; m < M; m += 8) s += __ldg(a + (size_t)m * N + n);
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    const int np = permC > 0 ? (n % permC) * permHW + n / permC : n;
    out[np] = t;
  }
}

int column_sums(const float* a, int M, int N, int permC, int permHW, float* out, cudaStream_t st) {
  k_column_sums<<<(N + 31) / 32, 256, 0, st>>>(a, M, N, permC, permHW, out);
  AE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 weight packs
// ------------------------------------------------------------------------------------------------
// w [Cs][Cb][3][3] -> fwd [(ky,kx,cb)][cs]; dgrad: 4 phase blocks, rows (tap,cs), cols cb
__global__ void k_pack_conv_simt(const float* __restrict__ w, int Cs, int Cb, float* __restrict__ fwd,
                                 float* __restrict__ dgrad) {
  const int total = Cs * Cb * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9;
    const int cb = (idx / 9) % Cb;
    const int cs = idx / (9 * Cb);
    const float v = w[idx];
    if (fwd) fwd[((size_t)tap * Cb + cb) * Cs + cs] = v;
    if (dgrad) {
      const int ky = tap / 3, kx = tap % 3;
      const int py = (ky == 1) ? 0 : 1, px = (kx == 1) ? 0 : 1;
      const int a = (ky == 2) ? 1 : 0, b = (kx == 2) ? 1 : 0;  // index inside the phase's tap list
      const int phase = py * 2 + px;
      const int offs[4] = {0, 1, 3, 5};
      const int t = a * (1 + px) + b;
      dgrad[(((size_t)offs[phase] + t) * Cs + cs) * Cb + cb] = v;
    }
  }
}

int pack_conv_simt(const float* w, int Cs, int Cb, float* fwd, float* dgrad, cudaStream_t st) {
  const int total = Cs * Cb * 9;
  k_pack_conv_simt<<<(total + 255) / 256, 256, 0, st>>>(w, Cs, Cb, fwd, dgrad);
  AE_LAUNCH_CHECK();
  return 0;
}

__global__ void k_pack_linear(const float* __restrict__ w, int N, int K, int permC, int permHW, int kind,
                              float* __restrict__ dst) {
  const int total = N * K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    pack_linear_elem(idx, w, N, K, permC, permHW, kind, dst);
}

int pack_linear(const float* w, int N, int K, int permC, int permHW, int kind, float* dst, cudaStream_t st) {
  const int total = N * K;
  k_pack_linear<<<(total + 255) / 256, 256, 0, st>>>(w, N, K, permC, permHW, kind, dst);
  AE_LAUNCH_CHECK();
  return 0;
}

__global__ void k_permute_vector(const float* __restrict__ src, int n, int permC, int permHW, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) permute_elem(i, src, permC, permHW, dst);
}

int permute_vector(const float* src, int n, int permC, int permHW, float* dst, cudaStream_t st) {
  k_permute_vector<<<(n + 255) / 256, 256, 0, st>>>(src, n, permC, permHW, dst);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

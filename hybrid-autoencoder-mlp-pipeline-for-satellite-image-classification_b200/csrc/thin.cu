// The two 3-channel layers: Conv2d(3,32,k3,s2,p1) (NB:504) and ConvTranspose2d(32,3,k3,s2,p1,op1)+Sigmoid
// (NB:628-629).  K=27 / N=3 do not fill a tensor-core tile; these are bandwidth kernels on CUDA cores.
//   thin = [B,3,64,64] NCHW fp32 (the reference's image layout), wide = [B,32,32,32] NHWC fp32.
//   both layers store their weight as [32][3][3][3] = [c32][c3][ky][kx].
#include "common.cuh"

namespace ae {

static constexpr int TH = 64, TW = 64, WH = 32, WW = 32, WC = 32;

__device__ __forceinline__ float thin_value(const Operand& op, size_t idx) {
  const float a = __ldg(op.src + idx);
  if (op.mode == AE_OP_RAW) return a;
  const float s = __ldg(op.src2 + idx);                       // AE_OP_SIGMOID_BWD
  const float up = (op.scalar != 0.f) ? op.scalar * (s - a) : a;  // fused MSE gradient, or a given upstream gradient
  return up * s * (1.f - s);
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

// ---------------------------------------------------------------------------------------------
// thin -> wide gather (conv1 forward; convT4 data gradient).  One thread per wide pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_thin_gather(Operand thin, const float* __restrict__ w, Epilogue e,
                                                     float* __restrict__ out, int batch) {
  __shared__ __align__(16) float Wsm[27][32];
  __shared__ float sStat[2][32];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 27 * 32; i += 128) {
    const int c32 = i / 27, k = i - c32 * 27;
    Wsm[k][c32] = __ldg(w + i);
  }
  if (tid < 32) { sStat[0][tid] = 0.f; sStat[1][tid] = 0.f; }
  __syncthreads();

  const int m = blockIdx.x * 128 + tid;  // grid covers batch*1024 exactly
  const int ox = m & 31, oy = (m >> 5) & 31, n = m >> 10;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
  for (int c3 = 0; c3 < 3; ++c3) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        float v = 0.f;
        if (iy >= 0 && ix >= 0) v = thin_value(thin, (((size_t)n * 3 + c3) * TH + iy) * TW + ix);
        const int k = c3 * 9 + ky * 3 + kx;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 wv = *reinterpret_cast<const float4*>(&Wsm[k][c4 * 4]);
          acc[c4 * 4 + 0] = fmaf(v, wv.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(v, wv.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(v, wv.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(v, wv.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  const size_t row = (size_t)m * WC;
  float s2v[32];
  if (e.mode == AE_EPI_RELUBWD_STATS) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 y4 = __ldg(reinterpret_cast<const float4*>(e.y + row) + c4);
      const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c4 * 4 + j;
        const float z = fmaf(yv[j], __ldg(e.bnc + AE_BNC_SCALE * WC + c), __ldg(e.bnc + AE_BNC_SHIFT * WC + c));
        acc[c] = z > 0.f ? acc[c] : 0.f;
        s2v[c] = acc[c] * ((yv[j] - __ldg(e.bnc + AE_BNC_MEAN * WC + c)) * __ldg(e.bnc + AE_BNC_RSTD * WC + c));
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (e.bias) acc[c] += __ldg(e.bias + c);
      s2v[c] = acc[c] * acc[c];
    }
  }
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    reinterpret_cast<float4*>(out + row)[c4] = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
  if (e.mode != AE_EPI_STORE && e.stats) {
    const float a = warp_colsum32(acc, lane);
    const float b = warp_colsum32(s2v, lane);
    atomicAdd(&sStat[0][lane], a);
    atomicAdd(&sStat[1][lane], b);
    __syncthreads();
    if (tid < 32) {
      atomicAdd(e.stats + tid, (double)sStat[0][tid]);
      atomicAdd(e.stats + WC + tid, (double)sStat[1][tid]);
    }
  }
}

int thin_gather_fwd(const Operand& thin, const float* w, const Epilogue& epi, float* out, int batch, cudaStream_t st) {
  k_thin_gather<<<batch * 8, 128, 0, st>>>(thin, w, epi, out, batch);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// wide -> thin scatter + sigmoid (+ squared error): convT4 forward.  One thread per wide pixel,
// producing the 2x2x3 output quad it alone owns (no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_thin_scatter_sigmoid(Operand wide, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ x_hat,
                                                              const float* __restrict__ x, double* __restrict__ sse,
                                                              int batch) {
  __shared__ __align__(16) float Wsm[9][3][32];  // [tap][co][ci]
  __shared__ float red[4];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 27 * 32; i += 128) {
    const int ci = i / 27, r = i - ci * 27, co = r / 9, tap = r - co * 9;
    Wsm[tap][co][ci] = __ldg(w + i);
  }
  __syncthreads();
  const int m = blockIdx.x * 128 + tid;
  const int ix = m & 31, iy = (m >> 5) & 31, n = m >> 10;

  float acc[2][2][3];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[a][b][c] = 0.f;

  // source pixel (iy+dy, ix+dx) contributes to output parity (py,px) through tap (ky,kx):
  //   dy=0: py=0 -> ky=1 ; py=1 -> ky=2        dy=1: py=1 -> ky=0
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int sy = iy + dy, sx = ix + dx;
      const bool valid = sy < WH && sx < WW;
      const size_t off = (((size_t)n * WH + sy) * WW + sx) * WC;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 v4 = load_operand4(wide, off + c4 * 4, c4 * 4, valid);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int py = dy; py < 2; ++py) {
          const int ky = dy ? 0 : (py ? 2 : 1);
#pragma unroll
          for (int px = dx; px < 2; ++px) {
            const int kx = dx ? 0 : (px ? 2 : 1);
#pragma unroll
            for (int co = 0; co < 3; ++co) {
              const float4 wv = *reinterpret_cast<const float4*>(&Wsm[ky * 3 + kx][co][c4 * 4]);
              acc[py][px][co] = fmaf(v[0], wv.x, fmaf(v[1], wv.y, fmaf(v[2], wv.z, fmaf(v[3], wv.w, acc[py][px][co]))));
            }
          }
        }
      }
    }
  }
  float err = 0.f;
#pragma unroll
  for (int co = 0; co < 3; ++co) {
    const float b = __ldg(bias + co);
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const size_t o = (((size_t)n * 3 + co) * TH + 2 * iy + py) * TW + 2 * ix;
      const float s0 = 1.f / (1.f + expf(-(acc[py][0][co] + b)));
      const float s1 = 1.f / (1.f + expf(-(acc[py][1][co] + b)));
      *reinterpret_cast<float2*>(x_hat + o) = make_float2(s0, s1);
      if (x) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(x + o));
        err += (s0 - t.x) * (s0 - t.x) + (s1 - t.y) * (s1 - t.y);
      }
    }
  }
  if (x && sse) {
    err = warp_sum(err);
    if (lane == 0) red[tid >> 5] = err;
    __syncthreads();
    if (tid == 0) atomicAdd(sse, (double)red[0] + (double)red[1] + (double)red[2] + (double)red[3]);
  }
}

int thin_scatter_sigmoid_fwd(const Operand& wide, const float* w, const float* bias, float* x_hat, const float* x,
                             double* sse, int batch, cudaStream_t st) {
  k_thin_scatter_sigmoid<<<batch * 8, 128, 0, st>>>(wide, w, bias, x_hat, x, sse, batch);
  AE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of either thin layer: dw[c32][k] = sum_pixels wide(p,c32) * patch(p,k), k=(c3,ky,kx).
// Persistent blocks of 288 threads; per 128-pixel tile the operands are staged in shared memory,
// each thread owns a 4(c32) x 3(k) register tile for one quarter of the tile's pixels; the four
// quarters are summed through shared memory, every block writes one partial, a fixed-order
// reduction kernel sums the partials (deterministic).
// ---------------------------------------------------------------------------------------------
static constexpr int TW_THREADS = 288;
static constexpr int TW_PART = 868;  // 864 weights + 3 thin-bias sums + 1 pad

__global__ void __launch_bounds__(TW_THREADS) k_thin_wgrad(Operand wide, Operand thin, float* __restrict__ partial,
                                                           int batch) {
  __shared__ __align__(16) float Ws[128][36];
  __shared__ float Ts[128][28];
  __shared__ float red[4][864];
  __shared__ float bsum[TW_THREADS / 32][3];
  const int tid = threadIdx.x;
  const int grp = tid / 72, u = tid - grp * 72;
  const int c4 = u & 7, kg = u >> 3;
  float acc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;
  float bs[3] = {0.f, 0.f, 0.f};

  const int tiles = batch * 8;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int m0 = tile * 128;
    for (int i = tid; i < 128 * 8; i += TW_THREADS) {
      const int p = i >> 3, q = i & 7;
      *reinterpret_cast<float4*>(&Ws[p][q * 4]) = load_operand4(wide, (size_t)(m0 + p) * WC + q * 4, q * 4, true);
    }
    for (int i = tid; i < 128 * 27; i += TW_THREADS) {
      const int p = i / 27, k = i - p * 27;
      const int m = m0 + p;
      const int ox = m & 31, oy = (m >> 5) & 31, n = m >> 10;
      const int c3 = k / 9, r = k - c3 * 9, ky = r / 3, kx = r - ky * 3;
      const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
      float v = 0.f;
      if (iy >= 0 && ix >= 0) v = thin_value(thin, (((size_t)n * 3 + c3) * TH + iy) * TW + ix);
      Ts[p][k] = v;
      // taps (ky,kx) in {1,2}^2 enumerate the 2x2 big-image quad owned by this small pixel exactly once
      if (ky >= 1 && kx >= 1) {
        if (c3 == 0) bs[0] += v; else if (c3 == 1) bs[1] += v; else bs[2] += v;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < 32; ++pp) {
      const int p = grp * 32 + pp;
      const float4 a = *reinterpret_cast<const float4*>(&Ws[p][c4 * 4]);
      const float b0 = Ts[p][kg * 3 + 0], b1 = Ts[p][kg * 3 + 1], b2 = Ts[p][kg * 3 + 2];
      acc[0][0] = fmaf(a.x, b0, acc[0][0]); acc[0][1] = fmaf(a.x, b1, acc[0][1]); acc[0][2] = fmaf(a.x, b2, acc[0][2]);
      acc[1][0] = fmaf(a.y, b0, acc[1][0]); acc[1][1] = fmaf(a.y, b1, acc[1][1]); acc[1][2] = fmaf(a.y, b2, acc[1][2]);
      acc[2][0] = fmaf(a.z, b0, acc[2][0]); acc[2][1] = fmaf(a.z, b1, acc[2][1]); acc[2][2] = fmaf(a.z, b2, acc[2][2]);
      acc[3][0] = fmaf(a.w, b0, acc[3][0]); acc[3][1] = fmaf(a.w, b1, acc[3][1]); acc[3][2] = fmaf(a.w, b2, acc[3][2]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) red[grp][(c4 * 4 + i) * 27 + kg * 3 + j] = acc[i][j];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = warp_sum(bs[c]);
    if ((tid & 31) == 0) bsum[tid >> 5][c] = s;
  }
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * TW_PART;
  for (int i = tid; i < 864; i += TW_THREADS) dst[i] = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
  if (tid < 3) {
    float s = 0.f;
    for (int wq = 0; wq < TW_THREADS / 32; ++wq) s += bsum[wq][tid];
    dst[864 + tid] = s;
  }
  if (tid == 3) dst[867] = 0.f;
}

__global__ void k_thin_wgrad_reduce(const float* __restrict__ partial, int nparts, float* __restrict__ dw,
                                    float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 867) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += __ldg(partial + (size_t)p * TW_PART + i);
  if (i < 864) dw[i] = s;
  else if (dbias) dbias[i - 864] = s;
}

static int thin_wgrad_blocks(int batch) {
  const int tiles = batch * 8;
  return tiles < 296 ? tiles : 296;
}
size_t thin_wgrad_workspace_bytes(int batch) { return (size_t)thin_wgrad_blocks(batch) * TW_PART * sizeof(float); }

int thin_wgrad(const Operand& wide, const Operand& thin, float* dw, float* dbias, void* partials, size_t bytes,
               int batch, cudaStream_t st) {
  const int blocks = thin_wgrad_blocks(batch);
  AE_CHECK(bytes >= (size_t)blocks * TW_PART * sizeof(float), "thin_wgrad: workspace too small");
  k_thin_wgrad<<<blocks, TW_THREADS, 0, st>>>(wide, thin, static_cast<float*>(partials), batch);
  AE_LAUNCH_CHECK();
  k_thin_wgrad_reduce<<<(867 + 127) / 128, 128, 0, st>>>(static_cast<const float*>(partials), blocks, dw, dbias);
  AE_LAUNCH_CHECK();
  return 0;
}

}  // namespace ae

// The latent classifier MLP (NB:2970-2987): Linear(D,128) BN1d ReLU Dropout(0.3) Linear(128,64) BN1d ReLU
// Linear(64,C), with softmax cross-entropy (NB:3463), forward AND backward in ONE persistent kernel.
//
// One thread-block cluster of 8 CTAs owns the whole batch: CTA r handles a contiguous slice of rows, all
// three weight matrices stay resident in its shared memory, BatchNorm1d batch statistics, the loss and
// the weight-gradient partials are combined across the cluster through distributed shared memory in a
// fixed order (deterministic).  The only global traffic is x, the row-local intermediates and the results.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ae {

static constexpr int H1 = 128, H2 = 64, CP = 16, CHUNK = 32, NT = 256;   // cluster size: template parameter of k_mlp
static constexpr int LD1 = H1 + 1, LD2 = H2 + 1, LD3 = CP + 1;
static constexpr float BN_EPS_F = 1e-5f;

enum { MLP_FWD = 1, MLP_CE = 2, MLP_BWD = 4, MLP_TRAIN = 8 };

struct MlpArgs {
  const float* params;
  float* grads;
  float* running;           // [mean1(128), var1(128), mean2(64), var2(64)]
  const float* x;
  const int64_t* labels;
  const uint8_t* keep_in;   // optional explicit dropout keep mask [B][128]
  const float* dlogits_in;  // backward-only mode
  float* logits;
  float* loss;
  int* correct;
  // workspace (row-local intermediates, saved between a forward-only and a backward-only call)
  float* h1; float* h2; float* d1; float* d2; float* dlog; uint8_t* keep; float* bnc;  // bnc: [2][4][128]
  int B, D, C, flags;
  unsigned long long seed;
  // graph-replayable step (ae_mlp_train_step): the dropout seed of launch i is seed + seed_dev[0], and the launch advances
  // seed_dev[0] and both BatchNorm layers' num_batches_tracked itself
  unsigned long long* seed_dev;
  // epoch mode (ae_mlp_train_step_indexed): row r of the batch is row order[cursor[0] + r] of x / labels; the launch records
  // (loss, correct) in hist[cursor[1]] and advances cursor by (B rows, 1 step) -- a whole epoch is then graph replays only
  const long long* order;
  long long* cursor;
  float* hist;
  int64_t* nbt;             // [2] or NULL
  float p;
  int64_t off[10];
};

__device__ __forceinline__ unsigned hash_u32(unsigned long long seed, unsigned r, unsigned c) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)r * 131u + c + 1u);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (unsigned)(z >> 32);
}

// Sum `n` per-CTA fp32 partials (S slots each) over the cluster in a fixed order -> tot[n] (double).
template <int NCTA>
__device__ void cluster_sum(cg::cluster_group& cl, float* part, int slots, int n, double* tot) {
  cl.sync();
  for (int i = threadIdx.x; i < n; i += NT) {
    double s = 0.0;
    for (unsigned r = 0; r < NCTA; ++r) {
      const float* rp = cl.map_shared_rank(part, r);
      for (int k = 0; k < slots; ++k) s += (double)rp[k * n + i];
    }
    tot[i] = s;
  }
  cl.sync();
}

// out[r][j] = bias[j] + sum_k inT[k][r] * Ws[k][j] for the 32 staged rows; NOUT in {128, 64}.
// Also accumulates per-column sum / sum of squares over the valid rows into st (slot = row group).
template <int NOUT>
__device__ __forceinline__ void tile_linear(const float* __restrict__ inT, int K, const float* __restrict__ Ws, int ld,
                                            const float* __restrict__ bias, float* __restrict__ out, int row0, int nvalid,
                                            float* s1, float* s2) {
  constexpr int G = NT / NOUT;         // row groups
  constexpr int RPT = CHUNK / G;       // rows per thread
  const int j = threadIdx.x % NOUT, rg = threadIdx.x / NOUT;
  float acc[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float w = Ws[k * ld + j];
#pragma unroll
    for (int q = 0; q < RPT / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(inT + k * CHUNK + rg * RPT + q * 4);
      acc[q * 4 + 0] = fmaf(v.x, w, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(v.y, w, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(v.z, w, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(v.w, w, acc[q * 4 + 3]);
    }
  }
  const float b = bias[j];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg * RPT + i;
    if (r < nvalid) {
      const float v = acc[i] + b;
      out[(size_t)(row0 + r) * NOUT + j] = v;
      *s1 += v;
      *s2 += v * v;
    }
  }
}

// Stage a torch Linear weight W [J][K] (row-major, K a multiple of 4) transposed into shared memory Ws [K][LD]: all of a
// thread's 16-byte global loads are issued before the first shared-memory store (a plain element loop exposes one global
// latency per iteration: 64 iterations for the two hidden layers, most of the eval kernel's time).
template <int NTHREADS = NT>
__device__ __forceinline__ void stage_weight_T(const float* __restrict__ W, float* __restrict__ Ws, int J, int K, int LD, int tid) {
  const int total4 = J * K / 4;
  for (int base = 0; base < total4; base += 8 * NTHREADS) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i4 = base + u * NTHREADS + tid;
      v[u] = i4 < total4 ? __ldg(reinterpret_cast<const float4*>(W) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i4 = base + u * NTHREADS + tid;
      if (i4 < total4) {
        const int i = i4 * 4, j = i / K, k = i - j * K;
        Ws[k * LD + j] = v[u].x; Ws[(k + 1) * LD + j] = v[u].y; Ws[(k + 2) * LD + j] = v[u].z; Ws[(k + 3) * LD + j] = v[u].w;
      }
    }
  }
}

// NCTA = CTAs of the one cluster that owns the batch (launch attribute): 8 for large batches; small training batches
// (the reference trains the MLP at batch 64, NB:3443) are bound by the cluster-wide exchanges, not by arithmetic.
template <int NCTA>
__global__ void __launch_bounds__(NT, 1) k_mlp(MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  const unsigned rank = cl.block_rank();
  const int D = a.D, C = a.C, B = a.B;
  const bool training = (a.flags & MLP_TRAIN) != 0;

  float* W1s = sm;                         // [D][LD1]
  float* W2s = W1s + D * LD1;              // [128][LD2]
  float* W3s = W2s + H1 * LD2;             // [64][LD3]
  float* stg0 = W3s + H2 * LD3;            // [128][32]
  float* stg1 = stg0 + H1 * CHUNK;         // [128][32]
  float* part = stg1 + H1 * CHUNK;         // [4 slots][2][128]
  double* tot = reinterpret_cast<double*>(part + 4 * 2 * H1);  // [2][128]
  float* coef = reinterpret_cast<float*>(tot + 2 * H1);        // [2 layers][7][128]
  float* dls = coef + 2 * 7 * H1;          // [32][LD3]
  float* gW = dls + CHUNK * LD3;           // [8192] gradient exchange
  float* misc = gW + 8192;                 // [64]

  const float* W1 = a.params + a.off[0]; const float* b1 = a.params + a.off[1];
  const float* g1 = a.params + a.off[2]; const float* be1 = a.params + a.off[3];
  const float* W2 = a.params + a.off[4]; const float* b2 = a.params + a.off[5];
  const float* g2 = a.params + a.off[6]; const float* be2 = a.params + a.off[7];
  const float* W3 = a.params + a.off[8]; const float* b3 = a.params + a.off[9];

  // resident weights, k-major with an odd leading dimension (conflict-free for both access directions)
  stage_weight_T(W1, W1s, H1, D, LD1, tid);
  stage_weight_T(W2, W2s, H2, H1, LD2, tid);
  for (int i = tid; i < H2 * LD3; i += NT) W3s[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < C * H2; i += NT) { const int c = i / H2, k = i - c * H2; W3s[k * LD3 + c] = W3[i]; }

  const int rows_per = (B + NCTA - 1) / NCTA;
  const int r0 = min(B, (int)rank * rows_per), r1 = min(B, r0 + rows_per);
  const float keep_scale = 1.f / (1.f - a.p);
  const unsigned long long seed = a.seed + (a.seed_dev ? a.seed_dev[0] : 0ull);
  const long long row_base = a.order ? a.cursor[0] : 0;
  auto src_row = [&](int r) -> size_t { return a.order ? (size_t)a.order[row_base + r] : (size_t)r; };
  float* c1 = coef;            // layer 1: scale, shift, mean, rstd, A, B, C  (rows of 128)
  float* c2 = coef + 7 * H1;   // layer 2
  __syncthreads();

  // =============================== forward ===============================
  if (a.flags & MLP_FWD) {
    // ---- layer 1: h1 = x W1^T + b1, batch statistics ----
    {
      float s1 = 0.f, s2 = 0.f;
      for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int nv = min(CHUNK, r1 - c0);
        for (int i = tid; i < D * CHUNK; i += NT) {
          const int r = i / D, k = i - r * D;
          stg0[k * CHUNK + r] = r < nv ? a.x[src_row(c0 + r) * D + k] : 0.f;
        }
        __syncthreads();
        tile_linear<H1>(stg0, D, W1s, LD1, b1, a.h1, c0, nv, &s1, &s2);
        __syncthreads();
      }
      const int j = tid % H1, rg = tid / H1;
      part[(rg * 2 + 0) * H1 + j] = s1;
      part[(rg * 2 + 1) * H1 + j] = s2;
    }
    if (training) {
      // slots: 2 row groups x {sum, sumsq} x 128 -> view as 2 slots of n = 256
      cluster_sum<NCTA>(cl, part, 2, 2 * H1, tot);
    } else {
      __syncthreads();
    }
    for (int j = tid; j < H1; j += NT) {
      float mean, var;
      if (training) {
        const double m = tot[j] / B;
        double v = tot[H1 + j] / B - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        if (rank == 0 && a.running) {
          const double unb = B > 1 ? v * B / (B - 1.0) : v;
          a.running[j] = (float)(0.9 * (double)a.running[j] + 0.1 * m);
          a.running[H1 + j] = (float)(0.9 * (double)a.running[H1 + j] + 0.1 * unb);
        }
      } else {
        mean = a.running[j]; var = a.running[H1 + j];
      }
      const float rstd = 1.f / sqrtf(var + BN_EPS_F);
      const float sc = g1[j] * rstd;
      c1[0 * H1 + j] = sc; c1[1 * H1 + j] = be1[j] - mean * sc; c1[2 * H1 + j] = mean; c1[3 * H1 + j] = rstd;
      if (rank == 0 && a.bnc) { a.bnc[0 * H1 + j] = sc; a.bnc[1 * H1 + j] = be1[j] - mean * sc; a.bnc[2 * H1 + j] = mean; a.bnc[3 * H1 + j] = rstd; }
    }
    __syncthreads();
    // ---- layer 2: a1 = dropout(relu(bn1(h1))); h2 = a1 W2^T + b2 ----
    {
      float s1 = 0.f, s2 = 0.f;
      for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int nv = min(CHUNK, r1 - c0);
        for (int i = tid; i < H1 * CHUNK; i += NT) {
          const int r = i / H1, k = i - r * H1;
          float v = 0.f;
          if (r < nv) {
            const size_t g = (size_t)(c0 + r) * H1 + k;
            v = fmaxf(fmaf(a.h1[g], c1[k], c1[H1 + k]), 0.f);
            if (training && a.p > 0.f) {
              unsigned char kp;
              if (a.keep_in) kp = a.keep_in[g];
              else kp = (hash_u32(seed, c0 + r, k) * (1.0f / 4294967296.0f)) >= a.p ? 1 : 0;
              if (a.keep) a.keep[g] = kp;
              v = kp ? v * keep_scale : 0.f;
            }
          }
          stg0[k * CHUNK + r] = v;
        }
        __syncthreads();
        tile_linear<H2>(stg0, H1, W2s, LD2, b2, a.h2, c0, nv, &s1, &s2);
        __syncthreads();
      }
      const int j = tid % H2, rg = tid / H2;
      part[(rg * 2 + 0) * H2 + j] = s1;
      part[(rg * 2 + 1) * H2 + j] = s2;
    }
    if (training) cluster_sum<NCTA>(cl, part, 4, 2 * H2, tot);
    else __syncthreads();
    for (int j = tid; j < H2; j += NT) {
      float mean, var;
      if (training) {
        const double m = tot[j] / B;
        double v = tot[H2 + j] / B - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        if (rank == 0 && a.running) {
          const double unb = B > 1 ? v * B / (B - 1.0) : v;
          a.running[2 * H1 + j] = (float)(0.9 * (double)a.running[2 * H1 + j] + 0.1 * m);
          a.running[2 * H1 + H2 + j] = (float)(0.9 * (double)a.running[2 * H1 + H2 + j] + 0.1 * unb);
        }
      } else {
        mean = a.running[2 * H1 + j]; var = a.running[2 * H1 + H2 + j];
      }
      const float rstd = 1.f / sqrtf(var + BN_EPS_F);
      const float sc = g2[j] * rstd;
      c2[0 * H1 + j] = sc; c2[1 * H1 + j] = be2[j] - mean * sc; c2[2 * H1 + j] = mean; c2[3 * H1 + j] = rstd;
      if (rank == 0 && a.bnc) { float* o = a.bnc + 4 * H1; o[0 * H1 + j] = sc; o[1 * H1 + j] = be2[j] - mean * sc; o[2 * H1 + j] = mean; o[3 * H1 + j] = rstd; }
    }
    __syncthreads();
    // ---- layer 3: logits = relu(bn2(h2)) W3^T + b3 ; cross-entropy ----
    float lsum = 0.f;
    int ok = 0;
    for (int c0 = r0; c0 < r1; c0 += CHUNK) {
      const int nv = min(CHUNK, r1 - c0);
      for (int i = tid; i < H2 * CHUNK; i += NT) {
        const int r = i / H2, k = i - r * H2;
        stg1[k * CHUNK + r] = r < nv ? fmaxf(fmaf(a.h2[(size_t)(c0 + r) * H2 + k], c2[k], c2[H1 + k]), 0.f) : 0.f;
      }
      __syncthreads();
      for (int i = tid; i < CHUNK * C; i += NT) {
        const int r = i / C, c = i - r * C;
        float acc = b3[c];
        for (int k = 0; k < H2; ++k) acc = fmaf(stg1[k * CHUNK + r], W3s[k * LD3 + c], acc);
        dls[r * LD3 + c] = acc;
        if (r < nv && a.logits) a.logits[(size_t)(c0 + r) * C + c] = acc;
      }
      __syncthreads();
      if ((a.flags & MLP_CE) && tid < nv) {
        const float* row = dls + tid * LD3;
        float mx = row[0]; int am = 0;
        for (int c = 1; c < C; ++c) if (row[c] > mx) { mx = row[c]; am = c; }
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
        const int lab = (int)a.labels[src_row(c0 + tid)];
        lsum += logf(se) + mx - row[lab];
        ok += (am == lab);
        if (a.flags & MLP_BWD) {
          const float inv = 1.f / ((float)B * se);
          for (int c = 0; c < C; ++c)
            a.dlog[(size_t)(c0 + tid) * CP + c] = expf(row[c] - mx) * inv - (c == lab ? 1.f / (float)B : 0.f);
        }
      }
      __syncthreads();
    }
    if (a.flags & MLP_CE) {
      // loss / correct: warp 0 of each CTA holds the per-row terms
      if (tid < 32) {
        lsum = warp_sum(lsum);
        for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
        if (tid == 0) { part[0] = lsum; part[1] = (float)ok; }
      }
      cluster_sum<NCTA>(cl, part, 1, 2, tot);
      if (rank == 0 && tid == 0) {
        if (a.loss) a.loss[0] = (float)(tot[0] / B);
        if (a.correct) a.correct[0] = (int)(tot[1] + 0.5);
      }
    }
  }

  // =============================== backward ===============================
  if (a.flags & MLP_BWD) {
    float* gp = a.grads;
    if (!(a.flags & MLP_FWD)) {
      // backward-only call: reload the BatchNorm coefficients saved by the forward call
      for (int i = tid; i < 4 * H1; i += NT) { c1[i] = a.bnc[i]; c2[i] = a.bnc[4 * H1 + i]; }
      __syncthreads();
    }
    const float* dl_src = (a.flags & MLP_CE) ? a.dlog : a.dlogits_in;
    const int dl_ld = (a.flags & MLP_CE) ? CP : C;

    // ---- pass 1: dW3, db3, dz2 = (dlogits W3) * relu'(.), statistics of dz2 ----
    {
      float accw[3] = {0.f, 0.f, 0.f};   // outputs tid, tid+256, tid+512 of the C*64 (<= 768) weight gradients
      float accb = 0.f;                  // tid < C: db3
      float s1 = 0.f, s2 = 0.f;
      for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int nv = min(CHUNK, r1 - c0);
        for (int i = tid; i < CHUNK * CP; i += NT) {
          const int r = i / CP, c = i - r * CP;
          dls[r * LD3 + c] = (r < nv && c < C) ? dl_src[(size_t)(c0 + r) * dl_ld + c] : 0.f;
        }
        for (int i = tid; i < H2 * CHUNK; i += NT) {
          const int r = i / H2, k = i - r * H2;
          stg1[k * CHUNK + r] = r < nv ? a.h2[(size_t)(c0 + r) * H2 + k] : 0.f;   // raw h2
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int o = tid + q * NT;
          if (o < C * H2) {
            const int c = o / H2, k = o - c * H2;
            float s = 0.f;
            for (int r = 0; r < CHUNK; ++r)
              s = fmaf(dls[r * LD3 + c], fmaxf(fmaf(stg1[k * CHUNK + r], c2[k], c2[H1 + k]), 0.f), s);
            accw[q] += s;
          }
        }
        if (tid < C) { float s = 0.f; for (int r = 0; r < CHUNK; ++r) s += dls[r * LD3 + tid]; accb += s; }
        {
          const int k = tid % H2, rg = tid / H2;   // 4 row groups x 8 rows
          for (int i = 0; i < 8; ++i) {
            const int r = rg * 8 + i;
            if (r < nv) {
              float d = 0.f;
              for (int c = 0; c < C; ++c) d = fmaf(dls[r * LD3 + c], W3s[k * LD3 + c], d);
              const float hv = stg1[k * CHUNK + r];
              const float zn = fmaf(hv, c2[k], c2[H1 + k]);
              d = zn > 0.f ? d : 0.f;
              a.d2[(size_t)(c0 + r) * H2 + k] = d;
              s1 += d;
              s2 += d * ((hv - c2[2 * H1 + k]) * c2[3 * H1 + k]);
            }
          }
        }
        __syncthreads();
      }
      {
        const int k = tid % H2, rg = tid / H2;
        part[(rg * 2 + 0) * H2 + k] = s1;
        part[(rg * 2 + 1) * H2 + k] = s2;
      }
      cluster_sum<NCTA>(cl, part, 4, 2 * H2, tot);
      for (int k = tid; k < H2; k += NT) {
        const double S1 = tot[k], S2 = tot[H2 + k];
        const double rstd = c2[3 * H1 + k];
        const double A = (double)g2[k] * rstd, Bc = -A * rstd * S2 / B, Cc = -A * S1 / B;
        c2[4 * H1 + k] = (float)A; c2[5 * H1 + k] = (float)Bc; c2[6 * H1 + k] = (float)Cc;
        if (rank == 0) { gp[a.off[6] + k] = (float)S2; gp[a.off[7] + k] = (float)S1; gp[a.off[5] + k] = 0.f; }
      }
      // exchange dW3 / db3 partials
#pragma unroll
      for (int q = 0; q < 3; ++q) { const int o = tid + q * NT; if (o < C * H2) gW[o] = accw[q]; }
      if (tid < C) gW[C * H2 + tid] = accb;
      cl.sync();
      {
        const int n = C * H2 + C;
        for (int i = tid + (int)rank * NT; i < n; i += NT * NCTA) {
          float s = 0.f;
          for (unsigned r = 0; r < NCTA; ++r) s += cl.map_shared_rank(gW, r)[i];
          if (i < C * H2) gp[a.off[8] + i] = s; else gp[a.off[9] + i - C * H2] = s;
        }
      }
      cl.sync();
    }

    // ---- pass 2: dh2 = BN backward(dz2); dW2 = dh2^T a1; dz1 = (dh2 W2) * dropout * relu'; statistics ----
    {
      float accw[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) accw[i] = 0.f;
      float s1 = 0.f, s2 = 0.f;
      const int kk = tid % H1, jg = tid / H1;   // dW2: k = kk, j in [jg*32, jg*32+32)
      for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int nv = min(CHUNK, r1 - c0);
        for (int i = tid; i < H2 * CHUNK; i += NT) {   // dh2^T
          const int r = i / H2, k = i - r * H2;
          float v = 0.f;
          if (r < nv) {
            const size_t g = (size_t)(c0 + r) * H2 + k;
            v = fmaf(c2[4 * H1 + k], a.d2[g], fmaf(c2[5 * H1 + k], a.h2[g] - c2[2 * H1 + k], c2[6 * H1 + k]));
          }
          stg1[k * CHUNK + r] = v;
        }
        for (int i = tid; i < H1 * CHUNK; i += NT) {   // a1^T (recomputed)
          const int r = i / H1, k = i - r * H1;
          float v = 0.f;
          if (r < nv) {
            const size_t g = (size_t)(c0 + r) * H1 + k;
            v = fmaxf(fmaf(a.h1[g], c1[k], c1[H1 + k]), 0.f);
            if (training && a.p > 0.f) v = a.keep[g] ? v * keep_scale : 0.f;
          }
          stg0[k * CHUNK + r] = v;
        }
        __syncthreads();
        for (int q = 0; q < CHUNK / 4; ++q) {
          const float4 av = *reinterpret_cast<const float4*>(stg0 + kk * CHUNK + q * 4);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 dv = *reinterpret_cast<const float4*>(stg1 + (jg * 32 + j) * CHUNK + q * 4);
            accw[j] = fmaf(av.x, dv.x, fmaf(av.y, dv.y, fmaf(av.z, dv.z, fmaf(av.w, dv.w, accw[j]))));
          }
        }
        {  // dz1 for 16 rows per thread (4 at a time): k = kk, rows jg*16 ..
          for (int q = 0; q < 4; ++q) {
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < H2; ++j) {
              const float w = W2s[kk * LD2 + j];
              const float4 dv = *reinterpret_cast<const float4*>(stg1 + j * CHUNK + jg * 16 + q * 4);
              d4[0] = fmaf(dv.x, w, d4[0]); d4[1] = fmaf(dv.y, w, d4[1]);
              d4[2] = fmaf(dv.z, w, d4[2]); d4[3] = fmaf(dv.w, w, d4[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = jg * 16 + q * 4 + i;
              if (r < nv) {
                float d = d4[i];
                const size_t g = (size_t)(c0 + r) * H1 + kk;
                const float hv = a.h1[g];
                if (training && a.p > 0.f) d = a.keep[g] ? d * keep_scale : 0.f;
                d = fmaf(hv, c1[kk], c1[H1 + kk]) > 0.f ? d : 0.f;
                a.d1[g] = d;
                s1 += d;
                s2 += d * ((hv - c1[2 * H1 + kk]) * c1[3 * H1 + kk]);
              }
            }
          }
        }
        __syncthreads();
      }
      part[(jg * 2 + 0) * H1 + kk] = s1;
      part[(jg * 2 + 1) * H1 + kk] = s2;
      cluster_sum<NCTA>(cl, part, 2, 2 * H1, tot);
      for (int k = tid; k < H1; k += NT) {
        const double S1 = tot[k], S2 = tot[H1 + k];
        const double rstd = c1[3 * H1 + k];
        const double A = (double)g1[k] * rstd, Bc = -A * rstd * S2 / B, Cc = -A * S1 / B;
        c1[4 * H1 + k] = (float)A; c1[5 * H1 + k] = (float)Bc; c1[6 * H1 + k] = (float)Cc;
        if (rank == 0) { gp[a.off[2] + k] = (float)S2; gp[a.off[3] + k] = (float)S1; gp[a.off[1] + k] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) gW[(jg * 32 + j) * H1 + kk] = accw[j];   // dW2[j][k]
      cl.sync();
      for (int i = tid + (int)rank * NT; i < H2 * H1; i += NT * NCTA) {
        float s = 0.f;
        for (unsigned r = 0; r < NCTA; ++r) s += cl.map_shared_rank(gW, r)[i];
        gp[a.off[4] + i] = s;
      }
      cl.sync();
    }

    // ---- pass 3: dh1 = BN backward(dz1); dW1 = dh1^T x ----
    {
      float accw[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) accw[i] = 0.f;
      const int jj = tid % H1, kg = tid / H1;   // dW1: j = jj, k in [kg*D/2, (kg+1)*D/2)
      const int kh = D / 2;
      for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int nv = min(CHUNK, r1 - c0);
        for (int i = tid; i < H1 * CHUNK; i += NT) {
          const int r = i / H1, k = i - r * H1;
          float v = 0.f;
          if (r < nv) {
            const size_t g = (size_t)(c0 + r) * H1 + k;
            v = fmaf(c1[4 * H1 + k], a.d1[g], fmaf(c1[5 * H1 + k], a.h1[g] - c1[2 * H1 + k], c1[6 * H1 + k]));
          }
          stg0[k * CHUNK + r] = v;
        }
        for (int i = tid; i < D * CHUNK; i += NT) {
          const int r = i / D, k = i - r * D;
          stg1[k * CHUNK + r] = r < nv ? a.x[src_row(c0 + r) * D + k] : 0.f;
        }
        __syncthreads();
        for (int q = 0; q < CHUNK / 4; ++q) {
          const float4 dv = *reinterpret_cast<const float4*>(stg0 + jj * CHUNK + q * 4);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            if (k < kh) {
              const float4 xv = *reinterpret_cast<const float4*>(stg1 + (kg * kh + k) * CHUNK + q * 4);
              accw[k] = fmaf(dv.x, xv.x, fmaf(dv.y, xv.y, fmaf(dv.z, xv.z, fmaf(dv.w, xv.w, accw[k]))));
            }
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) if (k < kh) gW[jj * D + kg * kh + k] = accw[k];   // dW1[j][k]
      cl.sync();
      for (int i = tid + (int)rank * NT; i < H1 * D; i += NT * NCTA) {
        float s = 0.f;
        for (unsigned r = 0; r < NCTA; ++r) s += cl.map_shared_rank(gW, r)[i];
        gp[a.off[0] + i] = s;
      }
      cl.sync();
    }
  }
  (void)misc;
  if (a.seed_dev || a.nbt || a.cursor) {
    cl.sync();                                 // every CTA has read the seed and the cursor
    if (rank == 0 && tid == 0) {
      if (a.seed_dev) a.seed_dev[0] += 0x9E3779B97F4A7C15ull;
      if (a.nbt && training) { a.nbt[0] += 1; a.nbt[1] += 1; }
      if (a.cursor) {
        if (a.hist) { a.hist[2 * a.cursor[1]] = a.loss[0]; a.hist[2 * a.cursor[1] + 1] = (float)a.correct[0]; }   // written by this thread above
        a.cursor[0] += B; a.cursor[1] += 1;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// k_mlp_small: the whole training step's forward + cross-entropy + backward for a batch of at most 64 rows (the reference
// trains the MLP at batch 64, NB:3443) in ONE CTA of 512 threads with every intermediate resident in shared memory.
// The cluster kernel above spreads the rows over up to 8 CTAs and pays ~10 cluster-wide exchanges plus a global round trip
// of every intermediate per phase; at this size that latency, not arithmetic, was the step (82 us).  Here the only
// synchronisation is __syncthreads, activations are kept k-major ([feature][row]) so that GEMM operands are 16-byte loads,
// and each thread's outputs stay in registers across the BatchNorm statistics.  Same arithmetic as k_mlp (same dropout
// hash, fp32 partial sums combined in double, BatchNorm backward as A*dz + B*(h-mean) + C).
// ---------------------------------------------------------------------------------------------
static constexpr int SB = 64, NTS = 512, LW1 = 130, LW2 = 66;
static constexpr int SP = 68;   // row pitch of the [feature][row] buffers: 16-byte loads of 32 consecutive features hit distinct banks
static constexpr size_t MLP_SMALL_SMEM =
    sizeof(float) * (64 * LW1 + H1 * LW2 + H2 * LD3 + 64 * SP + 2 * H1 * SP + 2 * H2 * SP + SB * LD3 + 16 * SP + 2048 + 14 * H1 + 64);

__device__ __forceinline__ void fma2s(float& d0, float& d1, float a, float b0, float b1) {
  const float2 r = __ffma2_rn(make_float2(a, a), make_float2(b0, b1), make_float2(d0, d1));
  d0 = r.x; d1 = r.y;
}

__global__ void __launch_bounds__(NTS, 1) k_mlp_small(MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, D = a.D, C = a.C, B = a.B;
  float* W1s = sm;                       // [D][LW1]   W1s[k][j] = W1[j][k]
  float* W2s = W1s + 64 * LW1;           // [128][LW2] W2s[k][j] = W2[j][k]
  float* W3s = W2s + H1 * LW2;           // [64][LD3]  W3s[k][c] = W3[c][k]
  float* xT = W3s + H2 * LD3;            // [D][SB]
  float* h1T = xT + 64 * SP;             // [128][SB] raw layer-1 outputs
  float* a1T = h1T + H1 * SP;            // [128][SB] dropout(relu(bn1(h1)));  later dh1
  float* h2T = a1T + H1 * SP;            // [64][SB]
  float* a2T = h2T + H2 * SP;            // [64][SB] relu(bn2(h2));  later dh2
  float* dls = a2T + H2 * SP;            // [SB][LD3] logits, then dlogits
  float* dlT = dls + SB * LD3;           // [16][SB]  dlogits, class-major
  float* part = dlT + 16 * SP;           // [row groups][2][columns] partial statistics (2048 floats)
  float* c1 = part + 2048;               // [7][128]: scale, shift, mean, rstd, A, B, C
  float* c2 = c1 + 7 * H1;
  float* misc = c2 + 7 * H1;             // [64]

  const float* W1 = a.params + a.off[0]; const float* b1 = a.params + a.off[1];
  const float* g1 = a.params + a.off[2]; const float* be1 = a.params + a.off[3];
  const float* W2 = a.params + a.off[4]; const float* b2 = a.params + a.off[5];
  const float* g2 = a.params + a.off[6]; const float* be2 = a.params + a.off[7];
  const float* W3 = a.params + a.off[8]; const float* b3 = a.params + a.off[9];
  float* gp = a.grads;
  const unsigned long long seed = a.seed + (a.seed_dev ? a.seed_dev[0] : 0ull);
  const long long row_base = a.order ? a.cursor[0] : 0;
  const float keep_scale = 1.f / (1.f - a.p);
  const bool drop = a.p > 0.f;
#ifdef AE_TRACE
  long long tr_t[12]; int tr_n = 0;
#define MLP_TR() do { if (tid == 0) tr_t[tr_n++] = clock64(); } while (0)
#else
#define MLP_TR() do {} while (0)
#endif
  MLP_TR();

  // ---- stage: weights (transposed), the batch (gathered, transposed, zero rows behind B) ----
  stage_weight_T<NTS>(W1, W1s, H1, D, LW1, tid);
  stage_weight_T<NTS>(W2, W2s, H2, H1, LW2, tid);
  for (int i = tid; i < H2 * LD3; i += NTS) {
    const int k = i / LD3, c = i - k * LD3;
    W3s[i] = c < C ? W3[c * H2 + k] : 0.f;
  }
  {
    const int d4 = D / 4;
    for (int i = tid; i < SB * d4; i += NTS) {
      const int r = i / d4, k = (i - r * d4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < B) {
        const size_t row = a.order ? (size_t)a.order[row_base + r] : (size_t)r;
        v = __ldg(reinterpret_cast<const float4*>(a.x + row * D + k));
      }
      xT[(k + 0) * SP + r] = v.x; xT[(k + 1) * SP + r] = v.y; xT[(k + 2) * SP + r] = v.z; xT[(k + 3) * SP + r] = v.w;
    }
  }
  __syncthreads();

  MLP_TR();
  // =============================== forward ===============================
  // ---- layer 1: thread = 8 rows x 2 columns of h1 [64 x 128] ----
  {
    const int j0 = (tid & 63) * 2, rg = tid >> 6, r0 = rg * 8;
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(xT + k * SP + r0), x1 = *reinterpret_cast<const float4*>(xT + k * SP + r0 + 4);
      const float2 w = *reinterpret_cast<const float2*>(W1s + k * LW1 + j0);
      fma2s(acc[0][0], acc[0][1], x0.x, w.x, w.y); fma2s(acc[1][0], acc[1][1], x0.y, w.x, w.y);
      fma2s(acc[2][0], acc[2][1], x0.z, w.x, w.y); fma2s(acc[3][0], acc[3][1], x0.w, w.x, w.y);
      fma2s(acc[4][0], acc[4][1], x1.x, w.x, w.y); fma2s(acc[5][0], acc[5][1], x1.y, w.x, w.y);
      fma2s(acc[6][0], acc[6][1], x1.z, w.x, w.y); fma2s(acc[7][0], acc[7][1], x1.w, w.x, w.y);
    }
    const float bb[2] = {b1[j0], b1[j0 + 1]};
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        acc[i][c] += bb[c];
        if (r0 + i < B) { s1[c] += acc[i][c]; s2[c] += acc[i][c] * acc[i][c]; }
      }
    part[(rg * 2 + 0) * H1 + j0] = s1[0]; part[(rg * 2 + 0) * H1 + j0 + 1] = s1[1];
    part[(rg * 2 + 1) * H1 + j0] = s2[0]; part[(rg * 2 + 1) * H1 + j0 + 1] = s2[1];
    __syncthreads();
    // layer 1's weights are not needed again: their space takes W2 in its natural [j][k] orientation (pitch LW1), which the
    // backward pass reads as conflict-free 8-byte pairs (W2s[k][j] with k = 2 * lane is a 4-way bank conflict there)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i4 = tid + u * NTS;                         // 2048 float4 = 64 x 128
      const float4 wv = __ldg(reinterpret_cast<const float4*>(W2) + i4);
      const int j = i4 >> 5, k = (i4 & 31) * 4;
      *reinterpret_cast<float2*>(W1s + j * LW1 + k) = make_float2(wv.x, wv.y);
      *reinterpret_cast<float2*>(W1s + j * LW1 + k + 2) = make_float2(wv.z, wv.w);
    }
    if (tid < H1) {
      double S1 = 0.0, S2 = 0.0;
      for (int g = 0; g < 8; ++g) { S1 += (double)part[(g * 2 + 0) * H1 + tid]; S2 += (double)part[(g * 2 + 1) * H1 + tid]; }
      const double m = S1 / B;
      double v = S2 / B - m * m;
      if (v < 0.0) v = 0.0;
      if (a.running) {
        const double unb = B > 1 ? v * B / (B - 1.0) : v;
        a.running[tid] = (float)(0.9 * (double)a.running[tid] + 0.1 * m);
        a.running[H1 + tid] = (float)(0.9 * (double)a.running[H1 + tid] + 0.1 * unb);
      }
      const float mean = (float)m, rstd = 1.f / sqrtf((float)v + BN_EPS_F), sc = g1[tid] * rstd;
      c1[tid] = sc; c1[H1 + tid] = be1[tid] - mean * sc; c1[2 * H1 + tid] = mean; c1[3 * H1 + tid] = rstd;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float sc = c1[j], sh = c1[H1 + j];
      float hv[8], av[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + i;
        float h = acc[i][c], v = fmaxf(fmaf(h, sc, sh), 0.f);
        if (drop) v = (hash_u32(seed, r, j) * (1.0f / 4294967296.0f)) >= a.p ? v * keep_scale : 0.f;
        if (r >= B) { h = 0.f; v = 0.f; }
        hv[i] = h; av[i] = v;
      }
      *reinterpret_cast<float4*>(h1T + j * SP + r0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(h1T + j * SP + r0 + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      *reinterpret_cast<float4*>(a1T + j * SP + r0) = make_float4(av[0], av[1], av[2], av[3]);
      *reinterpret_cast<float4*>(a1T + j * SP + r0 + 4) = make_float4(av[4], av[5], av[6], av[7]);
    }
    __syncthreads();
  }
  MLP_TR();
  // ---- layer 2: thread = 4 rows x 2 columns of h2 [64 x 64] ----
  {
    const int j0 = (tid & 31) * 2, rg = tid >> 5, r0 = rg * 4;
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < H1; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(a1T + k * SP + r0);
      const float2 w = *reinterpret_cast<const float2*>(W2s + k * LW2 + j0);
      fma2s(acc[0][0], acc[0][1], x0.x, w.x, w.y); fma2s(acc[1][0], acc[1][1], x0.y, w.x, w.y);
      fma2s(acc[2][0], acc[2][1], x0.z, w.x, w.y); fma2s(acc[3][0], acc[3][1], x0.w, w.x, w.y);
    }
    const float bb[2] = {b2[j0], b2[j0 + 1]};
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        acc[i][c] += bb[c];
        if (r0 + i < B) { s1[c] += acc[i][c]; s2[c] += acc[i][c] * acc[i][c]; }
      }
    part[(rg * 2 + 0) * H2 + j0] = s1[0]; part[(rg * 2 + 0) * H2 + j0 + 1] = s1[1];
    part[(rg * 2 + 1) * H2 + j0] = s2[0]; part[(rg * 2 + 1) * H2 + j0 + 1] = s2[1];
    __syncthreads();
    if (tid < H2) {
      double S1 = 0.0, S2 = 0.0;
      for (int g = 0; g < 16; ++g) { S1 += (double)part[(g * 2 + 0) * H2 + tid]; S2 += (double)part[(g * 2 + 1) * H2 + tid]; }
      const double m = S1 / B;
      double v = S2 / B - m * m;
      if (v < 0.0) v = 0.0;
      if (a.running) {
        const double unb = B > 1 ? v * B / (B - 1.0) : v;
        a.running[2 * H1 + tid] = (float)(0.9 * (double)a.running[2 * H1 + tid] + 0.1 * m);
        a.running[2 * H1 + H2 + tid] = (float)(0.9 * (double)a.running[2 * H1 + H2 + tid] + 0.1 * unb);
      }
      const float mean = (float)m, rstd = 1.f / sqrtf((float)v + BN_EPS_F), sc = g2[tid] * rstd;
      c2[tid] = sc; c2[H1 + tid] = be2[tid] - mean * sc; c2[2 * H1 + tid] = mean; c2[3 * H1 + tid] = rstd;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float sc = c2[j], sh = c2[H1 + j];
      float hv[4], av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = r0 + i < B;
        hv[i] = ok ? acc[i][c] : 0.f;
        av[i] = ok ? fmaxf(fmaf(acc[i][c], sc, sh), 0.f) : 0.f;
      }
      *reinterpret_cast<float4*>(h2T + j * SP + r0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(a2T + j * SP + r0) = make_float4(av[0], av[1], av[2], av[3]);
    }
    __syncthreads();
  }
  MLP_TR();
  // ---- layer 3 + cross-entropy ----
  for (int o = tid; o < SB * C; o += NTS) {
    const int r = o & (SB - 1), c = o >> 6;
    float acc = b3[c];
#pragma unroll 8
    for (int k = 0; k < H2; ++k) acc = fmaf(a2T[k * SP + r], W3s[k * LD3 + c], acc);
    dls[r * LD3 + c] = acc;
    if (r < B && a.logits) a.logits[(size_t)r * C + c] = acc;
  }
  __syncthreads();
  if (tid < SB) {
    float lterm = 0.f;
    int ok = 0;
    float* row = dls + tid * LD3;
    if (tid < B) {
      float mx = row[0]; int am = 0;
      for (int c = 1; c < C; ++c) if (row[c] > mx) { mx = row[c]; am = c; }
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
      const size_t src = a.order ? (size_t)a.order[row_base + tid] : (size_t)tid;
      const int lab = (int)a.labels[src];
      lterm = logf(se) + mx - row[lab];
      ok = am == lab;
      const float inv = 1.f / ((float)B * se);
      for (int c = 0; c < C; ++c) {
        const float d = expf(row[c] - mx) * inv - (c == lab ? 1.f / (float)B : 0.f);
        row[c] = d; dlT[c * SP + tid] = d;
      }
    } else {
      for (int c = 0; c < C; ++c) { row[c] = 0.f; dlT[c * SP + tid] = 0.f; }
    }
    lterm = warp_sum(lterm);
    for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
    if ((tid & 31) == 0) { misc[(tid >> 5) * 2] = lterm; misc[(tid >> 5) * 2 + 1] = (float)ok; }
  }
  __syncthreads();
  if (tid == 0) {
    if (a.loss) a.loss[0] = (float)(((double)misc[0] + (double)misc[2]) / B);
    if (a.correct) a.correct[0] = (int)(misc[1] + misc[3] + 0.5f);
  }

  MLP_TR();
  // =============================== backward ===============================
  // ---- dW3, db3; dz2 = (dlogits W3) * relu'(bn2(h2)) kept in registers, its statistics ----
  {
    for (int o = tid; o < C * H2; o += NTS) {
      const int c = o >> 6, k = o & (H2 - 1);
      float s = 0.f;
#pragma unroll 4
      for (int q = 0; q < SB / 4; ++q) {
        const float4 dv = *reinterpret_cast<const float4*>(dlT + c * SP + q * 4), av = *reinterpret_cast<const float4*>(a2T + k * SP + q * 4);
        s = fmaf(dv.x, av.x, fmaf(dv.y, av.y, fmaf(dv.z, av.z, fmaf(dv.w, av.w, s))));
      }
      gp[a.off[8] + o] = s;
    }
    if (tid < C) {
      float s = 0.f;
      for (int r = 0; r < SB; ++r) s += dlT[tid * SP + r];
      gp[a.off[9] + tid] = s;
    }
    const int j0 = (tid & 31) * 2, rg = tid >> 5, r0 = rg * 4;
    float d[4][2], xh[4][2];
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float sc = c2[j], sh = c2[H1 + j], mean = c2[2 * H1 + j], rstd = c2[3 * H1 + j];
      const float4 hv4 = *reinterpret_cast<const float4*>(h2T + j * SP + r0);
      const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float t = 0.f;
        for (int cc = 0; cc < C; ++cc) t = fmaf(dls[(r0 + i) * LD3 + cc], W3s[j * LD3 + cc], t);
        t = fmaf(hv[i], sc, sh) > 0.f ? t : 0.f;          // rows behind B: dlogits are zero
        d[i][c] = t;
        xh[i][c] = (hv[i] - mean) * rstd;
        s1[c] += t; s2[c] += t * xh[i][c];
      }
    }
    part[(rg * 2 + 0) * H2 + j0] = s1[0]; part[(rg * 2 + 0) * H2 + j0 + 1] = s1[1];
    part[(rg * 2 + 1) * H2 + j0] = s2[0]; part[(rg * 2 + 1) * H2 + j0 + 1] = s2[1];
    __syncthreads();                                         // dW3 has read a2T; the partials are in
    if (tid < H2) {
      double S1 = 0.0, S2 = 0.0;
      for (int g = 0; g < 16; ++g) { S1 += (double)part[(g * 2 + 0) * H2 + tid]; S2 += (double)part[(g * 2 + 1) * H2 + tid]; }
      const double rstd = c2[3 * H1 + tid];
      const double A = (double)g2[tid] * rstd, Bc = -A * rstd * S2 / B, Cc = -A * S1 / B;
      c2[4 * H1 + tid] = (float)A; c2[5 * H1 + tid] = (float)Bc; c2[6 * H1 + tid] = (float)Cc;
      gp[a.off[6] + tid] = (float)S2; gp[a.off[7] + tid] = (float)S1; gp[a.off[5] + tid] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float A = c2[4 * H1 + j], Bc = c2[5 * H1 + j], Cc = c2[6 * H1 + j], rinv = 1.f / c2[3 * H1 + j];
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = r0 + i < B ? fmaf(A, d[i][c], fmaf(Bc, xh[i][c] * rinv, Cc)) : 0.f;   // xh/rstd = h - mean
      *reinterpret_cast<float4*>(a2T + j * SP + r0) = make_float4(o[0], o[1], o[2], o[3]);                   // dh2^T
    }
    __syncthreads();
  }
  MLP_TR();
  // ---- dW2 = dh2^T a1; dz1 = (dh2 W2) * dropout * relu'(bn1(h1)) in registers, its statistics ----
  {
    {
      // dW2[j][k]: 4 x 4 register tile, k = lane + 32 i (consecutive lanes read consecutive rows: conflict-free with the
      // padded pitch), j = 4 jg + jj (broadcast loads)
      const int lane = tid & 31, jg = tid >> 5;
      float acc[4][4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[jj][i] = 0.f;
#pragma unroll 2
      for (int q = 0; q < SB / 4; ++q) {
        float4 av[4], dv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a1T + (lane + 32 * i) * SP + q * 4);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) dv[jj] = *reinterpret_cast<const float4*>(a2T + (jg * 4 + jj) * SP + q * 4);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[jj][i] = fmaf(av[i].x, dv[jj].x, fmaf(av[i].y, dv[jj].y, fmaf(av[i].z, dv[jj].z, fmaf(av[i].w, dv[jj].w, acc[jj][i]))));
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < 4; ++i) gp[a.off[4] + (jg * 4 + jj) * H1 + lane + 32 * i] = acc[jj][i];
    }
    const int j0 = (tid & 63) * 2, rg = tid >> 6, r0 = rg * 8;
    float d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i][0] = 0.f; d[i][1] = 0.f; }
#pragma unroll 4
    for (int j = 0; j < H2; ++j) {
      const float4 x0 = *reinterpret_cast<const float4*>(a2T + j * SP + r0), x1 = *reinterpret_cast<const float4*>(a2T + j * SP + r0 + 4);
      const float2 w01 = *reinterpret_cast<const float2*>(W1s + j * LW1 + j0);   // W2[j][j0], W2[j][j0 + 1]
      const float w0 = w01.x, w1 = w01.y;
      fma2s(d[0][0], d[0][1], x0.x, w0, w1); fma2s(d[1][0], d[1][1], x0.y, w0, w1);
      fma2s(d[2][0], d[2][1], x0.z, w0, w1); fma2s(d[3][0], d[3][1], x0.w, w0, w1);
      fma2s(d[4][0], d[4][1], x1.x, w0, w1); fma2s(d[5][0], d[5][1], x1.y, w0, w1);
      fma2s(d[6][0], d[6][1], x1.z, w0, w1); fma2s(d[7][0], d[7][1], x1.w, w0, w1);
    }
    float xh[8][2];
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float sc = c1[j], sh = c1[H1 + j], mean = c1[2 * H1 + j], rstd = c1[3 * H1 + j];
      const float4 h0 = *reinterpret_cast<const float4*>(h1T + j * SP + r0), h1v = *reinterpret_cast<const float4*>(h1T + j * SP + r0 + 4);
      const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1v.x, h1v.y, h1v.z, h1v.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + i;
        float t = d[i][c];
        if (drop) t = (hash_u32(seed, r, j) * (1.0f / 4294967296.0f)) >= a.p ? t * keep_scale : 0.f;
        t = (r < B && fmaf(hv[i], sc, sh) > 0.f) ? t : 0.f;
        d[i][c] = t;
        xh[i][c] = (hv[i] - mean) * rstd;
        s1[c] += t; s2[c] += t * xh[i][c];
      }
    }
    part[(rg * 2 + 0) * H1 + j0] = s1[0]; part[(rg * 2 + 0) * H1 + j0 + 1] = s1[1];
    part[(rg * 2 + 1) * H1 + j0] = s2[0]; part[(rg * 2 + 1) * H1 + j0 + 1] = s2[1];
    __syncthreads();                                         // dW2 has read a1T; the partials are in
    if (tid < H1) {
      double S1 = 0.0, S2 = 0.0;
      for (int g = 0; g < 8; ++g) { S1 += (double)part[(g * 2 + 0) * H1 + tid]; S2 += (double)part[(g * 2 + 1) * H1 + tid]; }
      const double rstd = c1[3 * H1 + tid];
      const double A = (double)g1[tid] * rstd, Bc = -A * rstd * S2 / B, Cc = -A * S1 / B;
      c1[4 * H1 + tid] = (float)A; c1[5 * H1 + tid] = (float)Bc; c1[6 * H1 + tid] = (float)Cc;
      gp[a.off[2] + tid] = (float)S2; gp[a.off[3] + tid] = (float)S1; gp[a.off[1] + tid] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      const float A = c1[4 * H1 + j], Bc = c1[5 * H1 + j], Cc = c1[6 * H1 + j], rinv = 1.f / c1[3 * H1 + j];
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = r0 + i < B ? fmaf(A, d[i][c], fmaf(Bc, xh[i][c] * rinv, Cc)) : 0.f;
      *reinterpret_cast<float4*>(a1T + j * SP + r0) = make_float4(o[0], o[1], o[2], o[3]);                   // dh1^T
      *reinterpret_cast<float4*>(a1T + j * SP + r0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
  }
  MLP_TR();
  // ---- dW1 = dh1^T x ----
  {
    // dW1[j][k]: 4 x 4 register tile, j = lane + 32 i, k = 4 kg + kk (threads whose k group lies behind D idle)
    const int lane = tid & 31, kg = tid >> 5;
    if (kg * 4 < D) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) acc[i][kk] = 0.f;
#pragma unroll 2
      for (int q = 0; q < SB / 4; ++q) {
        float4 dv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = *reinterpret_cast<const float4*>(a1T + (lane + 32 * i) * SP + q * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) xv[kk] = *reinterpret_cast<const float4*>(xT + (kg * 4 + kk) * SP + q * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            acc[i][kk] = fmaf(dv[i].x, xv[kk].x, fmaf(dv[i].y, xv[kk].y, fmaf(dv[i].z, xv[kk].z, fmaf(dv[i].w, xv[kk].w, acc[i][kk]))));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(gp + a.off[0] + (size_t)(lane + 32 * i) * D + kg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
  MLP_TR();
#ifdef AE_TRACE
  if (tid == 0 && a.cursor == nullptr && (a.seed_dev == nullptr || (a.seed_dev[0] & 0xff) == 0x15))
    printf("k_mlp_small clk: stage %lld  L1 %lld  L2 %lld  L3+CE %lld  bwd3 %lld  bwd2 %lld  bwd1 %lld\n", tr_t[1] - tr_t[0], tr_t[2] - tr_t[1],
           tr_t[3] - tr_t[2], tr_t[4] - tr_t[3], tr_t[5] - tr_t[4], tr_t[6] - tr_t[5], tr_t[7] - tr_t[6]);
#endif
  if (tid == 0) {
    if (a.seed_dev) a.seed_dev[0] += 0x9E3779B97F4A7C15ull;
    if (a.nbt) { a.nbt[0] += 1; a.nbt[1] += 1; }
    if (a.cursor) {
      if (a.hist) { a.hist[2 * a.cursor[1]] = a.loss[0]; a.hist[2 * a.cursor[1] + 1] = (float)a.correct[0]; }
      a.cursor[0] += B; a.cursor[1] += 1;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Inference (NB:3499, NB:3702): running statistics, no dropout.  No cross-row dependency, so the grid
// is one CTA per SM looping over 32-row chunks with the weights resident in shared memory.
// ---------------------------------------------------------------------------------------------
template <int NOUT>
__device__ __forceinline__ void tile_linear_act(const float* __restrict__ inT, int K, const float* __restrict__ Ws, int ld,
                                                const float* __restrict__ bias, const float* __restrict__ sc,
                                                const float* __restrict__ sh, float* __restrict__ outT) {
  constexpr int G = NT / NOUT;
  constexpr int RPT = CHUNK / G;
  const int j = threadIdx.x % NOUT, rg = threadIdx.x / NOUT;
  float acc[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float w = Ws[k * ld + j];
#pragma unroll
    for (int q = 0; q < RPT / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(inT + k * CHUNK + rg * RPT + q * 4);
      acc[q * 4 + 0] = fmaf(v.x, w, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(v.y, w, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(v.z, w, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(v.w, w, acc[q * 4 + 3]);
    }
  }
  const float b = bias[j], s = sc[j], t = sh[j];
#pragma unroll
  for (int i = 0; i < RPT; ++i) outT[j * CHUNK + rg * RPT + i] = fmaxf(fmaf(acc[i] + b, s, t), 0.f);
}

__global__ void __launch_bounds__(NT, 1) k_mlp_eval(const float* __restrict__ params, const float* __restrict__ running,
                                                    const float* __restrict__ x, int B, int D, int C, MlpArgs lay,
                                                    float* __restrict__ logits, int64_t* __restrict__ argmax) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  float* W1s = sm;
  float* W2s = W1s + D * LD1;
  float* W3s = W2s + H1 * LD2;
  float* stgA = W3s + H2 * LD3;
  float* stgB = stgA + H1 * CHUNK;
  float* cf = stgB + H1 * CHUNK;     // sc1, sh1 [128]; sc2, sh2 [64]
  float* dls = cf + 4 * H1;
  const float* W1 = params + lay.off[0]; const float* b1 = params + lay.off[1];
  const float* g1 = params + lay.off[2]; const float* be1 = params + lay.off[3];
  const float* W2 = params + lay.off[4]; const float* b2 = params + lay.off[5];
  const float* g2 = params + lay.off[6]; const float* be2 = params + lay.off[7];
  const float* W3 = params + lay.off[8]; const float* b3 = params + lay.off[9];
  stage_weight_T(W1, W1s, H1, D, LD1, tid);
  stage_weight_T(W2, W2s, H2, H1, LD2, tid);
  for (int i = tid; i < H2 * LD3; i += NT) W3s[i] = 0.f;
  for (int j = tid; j < H1; j += NT) {
    const float rstd = 1.f / sqrtf(running[H1 + j] + BN_EPS_F), s = g1[j] * rstd;
    cf[j] = s; cf[H1 + j] = be1[j] - running[j] * s;
  }
  for (int j = tid; j < H2; j += NT) {
    const float rstd = 1.f / sqrtf(running[2 * H1 + H2 + j] + BN_EPS_F), s = g2[j] * rstd;
    cf[2 * H1 + j] = s; cf[3 * H1 + j] = be2[j] - running[2 * H1 + j] * s;
  }
  __syncthreads();
  for (int i = tid; i < C * H2; i += NT) { const int c = i / H2, k = i - c * H2; W3s[k * LD3 + c] = W3[i]; }
  __syncthreads();
  const int nchunks = (B + CHUNK - 1) / CHUNK;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int c0 = ch * CHUNK, nv = min(CHUNK, B - c0);
    {
      // the chunk's rows: all of a thread's global loads first (8 x 256 = D * CHUNK at D = 64), then the transposing stores
      float xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * NT, r = i / D, k = i - r * D;
        xv[u] = (i < D * CHUNK && r < nv) ? __ldg(x + (size_t)(c0 + r) * D + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * NT, r = i / D, k = i - r * D;
        if (i < D * CHUNK) stgA[k * CHUNK + r] = xv[u];
      }
    }
    __syncthreads();
    tile_linear_act<H1>(stgA, D, W1s, LD1, b1, cf, cf + H1, stgB);
    __syncthreads();
    tile_linear_act<H2>(stgB, H1, W2s, LD2, b2, cf + 2 * H1, cf + 3 * H1, stgA);
    __syncthreads();
    for (int i = tid; i < CHUNK * C; i += NT) {
      const int r = i / C, c = i - r * C;
      float acc = b3[c];
      for (int k = 0; k < H2; ++k) acc = fmaf(stgA[k * CHUNK + r], W3s[k * LD3 + c], acc);
      dls[r * LD3 + c] = acc;
      if (r < nv && logits) logits[(size_t)(c0 + r) * C + c] = acc;
    }
    __syncthreads();
    if (argmax && tid < nv) {
      const float* row = dls + tid * LD3;
      float mx = row[0]; int am = 0;
      for (int c = 1; c < C; ++c) if (row[c] > mx) { mx = row[c]; am = c; }
      argmax[c0 + tid] = am;
    }
    __syncthreads();
  }
}

static size_t mlp_smem_bytes(int D) {
  size_t f = (size_t)D * LD1 + (size_t)H1 * LD2 + (size_t)H2 * LD3 + 2 * (size_t)H1 * CHUNK + 4 * 2 * H1;
  f += 2 * 2 * H1;        // tot (doubles)
  f += 2 * 7 * H1 + CHUNK * LD3 + 8192 + 64;
  return f * sizeof(float);
}

static int64_t mlp_layout(int D, int C, int64_t* off, int64_t* size) {
  const int64_t sz[10] = {(int64_t)H1 * D, H1, H1, H1, (int64_t)H2 * H1, H2, H2, H2, (int64_t)C * H2, C};
  int64_t o = 0;
  for (int i = 0; i < 10; ++i) {
    if (off) off[i] = o;
    if (size) size[i] = sz[i];
    o += (sz[i] + 3) & ~(int64_t)3;
  }
  return o;
}

struct MlpWs { float *h1, *h2, *d1, *d2, *dlog, *bnc; uint8_t* keep; };
static size_t mlp_carve(int B, char* base, MlpWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; char* p = base ? base + off : nullptr; off += bytes; return p; };
  float* h1 = (float*)take((size_t)B * H1 * 4); float* h2 = (float*)take((size_t)B * H2 * 4);
  float* d1 = (float*)take((size_t)B * H1 * 4); float* d2 = (float*)take((size_t)B * H2 * 4);
  float* dlog = (float*)take((size_t)B * CP * 4); float* bnc = (float*)take(2 * 4 * H1 * 4);
  uint8_t* keep = (uint8_t*)take((size_t)B * H1);
  if (w) { w->h1 = h1; w->h2 = h2; w->d1 = d1; w->d2 = d2; w->dlog = dlog; w->bnc = bnc; w->keep = keep; }
  return off + 256;
}

template <int NCTA>
static int mlp_launch_n(MlpArgs& a, cudaStream_t st) {
  const size_t smem = mlp_smem_bytes(a.D);
  static bool attr_set = false;
  if (!attr_set) {
    AE_CUDA(cudaFuncSetAttribute(k_mlp<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_smem_bytes(64)));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NCTA, 1, 1); cfg.blockDim = dim3(NT, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AE_CUDA(cudaLaunchKernelEx(&cfg, k_mlp<NCTA>, a));
  return 0;
}

static int mlp_launch(MlpArgs& a, cudaStream_t st) {
  AE_CHECK(a.D >= 4 && a.D <= 64 && a.D % 4 == 0, "mlp: input_dim=%d must be a multiple of 4 in [4,64]", a.D);
  AE_CHECK(a.C >= 2 && a.C <= 12, "mlp: num_classes=%d must be in [2,12]", a.C);
  AE_CHECK(a.B >= 1, "mlp: empty batch");
  mlp_layout(a.D, a.C, a.off, nullptr);
  // cluster size by batch (AE_B200_MLP_CLUSTER overrides, for measurements).  Measured at the reference's batch 64, whole
  // step replayed as a graph: 148 / 101 / 97 / 97 us with 1 / 2 / 4 / 8 CTAs -- one CTA is bound by arithmetic (256 threads),
  // from 4 CTAs up by the cluster-wide exchanges.
  int nc = a.B >= 48 ? 8 : a.B >= 24 ? 4 : a.B >= 12 ? 2 : 1;
  const char* ev = getenv("AE_B200_MLP_CLUSTER");
  if (ev) { const int v = atoi(ev); if (v == 1 || v == 2 || v == 4 || v == 8) nc = v; else ev = nullptr; }
  // the whole fused training step of a batch of at most 64 rows: one CTA, everything in shared memory (k_mlp_small);
  // AE_B200_MLP_CLUSTER=1/2/4/8 forces the cluster kernel (measurements, tests of every cluster size)
  if (!ev && a.B <= SB && a.flags == (MLP_FWD | MLP_TRAIN | MLP_CE | MLP_BWD) && !a.keep_in && !a.dlogits_in && a.labels && a.grads &&
      a.loss && a.correct) {
    static bool small_attr = false;
    if (!small_attr) {
      AE_CUDA(cudaFuncSetAttribute(k_mlp_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMALL_SMEM));
      small_attr = true;
    }
    k_mlp_small<<<1, NTS, MLP_SMALL_SMEM, st>>>(a);
    AE_LAUNCH_CHECK();
    return 0;
  }
  switch (nc) {
    case 1: return mlp_launch_n<1>(a, st);
    case 2: return mlp_launch_n<2>(a, st);
    case 4: return mlp_launch_n<4>(a, st);
    default: return mlp_launch_n<8>(a, st);
  }
}

int adam_step_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                   float gscale, int* step_dev, cudaStream_t st);   // elementwise.cu

}  // namespace ae

using namespace ae;

extern "C" {

int64_t ae_mlp_param_layout(int input_dim, int num_classes, int64_t* offsets, int64_t* sizes) {
  return mlp_layout(input_dim, num_classes, offsets, sizes);
}

size_t ae_mlp_workspace_bytes(int batch, int input_dim, int num_classes) {
  (void)input_dim; (void)num_classes;
  return mlp_carve(batch, nullptr, nullptr);
}

int ae_mlp_fwd_bwd_ce(const float* params, float* grads, float* bn_running, const float* x, const int64_t* labels,
                      const uint8_t* dropout_keep, uint64_t dropout_seed, float dropout_p, int batch, int input_dim,
                      int num_classes, int training, float* logits, float* loss, int* correct, void* workspace,
                      size_t workspace_bytes, ae_stream_t stream) {
  AE_CHECK(params && x && workspace, "ae_mlp_fwd_bwd_ce: null argument");
  AE_CHECK(workspace_bytes >= mlp_carve(batch, nullptr, nullptr), "ae_mlp_fwd_bwd_ce: workspace too small");
  AE_CHECK(training == 0 || bn_running != nullptr, "ae_mlp_fwd_bwd_ce: training needs the BatchNorm running buffers");
  AE_CHECK(dropout_p >= 0.f && dropout_p < 1.f, "ae_mlp_fwd_bwd_ce: dropout_p out of range");
  MlpWs w;
  mlp_carve(batch, (char*)workspace, &w);
  MlpArgs a{};
  a.params = params; a.grads = grads; a.running = bn_running; a.x = x; a.labels = labels; a.keep_in = dropout_keep;
  a.dlogits_in = nullptr; a.logits = logits; a.loss = loss; a.correct = correct;
  a.h1 = w.h1; a.h2 = w.h2; a.d1 = w.d1; a.d2 = w.d2; a.dlog = w.dlog; a.keep = w.keep; a.bnc = w.bnc;
  a.B = batch; a.D = input_dim; a.C = num_classes; a.seed = dropout_seed; a.p = training ? dropout_p : 0.f;
  a.flags = MLP_FWD;
  if (training) a.flags |= MLP_TRAIN;
  if (labels) a.flags |= MLP_CE;
  if (training && grads && labels) a.flags |= MLP_BWD;
  return mlp_launch(a, (cudaStream_t)stream);
}

int ae_mlp_forward_eval(const float* params, const float* bn_running, const float* x, int batch, int input_dim,
                        int num_classes, float* logits, int64_t* argmax, ae_stream_t stream) {
  AE_CHECK(params && bn_running && x && batch >= 1, "ae_mlp_forward_eval: bad argument");
  AE_CHECK(input_dim >= 4 && input_dim <= 64 && input_dim % 4 == 0, "mlp: input_dim=%d must be a multiple of 4 in [4,64]", input_dim);
  AE_CHECK(num_classes >= 2 && num_classes <= 12, "mlp: num_classes=%d must be in [2,12]", num_classes);
  const size_t smem = ((size_t)input_dim * LD1 + (size_t)H1 * LD2 + (size_t)H2 * LD3 + 2 * (size_t)H1 * CHUNK + 4 * H1 + CHUNK * LD3) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    AE_CUDA(cudaFuncSetAttribute(k_mlp_eval, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(((size_t)64 * LD1 + (size_t)H1 * LD2 + (size_t)H2 * LD3 + 2 * (size_t)H1 * CHUNK + 4 * H1 + CHUNK * LD3) * 4)));
    attr_set = true;
  }
  MlpArgs lay{};
  mlp_layout(input_dim, num_classes, lay.off, nullptr);
  const int nchunks = (batch + CHUNK - 1) / CHUNK;
  const int grid = nchunks < 148 ? nchunks : 148;
  k_mlp_eval<<<grid, NT, smem, (cudaStream_t)stream>>>(params, bn_running, x, batch, input_dim, num_classes, lay, logits, argmax);
  AE_LAUNCH_CHECK();
  return 0;
}

// One whole MLP training step (NB:3476-3482: zero_grad, forward, CE, backward, Adam.step) as two launches with nothing baked
// in that changes between steps: the dropout seed advances on the device, so the call can be captured in a CUDA graph and
// replayed.  Also bumps both BatchNorm layers' num_batches_tracked (bn_steps, int64[2]).
int ae_mlp_train_step(float* params, float* grads, float* bn_running, int64_t* bn_steps, const float* x, const int64_t* labels,
                      uint64_t dropout_seed, uint64_t* seed_dev, float dropout_p, int batch, int input_dim, int num_classes,
                      float* logits, float* loss, int* correct, void* workspace, size_t workspace_bytes,
                      const ae_adam_config_t* adam, float* adam_m, float* adam_v, int* step_dev, ae_stream_t stream) {
  AE_CHECK(params && grads && bn_running && x && labels && seed_dev && workspace && adam && adam_m && adam_v && step_dev,
           "ae_mlp_train_step: null argument");
  AE_CHECK(workspace_bytes >= mlp_carve(batch, nullptr, nullptr), "ae_mlp_train_step: workspace too small");
  AE_CHECK(dropout_p >= 0.f && dropout_p < 1.f, "ae_mlp_train_step: dropout_p out of range");
  MlpWs w;
  mlp_carve(batch, (char*)workspace, &w);
  MlpArgs a{};
  a.params = params; a.grads = grads; a.running = bn_running; a.x = x; a.labels = labels; a.keep_in = nullptr;
  a.dlogits_in = nullptr; a.logits = logits; a.loss = loss; a.correct = correct;
  a.h1 = w.h1; a.h2 = w.h2; a.d1 = w.d1; a.d2 = w.d2; a.dlog = w.dlog; a.keep = w.keep; a.bnc = w.bnc;
  a.B = batch; a.D = input_dim; a.C = num_classes; a.seed = dropout_seed; a.seed_dev = (unsigned long long*)seed_dev; a.nbt = bn_steps;
  a.p = dropout_p;
  a.flags = MLP_FWD | MLP_TRAIN | MLP_CE | MLP_BWD;
  AE_TRY(mlp_launch(a, (cudaStream_t)stream));
  int64_t off[10];
  const int64_t n = mlp_layout(input_dim, num_classes, off, nullptr);
  return adam_step_flat(params, grads, adam_m, adam_v, n, adam->lr, adam->beta1, adam->beta2, adam->eps, adam->weight_decay, 1.f,
                        step_dev, (cudaStream_t)stream);
}

// The same step reading its batch through an index: rows order[cursor[0] ..] of x_all / labels_all (cursor = int64[2]: rows
// consumed, steps done); (loss, correct) of step i land in hist[i].  One graph replay per batch runs an epoch.
int ae_mlp_train_step_indexed(float* params, float* grads, float* bn_running, int64_t* bn_steps, const float* x_all,
                              const int64_t* labels_all, const int64_t* order, int64_t* cursor, float* hist, uint64_t dropout_seed,
                              uint64_t* seed_dev, float dropout_p, int batch, int input_dim, int num_classes, float* logits,
                              float* loss, int* correct, void* workspace, size_t workspace_bytes, const ae_adam_config_t* adam,
                              float* adam_m, float* adam_v, int* step_dev, ae_stream_t stream) {
  AE_CHECK(params && grads && bn_running && x_all && labels_all && order && cursor && seed_dev && workspace && adam && adam_m &&
               adam_v && step_dev && loss && correct,
           "ae_mlp_train_step_indexed: null argument");
  AE_CHECK(workspace_bytes >= mlp_carve(batch, nullptr, nullptr), "ae_mlp_train_step_indexed: workspace too small");
  AE_CHECK(dropout_p >= 0.f && dropout_p < 1.f, "ae_mlp_train_step_indexed: dropout_p out of range");
  MlpWs w;
  mlp_carve(batch, (char*)workspace, &w);
  MlpArgs a{};
  a.params = params; a.grads = grads; a.running = bn_running; a.x = x_all; a.labels = labels_all; a.keep_in = nullptr;
  a.dlogits_in = nullptr; a.logits = logits; a.loss = loss; a.correct = correct;
  a.h1 = w.h1; a.h2 = w.h2; a.d1 = w.d1; a.d2 = w.d2; a.dlog = w.dlog; a.keep = w.keep; a.bnc = w.bnc;
  a.B = batch; a.D = input_dim; a.C = num_classes; a.seed = dropout_seed; a.seed_dev = (unsigned long long*)seed_dev; a.nbt = bn_steps;
  a.order = (const long long*)order; a.cursor = (long long*)cursor; a.hist = hist;
  a.p = dropout_p;
  a.flags = MLP_FWD | MLP_TRAIN | MLP_CE | MLP_BWD;
  AE_TRY(mlp_launch(a, (cudaStream_t)stream));
  int64_t off[10];
  const int64_t n = mlp_layout(input_dim, num_classes, off, nullptr);
  return adam_step_flat(params, grads, adam_m, adam_v, n, adam->lr, adam->beta1, adam->beta2, adam->eps, adam->weight_decay, 1.f,
                        step_dev, (cudaStream_t)stream);
}

// backward of a preceding training-mode forward-only call (labels == NULL, same workspace), given d(loss)/d(logits)
int ae_mlp_backward(const float* params, float* grads, const float* x, const float* d_logits, float dropout_p, int batch,
                    int input_dim, int num_classes, void* workspace, size_t workspace_bytes, ae_stream_t stream) {
  AE_CHECK(params && grads && x && d_logits && workspace, "ae_mlp_backward: null argument");
  AE_CHECK(workspace_bytes >= mlp_carve(batch, nullptr, nullptr), "ae_mlp_backward: workspace too small");
  MlpWs w;
  mlp_carve(batch, (char*)workspace, &w);
  MlpArgs a{};
  a.params = params; a.grads = grads; a.running = nullptr; a.x = x; a.labels = nullptr; a.keep_in = nullptr;
  a.dlogits_in = d_logits; a.logits = nullptr; a.loss = nullptr; a.correct = nullptr;
  a.h1 = w.h1; a.h2 = w.h2; a.d1 = w.d1; a.d2 = w.d2; a.dlog = w.dlog; a.keep = w.keep; a.bnc = w.bnc;
  a.B = batch; a.D = input_dim; a.C = num_classes; a.seed = 0; a.p = dropout_p;
  a.flags = MLP_BWD | MLP_TRAIN;
  return mlp_launch(a, (cudaStream_t)stream);
}

}  // extern "C"

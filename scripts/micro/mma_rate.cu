// Issue rate of the legacy warp-level tensor path (mma.sync m16n8k8 tf32, m16n8k16 bf16) against FFMA / FFMA2 on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float d[8][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x + 5, threadIdx.x + 7};
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  float2 f[16];
  for (int i = 0; i < 16; ++i) f[i] = make_float2(0.f, 0.f);
  const float x = __uint_as_float(0x3f800000 + threadIdx.x), y = 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_tf32(d[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_bf16(d[i], a, b);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { f[i].x = fmaf(x, y, f[i].x); f[i].y = fmaf(y, x, f[i].y); }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __ffma2_rn(make_float2(x, x), make_float2(y, x), f[i]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  for (int i = 0; i < 16; ++i) s += f[i].x + f[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

int main() {
  float* out; long long* clk; long long h;
  cudaMalloc(&out, 148 * 16 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  const char* names[4] = {"mma.sync m16n8k8 tf32", "mma.sync m16n8k16 bf16", "FFMA", "FFMA2"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, clk);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      const double per_sm_inst = (double)warps * iters * (mode < 2 ? 8 : (mode == 2 ? 32 : 16));
      const double mac = mode == 0 ? 16 * 8 * 8 : mode == 1 ? 16 * 8 * 16 : mode == 2 ? 32 : 64;
      printf("%-24s %2d warps/SM: %8.3f warp-instr/clk/SM  %8.1f MAC/clk/SM\n", names[mode], warps, per_sm_inst / h, per_sm_inst * mac / h);
    }
  }
  return 0;
}

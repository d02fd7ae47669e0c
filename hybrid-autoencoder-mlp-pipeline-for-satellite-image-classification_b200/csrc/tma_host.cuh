// Host-side tensor-map helpers shared by the TMA-fed kernels (tma_gemm.cu, rowgemm2.cu).
#pragma once
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace ae {

// ---------------------------------------------------------------------------------------------
// tensor maps (host)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// Lattice of an NHWC split-bf16 tensor: pixels (y0 + sy*j, x0 + sx*i), j < nH, i < nW, of a [B][H][W][C] image.
// Box = (kc channels, bx, by, bn, nsplit).
inline int encode_map(CUtensorMap* map, const void* planes, int B, int H, int W, int C, int nsplit, int y0, int x0,
                      int sy, int sx, int nH, int nW, int kc, int bx, int by, int bn) {
  EncodeTiledFn fn = encode_fn();
  AE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const char* base = static_cast<const char*>(planes) + ((size_t)y0 * W + x0) * C * 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)nW, (cuuint64_t)nH, (cuuint64_t)B, (cuuint64_t)nsplit};
  cuuint64_t strides[4] = {(cuuint64_t)sx * C * 2, (cuuint64_t)sy * W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)B * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn, (cuuint32_t)nsplit};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  AE_CHECK(kc * 2 == 128 || kc * 2 == 64, "encode_map: box of %d channels is neither 64 nor 128 bytes", kc);
  AE_CHECK(((uintptr_t)base & 15) == 0, "encode_map: tensor base must be 16-byte aligned");
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<char*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (C=%d W=%d H=%d B=%d box %d,%d,%d,%d)", (int)r, C,
           nW, nH, B, kc, bx, by, bn);
  return 0;
}

// pixel box of `rows` consecutive small pixels (row-major over n, y, x): bx * by * bn == rows
inline void pixel_box(int Hs, int Ws, int rows, int* bx, int* by, int* bn) {
  *bx = Ws < rows ? Ws : rows;
  int r = rows / *bx;
  *by = Hs < r ? Hs : r;
  *bn = r / *by;
}


// fp32 NHWC tensor [B][H][W][C] seen as the pixel lattice (y0 + sy*j, x0 + sx*i), j < nH, i < nW: the form in which the
// second-generation row GEMM stores its output tiles and loads the forward activations of its ReLU-backward epilogue.
// Box = (32 channels = 128 bytes, bx, by, bn), SWIZZLE_128B.
inline int encode_map_f32(CUtensorMap* map, const float* tensor, int B, int H, int W, int C, int y0, int x0, int sy, int sx,
                          int nH, int nW, int bx, int by, int bn) {
  EncodeTiledFn fn = encode_fn();
  AE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  AE_CHECK(C % 32 == 0, "encode_map_f32: %d channels (must be a multiple of 32)", C);
  const char* base = reinterpret_cast<const char*>(tensor) + ((size_t)y0 * W + x0) * C * 4;
  AE_CHECK(((uintptr_t)base & 15) == 0, "encode_map_f32: tensor base must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)nW, (cuuint64_t)nH, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sx * C * 4, (cuuint64_t)sy * W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<char*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d (C=%d W=%d H=%d B=%d box %d,%d,%d)", (int)r, C, nW,
           nH, B, bx, by, bn);
  return 0;
}

}  // namespace ae

// Device-side input transforms of the reference: the training pipeline NB:386-391
//   RandomHorizontalFlip -> RandomCrop(64, padding=4) -> ToTensor -> AddGaussianNoise(0, 0.03)   (class NB:361-368)
// and the eval pipeline NB:393-395 (ToTensor), over uint8 HWC images that stay resident in HBM.  One pass: gather the
// batch by index (the shuffled DataLoader batch, NB:420), flip / shift with zero fill, /255, add noise, write the fp32
// NCHW tensor the encoder reads.  HBM-bound: 12 KB read + 48 KB written per image (a fourth of the bytes an fp32 batch
// costs to bring over PCIe / NVLink).  The random draws (flip bit, crop offsets) are inputs; the noise is either an
// input (parity tests) or drawn here from Philox4x32-10 keyed by (seed, element index).
#include "common.cuh"

namespace ae {

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __device__ static uint4 run(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
      const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
      ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
      key.x += W0; key.y += W1;
    }
    return ctr;
  }
};

// four standard normals from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t block) {
  const uint4 r = Philox::run(make_uint4((uint32_t)block, (uint32_t)(block >> 32), 0u, 0u),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;   // 2^-32
  const float u0 = fmaf((float)r.x, k, 0.5f * k), u1 = (float)r.y * k;
  const float u2 = fmaf((float)r.z, k, 0.5f * k), u3 = (float)r.w * k;
  const float a = sqrtf(-2.f * logf(u0)), b = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  return make_float4(a * c0, a * s0, b * c1, b * s1);
}

static constexpr int AH = 64, AW = 64;

// one thread: 4 consecutive output columns of one row, all three channels
__global__ void __launch_bounds__(256) k_augment_u8(const uint8_t* __restrict__ src, const int64_t* __restrict__ index,
                                                   const uint8_t* __restrict__ flip, const int32_t* __restrict__ off_y,
                                                   const int32_t* __restrict__ off_x, int pad,
                                                   const float* __restrict__ noise, uint64_t seed, float mean, float stdv,
                                                   float* __restrict__ out, int batch) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)batch * AH * (AW / 4)) return;
  const int x0 = (int)(t % (AW / 4)) * 4;
  const int y = (int)((t / (AW / 4)) % AH);
  const int b = (int)(t / (AH * (AW / 4)));
  const int64_t img = index ? index[b] : b;
  const bool fl = flip && flip[b] != 0;
  const int sy = y + (off_y ? off_y[b] : pad) - pad;           // row of the (flipped) source image
  const int dx = (off_x ? off_x[b] : pad) - pad;
  const uint8_t* row = src + (img * AH + sy) * (int64_t)(AW * 3);
  float v[3][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int fx = x0 + k + dx;                                // column of the flipped image
    const bool inb = sy >= 0 && sy < AH && fx >= 0 && fx < AW;
    const int sx = fl ? AW - 1 - fx : fx;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c][k] = inb ? __fdiv_rn((float)__ldg(row + sx * 3 + c), 255.f) : 0.f;   // ToTensor: .div(255)
  }
  const bool noisy = noise != nullptr || stdv != 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int64_t o = (((int64_t)b * 3 + c) * AH + y) * AW + x0;
    float4 r = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
    if (noisy) {
      const float4 n = noise ? __ldg(reinterpret_cast<const float4*>(noise + o)) : normal4(seed, (uint64_t)(o >> 2));
      // tensor + (noise * std + mean), each operation rounded on its own as torch does (NB:367-368)
      r.x = __fadd_rn(r.x, __fadd_rn(__fmul_rn(n.x, stdv), mean));
      r.y = __fadd_rn(r.y, __fadd_rn(__fmul_rn(n.y, stdv), mean));
      r.z = __fadd_rn(r.z, __fadd_rn(__fmul_rn(n.z, stdv), mean));
      r.w = __fadd_rn(r.w, __fadd_rn(__fmul_rn(n.w, stdv), mean));
    }
    *reinterpret_cast<float4*>(out + o) = r;
  }
}

}  // namespace ae

using namespace ae;

extern "C" int ae_augment_u8(const uint8_t* images, int64_t num_images, const int64_t* index, const uint8_t* flip,
                             const int32_t* off_y, const int32_t* off_x, int pad, const float* noise, uint64_t seed,
                             float noise_mean, float noise_std, float* out, int batch, ae_stream_t stream) {
  AE_CHECK(images && out && batch >= 1, "ae_augment_u8: bad argument");
  AE_CHECK(index != nullptr || batch <= num_images, "ae_augment_u8: batch %d exceeds the %lld images given and no index", batch,
           (long long)num_images);
  AE_CHECK(pad >= 0 && pad <= 32, "ae_augment_u8: pad=%d out of range", pad);
  AE_CHECK((((uintptr_t)out | (uintptr_t)noise) & 15) == 0, "ae_augment_u8: out / noise must be 16-byte aligned");
  const int64_t threads = (int64_t)batch * AH * (AW / 4);
  k_augment_u8<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(images, index, flip, off_y, off_x, pad, noise,
                                                                                   seed, noise_mean, noise_std, out, batch);
  AE_LAUNCH_CHECK();
  return 0;
}

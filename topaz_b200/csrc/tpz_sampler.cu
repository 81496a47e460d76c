// GPU training-crop sampler + augmentation.  Replaces the host data loader of `topaz train`:
// MultipleImageSetDataset.__getitem__ / MemoryMappedImage.get_crop / get_UN_crop (topaz/utils/data/
// memory_mapped_data.py:45-100, 195-233): positive-balanced sampling, 'pn' rejection of labelled pixels for negatives,
// zero-padded crops of ceil(crop*sqrt(2)) pixels, nearest-neighbour rotation about the crop centre by a uniform angle
// (torchvision.transforms.functional.rotate defaults), centre crop, random horizontal / vertical flips.
// Micrographs, the positive-pixel table and a positive mask stay resident in HBM; one launch draws the minibatch
// parameters (Philox counter RNG: reproducible from (seed, batch index)), one launch writes X [B][crop][crop] fp32.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <curand_kernel.h>

namespace {

struct SampleParams {
  int img, cy, cx, label;
  float angle_deg;
  int hflip, vflip, pad;
};

__global__ void sample_params_kernel(int B, unsigned long long seed, unsigned long long batch_index,
                                     const TpzSamplerImage* __restrict__ imgs, const int* __restrict__ set_begin,
                                     const float* __restrict__ set_cdf, int nsets, const int* __restrict__ positives /*[P][3]*/,
                                     int P, const unsigned char* __restrict__ pos_mask, float positive_balance, int split_pn,
                                     int rotate, int flip, SampleParams* __restrict__ out, double* __restrict__ labels) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, batch_index * (unsigned long long)B + b, 0, &st);
  SampleParams s;
  // image set first (the reference draws it even when a positive is sampled)
  const float us = curand_uniform(&st);
  int set = 0;
  while (set + 1 < nsets && us > set_cdf[set]) ++set;
  if (P > 0 && curand_uniform(&st) <= positive_balance) {
    // 32-bit draw + multiply-high: the positives table lists every labelled PIXEL (discs are expanded) and can exceed 2^24
    // entries, beyond which a float draw can no longer reach every index
    const int k = (int)__umulhi(curand(&st), (unsigned)P);
    s.img = positives[3 * k]; s.cy = positives[3 * k + 1]; s.cx = positives[3 * k + 2]; s.label = 1;
  } else {
    const int n_in_set = set_begin[set + 1] - set_begin[set];
    s.img = set_begin[set] + min((int)(curand_uniform(&st) * n_in_set), n_in_set - 1);
    const TpzSamplerImage im = imgs[s.img];
    for (int tries = 0; tries < 10000; ++tries) {
      s.cx = min((int)(curand_uniform(&st) * im.W), im.W - 1);
      s.cy = min((int)(curand_uniform(&st) * im.H), im.H - 1);
      if (!split_pn || pos_mask == nullptr || pos_mask[im.offset + (long long)s.cy * im.W + s.cx] == 0) break;
    }
    s.label = 0;
  }
  s.angle_deg = rotate ? 360.f * (1.f - curand_uniform(&st)) : 0.f;     // uniform [0, 360)
  s.hflip = (flip && curand_uniform(&st) > 0.5f) ? 1 : 0;
  s.vflip = (flip && curand_uniform(&st) > 0.5f) ? 1 : 0;
  s.pad = 0;
  out[b] = s;
  labels[b] = (double)s.label;
}

// X[b][i][j]: undo flips -> position in the rotated big crop -> inverse rotation (nearest, torchvision's affine grid
// convention: centred coordinates, round-half-even) -> big-crop pixel -> micrograph pixel (zero outside)
__global__ void make_crops_kernel(int B, int crop, int big, const TpzSamplerImage* __restrict__ imgs,
                                  const float* __restrict__ pixels, const SampleParams* __restrict__ params,
                                  float* __restrict__ X) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= crop * crop) return;
  const SampleParams s = params[b];
  int i = t / crop, j = t - i * crop;
  if (s.vflip) i = crop - 1 - i;        // vflip was applied last in the forward pipeline -> undone first
  if (s.hflip) j = crop - 1 - j;
  const int off = (big - crop) / 2;
  const int bi = i + off, bj = j + off;   // position in the rotated big crop
  float sy = (float)bi, sx = (float)bj;
  if (s.angle_deg != 0.f) {
    const float rot = -s.angle_deg * 0.017453292519943295f;
    const float ca = cosf(rot), sa = sinf(rot);
    const float x = (float)bj - 0.5f * (big - 1), y = (float)bi - 0.5f * (big - 1);
    // inverse matrix of torchvision._get_inverse_affine_matrix(center 0, angle=-a): [cos, sin; -sin, cos]
    sx = ca * x + sa * y + 0.5f * (big - 1);
    sy = -sa * x + ca * y + 0.5f * (big - 1);
  }
  const int rx = __float2int_rn(sx), ry = __float2int_rn(sy);
  float v = 0.f;
  if (rx >= 0 && rx < big && ry >= 0 && ry < big) {
    const TpzSamplerImage im = imgs[s.img];
    const int gy = s.cy - big / 2 + ry, gx = s.cx - big / 2 + rx;
    if (gy >= 0 && gy < im.H && gx >= 0 && gx < im.W) v = pixels[im.offset + (long long)gy * im.W + gx];
  }
  X[((long long)b * crop + (t / crop)) * crop + (t % crop)] = v;
}

}  // namespace

extern "C" int tpz_sample_crops(int B, unsigned long long seed, unsigned long long batch_index, const TpzSamplerImage* imgs,
                                const float* pixels, const int* set_begin, const float* set_cdf, int nsets,
                                const int* positives, int num_positives, const unsigned char* pos_mask,
                                float positive_balance, int split_pn, int rotate, int flip, int crop, int big_crop,
                                void* params_scratch, float* X, double* Y, void* stream) {
  TPZ_CHECK(B > 0 && crop > 0 && big_crop >= crop && (big_crop - crop) % 2 == 0, "tpz_sample_crops: bad sizes");
  TPZ_CHECK(nsets > 0, "tpz_sample_crops: no image sets");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SampleParams* p = reinterpret_cast<SampleParams*>(params_scratch);
  sample_params_kernel<<<tpz_div_up(B, 128), 128, 0, st>>>(B, seed, batch_index, imgs, set_begin, set_cdf, nsets, positives,
                                                           num_positives, pos_mask, positive_balance, split_pn, rotate, flip, p, Y);
  dim3 grid(tpz_div_up(crop * crop, 256), B);
  make_crops_kernel<<<grid, 256, 0, st>>>(B, crop, big_crop, imgs, pixels, p, X);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// deterministic variant used by the parity tests: the caller supplies the per-sample parameters
extern "C" int tpz_make_crops(int B, int crop, int big_crop, const TpzSamplerImage* imgs, const float* pixels,
                              const void* params /* B x {img,cy,cx,label,angle,hflip,vflip,pad} */, float* X, void* stream) {
  dim3 grid(tpz_div_up(crop * crop, 256), B);
  make_crops_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(B, crop, big_crop, imgs, pixels,
                                                                            reinterpret_cast<const SampleParams*>(params), X);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// Training-mode BatchNorm of the strided classifier (`topaz train` builds its ResNets with --bn on by default:
// reference topaz/commands/train.py:91, modules at topaz/model/features/resnet.py:68-70,134-141, called at
// resnet.py:101-104 and :185-204).  NHWC fp32 activations [P][C]; per-channel statistics are accumulated in fp64
// (sum, sum of squares) so that var = E[x^2] - E[x]^2 carries no cancellation error at fp32 level, and so that the
// multi-GPU path can all-reduce the raw sums (statistics of the GLOBAL minibatch, as in the single-process reference).
//
//   tpz_bn_stats_f32      sums[c] += sum_p x[p][c],  sums[C+c] += sum_p x[p][c]^2
//   tpz_bn_fwd_f32        y = relu?((x-mean)*invstd*gamma + beta); writes save = {mean, invstd}; running-stat update
//   tpz_bn_bwd_reduce_f32 sums[c] += sum_p g[p][c],  sums[C+c] += sum_p g[p][c]*xhat[p][c]
//   tpz_bn_bwd_f32        dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat));  dgamma += sum g*xhat, dbeta += sum g
// All four are HBM-bound single passes (4-8 B per element read, 4 B written).
//
// The other element-wise layers of the training nets live here too:
//   tpz_act_fwd_f32 / tpz_act_bwd_f32          PReLU (one learnable slope) / LeakyReLU of conv31/63/127 (basic.py:16,51,66)
//   tpz_dropout_fwd_f32 / tpz_dropout_bwd_f32  nn.Dropout in training (resnet.py:296-303): Philox keep-masks
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <curand_kernel.h>

namespace {

constexpr int BN_MAX_C = 2048;

// Block = 32 channels x 8 row-slices (same decomposition as bias_grad_kernel); WITH_G: reduce g and g*xhat instead.
template <bool WITH_G>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ save, long long P, int C,
                                                        double* __restrict__ sums) {
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, ss = 0.0;
  if (c < C) {
    float mean = 0.f, invstd = 1.f;
    if (WITH_G) { mean = save[c]; invstd = save[C + c]; }
    const long long step = (long long)gridDim.y * 8;
    long long p = (long long)blockIdx.y * 8 + slice;
    for (; p + 3 * step < P; p += 4 * step) {
      const float x0 = x[p * C + c], x1 = x[(p + step) * C + c], x2 = x[(p + 2 * step) * C + c], x3 = x[(p + 3 * step) * C + c];
      if (WITH_G) {
        const float g0 = g[p * C + c], g1 = g[(p + step) * C + c], g2 = g[(p + 2 * step) * C + c], g3 = g[(p + 3 * step) * C + c];
        s += ((double)g0 + (double)g1) + ((double)g2 + (double)g3);
        ss += ((double)g0 * (double)((x0 - mean) * invstd) + (double)g1 * (double)((x1 - mean) * invstd)) +
              ((double)g2 * (double)((x2 - mean) * invstd) + (double)g3 * (double)((x3 - mean) * invstd));
      } else {
        s += ((double)x0 + (double)x1) + ((double)x2 + (double)x3);
        ss += ((double)x0 * x0 + (double)x1 * x1) + ((double)x2 * x2 + (double)x3 * x3);
      }
    }
    for (; p < P; p += step) {
      const float x0 = x[p * C + c];
      if (WITH_G) {
        const float g0 = g[p * C + c];
        s += (double)g0;
        ss += (double)g0 * (double)((x0 - mean) * invstd);
      } else {
        s += (double)x0;
        ss += (double)x0 * x0;
      }
    }
  }
  __shared__ double sh[2][8][33];
  sh[0][slice][lane] = s;
  sh[1][slice][lane] = ss;
  __syncthreads();
  if (slice == 0 && c < C) {
    for (int k = 1; k < 8; ++k) { s += sh[0][k][lane]; ss += sh[1][k][lane]; }
    atomicAdd(&sums[c], s);
    atomicAdd(&sums[C + c], ss);
  }
}

// y = act((x - mean) * (invstd*gamma) + beta).  sums != nullptr: training mode, statistics from the sums over `count`
// elements per channel (block 0 also writes save = {mean, invstd} and updates the running statistics);
// sums == nullptr: mean / invstd are read from `save` (eval mode, running statistics prepared by the caller).
__global__ void __launch_bounds__(256) bn_fwd_kernel(const float* __restrict__ x, long long P, int C,
                                                     const double* __restrict__ sums, double inv_count, double unbias,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, float momentum, float* __restrict__ running_mean,
                                                     float* __restrict__ running_var, int relu, float* __restrict__ y,
                                                     float* save) {
  extern __shared__ float s_bn[];            // mean[C], scale[C], shift[C]
  float* s_mean = s_bn; float* s_scale = s_bn + C; float* s_shift = s_bn + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    if (sums != nullptr) {
      const double m = sums[c] * inv_count;
      double var = sums[C + c] * inv_count - m * m;
      if (var < 0.0) var = 0.0;
      const double is = 1.0 / sqrt(var + (double)eps);
      mean = (float)m; invstd = (float)is;
      if (blockIdx.x == 0) {
        save[c] = mean; save[C + c] = invstd;
        if (running_mean != nullptr) {
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
        }
      }
    } else {
      mean = save[c]; invstd = save[C + c];
    }
    s_mean[c] = mean;
    s_scale[c] = invstd * (gamma != nullptr ? gamma[c] : 1.f);
    s_shift[c] = beta != nullptr ? beta[c] : 0.f;
  }
  __syncthreads();
  const long long total = P * C;
  if ((C & 3) == 0) {
    const long long total4 = total >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)((i << 2) % C);
      float4 v = x4[i];
      v.x = (v.x - s_mean[c]) * s_scale[c] + s_shift[c];
      v.y = (v.y - s_mean[c + 1]) * s_scale[c + 1] + s_shift[c + 1];
      v.z = (v.z - s_mean[c + 2]) * s_scale[c + 2] + s_shift[c + 2];
      v.w = (v.w - s_mean[c + 3]) * s_scale[c + 3] + s_shift[c + 3];
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      y4[i] = v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % C);
      float v = (x[i] - s_mean[c]) * s_scale[c] + s_shift[c];
      if (relu) v = fmaxf(v, 0.f);
      y[i] = v;
    }
  }
}

// dx = gamma*invstd*(g - a - xhat*b), a = sum(g)/count, b = sum(g*xhat)/count (sums over the GLOBAL minibatch);
// block 0 accumulates this rank's parameter gradients from its LOCAL sums.  dx may alias g.
__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* g, const float* __restrict__ x, long long P, int C,
                                                     const float* __restrict__ save, const double* __restrict__ sums,
                                                     double inv_count, const float* __restrict__ gamma,
                                                     const double* __restrict__ local_sums, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, float* dx) {
  extern __shared__ float s_bn[];            // mean[C], invstd[C], scale[C], a[C], b[C]
  float* s_mean = s_bn; float* s_inv = s_bn + C; float* s_scale = s_bn + 2 * C; float* s_a = s_bn + 3 * C;
  float* s_b = s_bn + 4 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float invstd = save[C + c];
    s_mean[c] = save[c];
    s_inv[c] = invstd;
    s_scale[c] = invstd * (gamma != nullptr ? gamma[c] : 1.f);
    s_a[c] = (float)(sums[c] * inv_count);
    s_b[c] = (float)(sums[C + c] * inv_count);
    if (blockIdx.x == 0) {
      if (dbeta != nullptr) dbeta[c] += (float)local_sums[c];
      if (dgamma != nullptr) dgamma[c] += (float)local_sums[C + c];
    }
  }
  __syncthreads();
  const long long total = P * C;
  if ((C & 3) == 0) {
    const long long total4 = total >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* d4 = reinterpret_cast<float4*>(dx);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)((i << 2) % C);
      const float4 gv = g4[i];
      const float4 xv = x4[i];
      float4 o;
      o.x = s_scale[c] * (gv.x - s_a[c] - (xv.x - s_mean[c]) * s_inv[c] * s_b[c]);
      o.y = s_scale[c + 1] * (gv.y - s_a[c + 1] - (xv.y - s_mean[c + 1]) * s_inv[c + 1] * s_b[c + 1]);
      o.z = s_scale[c + 2] * (gv.z - s_a[c + 2] - (xv.z - s_mean[c + 2]) * s_inv[c + 2] * s_b[c + 2]);
      o.w = s_scale[c + 3] * (gv.w - s_a[c + 3] - (xv.w - s_mean[c + 3]) * s_inv[c + 3] * s_b[c + 3]);
      d4[i] = o;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % C);
      dx[i] = s_scale[c] * (g[i] - s_a[c] - (x[i] - s_mean[c]) * s_inv[c] * s_b[c]);
    }
  }
}

// PReLU with one learnable slope / LeakyReLU (the activation of the conv31/63/127 extractors: reference
// topaz/model/features/basic.py:16,51,66):  y = v > 0 ? v : a*v.   a = *slope_dev when given (nn.PReLU().weight), else
// slope_const (nn.LeakyReLU).
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ v, long long n, const float* __restrict__ slope_dev,
                                                      float slope_const, float* __restrict__ y) {
  const float a = slope_dev != nullptr ? slope_dev[0] : slope_const;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = v[i];
    y[i] = x > 0.f ? x : a * x;
  }
}

// g <- g * (v > 0 ? 1 : a);  dslope += sum over v <= 0 of g*v (torch's PReLU backward; one atomicAdd per block)
__global__ void __launch_bounds__(256) act_bwd_kernel(float* g, const float* __restrict__ v, long long n,
                                                      const float* __restrict__ slope_dev, float slope_const,
                                                      float* __restrict__ dslope) {
  const float a = slope_dev != nullptr ? slope_dev[0] : slope_const;
  double part = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = v[i], gi = g[i];
    if (!(x > 0.f)) {
      part += (double)gi * (double)x;
      g[i] = a * gi;
    }
  }
  if (dslope != nullptr) {          // uniform branch (kernel argument)
    __shared__ double sh[8];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += sh[w];
      atomicAdd(dslope, (float)tot);
    }
  }
}

// nn.Dropout in training (reference resnet.py:296-303, basic.py:58-59,71-72): keep each element with probability 1-p and
// scale the kept ones by 1/(1-p).  Philox counter RNG: element group q = i/4 draws from subsequence q at `offset`, so the
// mask is a pure function of (seed, offset, i) -- reproducible under torch.manual_seed, independent of the launch shape.
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const float* __restrict__ x, long long n, float p, float scale,
                                                          unsigned long long seed, unsigned long long offset,
                                                          float* __restrict__ y, unsigned char* __restrict__ mask) {
  const long long groups = (n + 3) >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += (long long)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)q, offset, &st);
    const float4 r4 = curand_uniform4(&st);          // uniform in (0, 1]
    const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = (q << 2) + e;
      if (i < n) {
        const bool keep = r[e] > p;
        y[i] = keep ? x[i] * scale : 0.f;
        mask[i] = keep ? 1 : 0;
      }
    }
  }
}

// g <- g * mask * scale
__global__ void __launch_bounds__(256) dropout_bwd_kernel(float* __restrict__ g, const unsigned char* __restrict__ mask,
                                                          long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] = mask[i] ? g[i] * scale : 0.f;
}

inline dim3 reduce_grid(long long P, int C) {
  return dim3(tpz_div_up(C, 32), (unsigned)(P < 4096 ? 1 : (P < 65536 ? 64 : 296)));
}

inline int elementwise_grid(long long total) {
  int grid = tpz_div_up(total, 256 * 8);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  return grid;
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tpz_bn_stats_f32(const float* x, long long P, int C, double* sums, void* stream) {
  TPZ_CHECK(P > 0 && C > 0 && C <= BN_MAX_C, "tpz_bn_stats_f32: bad shape P=%lld C=%d", P, C);
  bn_reduce_kernel<false><<<reduce_grid(P, C), 256, 0, ST(stream)>>>(x, nullptr, nullptr, P, C, sums);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_bn_fwd_f32(const float* x, long long P, int C, const double* sums, long long count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              int relu, float* y, float* save, void* stream) {
  TPZ_CHECK(P > 0 && C > 0 && C <= BN_MAX_C, "tpz_bn_fwd_f32: bad shape P=%lld C=%d", P, C);
  TPZ_CHECK(save != nullptr, "tpz_bn_fwd_f32: save (mean, invstd) buffer is required");
  TPZ_CHECK(sums == nullptr || count > 0, "tpz_bn_fwd_f32: count must be positive in training mode");
  TPZ_CHECK((running_mean == nullptr) == (running_var == nullptr), "tpz_bn_fwd_f32: running_mean/var must come together");
  const double inv_count = sums != nullptr ? 1.0 / (double)count : 0.0;
  const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  bn_fwd_kernel<<<elementwise_grid(P * C), 256, 3 * C * sizeof(float), ST(stream)>>>(
      x, P, C, sums, inv_count, unbias, gamma, beta, eps, momentum, running_mean, running_var, relu, y, save);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_bn_bwd_reduce_f32(const float* g, const float* x, long long P, int C, const float* save, double* sums,
                                     void* stream) {
  TPZ_CHECK(P > 0 && C > 0 && C <= BN_MAX_C, "tpz_bn_bwd_reduce_f32: bad shape P=%lld C=%d", P, C);
  bn_reduce_kernel<true><<<reduce_grid(P, C), 256, 0, ST(stream)>>>(x, g, save, P, C, sums);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_bn_bwd_f32(const float* g, const float* x, long long P, int C, const float* save, const double* sums,
                              long long count, const float* gamma, const double* local_sums, float* dgamma, float* dbeta,
                              float* dx, void* stream) {
  TPZ_CHECK(P > 0 && C > 0 && C <= BN_MAX_C, "tpz_bn_bwd_f32: bad shape P=%lld C=%d", P, C);
  TPZ_CHECK(count > 0, "tpz_bn_bwd_f32: count must be positive");
  TPZ_CHECK((dgamma == nullptr && dbeta == nullptr) || local_sums != nullptr, "tpz_bn_bwd_f32: local_sums required");
  bn_bwd_kernel<<<elementwise_grid(P * C), 256, 5 * C * sizeof(float), ST(stream)>>>(
      g, x, P, C, save, sums, 1.0 / (double)count, gamma, local_sums, dgamma, dbeta, dx);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_act_fwd_f32(const float* v, long long n, const float* slope_dev, float slope_const, float* y, void* stream) {
  TPZ_CHECK(n > 0, "tpz_act_fwd_f32: empty tensor");
  act_fwd_kernel<<<elementwise_grid(n), 256, 0, ST(stream)>>>(v, n, slope_dev, slope_const, y);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_act_bwd_f32(float* g, const float* v, long long n, const float* slope_dev, float slope_const,
                               float* dslope, void* stream) {
  TPZ_CHECK(n > 0, "tpz_act_bwd_f32: empty tensor");
  act_bwd_kernel<<<elementwise_grid(n), 256, 0, ST(stream)>>>(g, v, n, slope_dev, slope_const, dslope);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_dropout_fwd_f32(const float* x, long long n, float p, unsigned long long seed, unsigned long long offset,
                                   float* y, unsigned char* mask, void* stream) {
  TPZ_CHECK(n > 0 && p >= 0.f && p < 1.f, "tpz_dropout_fwd_f32: bad arguments n=%lld p=%g", n, (double)p);
  dropout_fwd_kernel<<<elementwise_grid(n), 256, 0, ST(stream)>>>(x, n, p, 1.f / (1.f - p), seed, offset, y, mask);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_dropout_bwd_f32(float* g, const unsigned char* mask, long long n, float p, void* stream) {
  TPZ_CHECK(n > 0 && p >= 0.f && p < 1.f, "tpz_dropout_bwd_f32: bad arguments n=%lld p=%g", n, (double)p);
  dropout_bwd_kernel<<<elementwise_grid(n), 256, 0, ST(stream)>>>(g, mask, n, 1.f / (1.f - p));
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// Micrograph preprocessing on the GPU: the reductions of the 2-component GMM normalisation
// (topaz/stats.py:86-214) and exact order statistics for its quantile initialisation (stats.py:91, np.quantile).
// Both are HBM-bound single passes over the image: 4 B/px read, fp64 block reductions, one atomic per block per sum.
// The Fourier-crop downsample (utils/image.py:38-61) is two dense products with precomputed real matrices and runs on
// tpz_gemm_f32 (tpz_train_mma.cu).
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums[0..6] += { sum Z, sum p0, sum p1, sum p0*xc, sum p1*xc, sum p0*xc^2, sum p1*xc^2 },  xc = x - shift
// mode 0 (initial hard split, stats.py:136-139): p0 = (x <= split), p1 = 1 - p0, Z = 0
// mode 1 (E step, stats.py:158-167 / 172-203): log_pk = -(x-mu_k)^2/2/var_k - 0.5 log(2 pi var_k) + log prior_k,
//         Z = logsumexp, p_k = exp(log_pk - Z).  Parameters arrive in shifted coordinates (mu_k - shift).
struct GmmParams { double shift, split, mu0, mu1, var0, var1, log_prior0, log_prior1; };

__global__ void __launch_bounds__(256) gmm_sums_kernel(const float* __restrict__ x, long long n, int mode, GmmParams P,
                                                       double* __restrict__ sums) {
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  const double c0 = -0.5 * log(2.0 * 3.14159265358979323846 * P.var0) + P.log_prior0;
  const double c1 = -0.5 * log(2.0 * 3.14159265358979323846 * P.var1) + P.log_prior1;
  const double h0 = 0.5 / P.var0, h1 = 0.5 / P.var1;
  const float splitf = (float)P.split;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xf = x[i];
    const double xc = (double)xf - P.shift;
    double p0, p1, Z = 0.0;
    if (mode == 0) {
      p0 = xf <= splitf ? 1.0 : 0.0;
      p1 = 1.0 - p0;
    } else {
      const double d0 = xc - P.mu0, d1 = xc - P.mu1;
      const double l0 = c0 - d0 * d0 * h0, l1 = c1 - d1 * d1 * h1;
      const double ma = fmax(l0, l1);
      const double e0 = exp(l0 - ma), e1 = exp(l1 - ma);
      const double s = e0 + e1;
      Z = ma + log(s);
      p0 = e0 / s;
      p1 = e1 / s;
    }
    a[0] += Z; a[1] += p0; a[2] += p1; a[3] += p0 * xc; a[4] += p1 * xc; a[5] += p0 * xc * xc; a[6] += p1 * xc * xc;
  }
  __shared__ double red[7][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const double v = warp_sum(a[k]);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(&sums[threadIdx.x], v);
  }
}

// monotone float -> uint32 key (ascending order preserved; -0 < +0 is harmless here)
__device__ __forceinline__ unsigned order_key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// radix-select histograms.  level 0: hist[4096] over key>>20.  level 1: for slot s with key>>20 == prefixes[s]:
// hist[s*4096 + ((key>>8)&0xFFF)].  level 2: for slot s with key>>8 == prefixes[s]: hist[s*256 + (key&0xFF)].
__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ x, long long n, int level,
                                                          const unsigned* __restrict__ prefixes, int nprefix,
                                                          unsigned* __restrict__ hist) {
  __shared__ unsigned sh[4096];
  __shared__ unsigned pf[64];
  if (level == 0) {
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
  } else {
    for (int i = threadIdx.x; i < nprefix; i += blockDim.x) pf[i] = prefixes[i];
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = order_key(x[i]);
    if (level == 0) {
      atomicAdd(&sh[k >> 20], 1u);
    } else if (level == 1) {
      const unsigned top = k >> 20;
      for (int s = 0; s < nprefix; ++s)
        if (pf[s] == top) atomicAdd(&hist[(size_t)s * 4096 + ((k >> 8) & 0xFFFu)], 1u);
    } else {
      const unsigned top = k >> 8;
      for (int s = 0; s < nprefix; ++s)
        if (pf[s] == top) atomicAdd(&hist[(size_t)s * 256 + (k & 0xFFu)], 1u);
    }
  }
  if (level == 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
      if (sh[i]) atomicAdd(&hist[i], sh[i]);
  }
}

}  // namespace

extern "C" int tpz_gmm_sums(const float* x, long long n, int mode, const double* params8, double* sums7, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(n > 0 && (mode == 0 || mode == 1), "tpz_gmm_sums: bad arguments n=%lld mode=%d", n, mode);
  GmmParams P;
  P.shift = params8[0]; P.split = params8[1]; P.mu0 = params8[2]; P.mu1 = params8[3]; P.var0 = params8[4];
  P.var1 = params8[5]; P.log_prior0 = params8[6]; P.log_prior1 = params8[7];
  TPZ_CUDA(cudaMemsetAsync(sums7, 0, 7 * sizeof(double), stream));
  const int blocks = (int)(tpz_div_up(n, 256 * 8) < 148 * 8 ? tpz_div_up(n, 256 * 8) : 148 * 8);
  gmm_sums_kernel<<<blocks, 256, 0, stream>>>(x, n, mode, P, sums7);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_select_hist(const float* x, long long n, int level, const unsigned* prefixes, int nprefix,
                               unsigned* hist, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(n > 0 && level >= 0 && level <= 2 && nprefix >= 0 && nprefix <= 64, "tpz_select_hist: bad arguments");
  const size_t bins = level == 0 ? 4096 : (size_t)nprefix * (level == 1 ? 4096 : 256);
  TPZ_CUDA(cudaMemsetAsync(hist, 0, bins * sizeof(unsigned), stream));
  const int blocks = (int)(tpz_div_up(n, 256 * 8) < 148 * 8 ? tpz_div_up(n, 256 * 8) : 148 * 8);
  select_hist_kernel<<<blocks, 256, 0, stream>>>(x, n, level, prefixes, nprefix, hist);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

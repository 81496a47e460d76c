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

// One pass over the pixels for up to 12 parameter sets at once (the 12 initialisations of stats.py:89-110 advance in
// lockstep).  For set s:  sums[s][0..6] += { sum Z, sum p0, sum p1, sum p0*xc, sum p1*xc, sum p0*xc^2, sum p1*xc^2 },
// xc = x - shift.
// mode 0 (initial hard split, stats.py:136-139): p0 = (x <= split), p1 = 1 - p0, Z = 0
// mode 1 (E step, stats.py:158-167 / 172-203): log_pk = -(x-mu_k)^2/2/var_k - 0.5 log(2 pi var_k) + log prior_k,
//         Z = logsumexp, p_k = exp(log_pk - Z).  Means arrive in shifted coordinates (mu_k - shift).
constexpr int GMM_MAX_SETS = 12, GMM_PER_THREAD = 16;
struct GmmSet { double mode, split, mu0, mu1, var0, var1, log_prior0, log_prior1; };
struct GmmParams { int nsets; double shift; GmmSet set[GMM_MAX_SETS]; };

__global__ void __launch_bounds__(256) gmm_sums_kernel(const float* __restrict__ x, long long n, const GmmParams P,
                                                       double* __restrict__ sums) {
  __shared__ double acc[GMM_MAX_SETS][7][8];
  for (int i = threadIdx.x; i < GMM_MAX_SETS * 7 * 8; i += blockDim.x) (&acc[0][0][0])[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr long long TILE = 256 * GMM_PER_THREAD;
  for (long long base = (long long)blockIdx.x * TILE; base < n; base += (long long)gridDim.x * TILE) {
    float xf[GMM_PER_THREAD];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < GMM_PER_THREAD; ++k) {
      const long long i = base + k * 256 + threadIdx.x;
      xf[k] = i < n ? x[i] : 0.f;
      valid |= (i < n ? 1u : 0u) << k;
    }
    for (int s = 0; s < P.nsets; ++s) {
      const GmmSet& Q = P.set[s];
      double a[7] = {0, 0, 0, 0, 0, 0, 0};
      if (Q.mode == 0.0) {
        const float splitf = (float)Q.split;
#pragma unroll
        for (int k = 0; k < GMM_PER_THREAD; ++k) {
          if (!((valid >> k) & 1u)) continue;
          const double xc = (double)xf[k] - P.shift;
          const double p0 = xf[k] <= splitf ? 1.0 : 0.0, p1 = 1.0 - p0;
          a[1] += p0; a[2] += p1; a[3] += p0 * xc; a[4] += p1 * xc; a[5] += p0 * xc * xc; a[6] += p1 * xc * xc;
        }
      } else {
        const double c0 = -0.5 * log(2.0 * 3.14159265358979323846 * Q.var0) + Q.log_prior0;
        const double c1 = -0.5 * log(2.0 * 3.14159265358979323846 * Q.var1) + Q.log_prior1;
        const double h0 = 0.5 / Q.var0, h1 = 0.5 / Q.var1;
#pragma unroll 4
        for (int k = 0; k < GMM_PER_THREAD; ++k) {
          if (!((valid >> k) & 1u)) continue;
          const double xc = (double)xf[k] - P.shift;
          const double d0 = xc - Q.mu0, d1 = xc - Q.mu1;
          const double l0 = c0 - d0 * d0 * h0, l1 = c1 - d1 * d1 * h1;
          const double ma = fmax(l0, l1);
          const double e0 = exp(l0 - ma), e1 = exp(l1 - ma);
          const double sm = e0 + e1;
          const double p0 = e0 / sm, p1 = e1 / sm;
          a[0] += ma + log(sm);
          a[1] += p0; a[2] += p1; a[3] += p0 * xc; a[4] += p1 * xc; a[5] += p0 * xc * xc; a[6] += p1 * xc * xc;
        }
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const double v = warp_sum(a[j]);
        if (lane == 0) acc[s][j][warp] += v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < P.nsets * 7) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += acc[threadIdx.x / 7][threadIdx.x % 7][w];
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

extern "C" int tpz_gmm_sums(const float* x, long long n, double shift, const double* sets8, int nsets, double* sums,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(n > 0 && nsets >= 1 && nsets <= GMM_MAX_SETS, "tpz_gmm_sums: bad arguments n=%lld nsets=%d", n, nsets);
  GmmParams P;
  P.nsets = nsets;
  P.shift = shift;
  for (int s = 0; s < nsets; ++s) {
    const double* q = sets8 + 8 * s;
    TPZ_CHECK(q[0] == 0.0 || q[0] == 1.0, "tpz_gmm_sums: set %d has mode %g", s, q[0]);
    P.set[s] = GmmSet{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]};
  }
  TPZ_CUDA(cudaMemsetAsync(sums, 0, (size_t)nsets * 7 * sizeof(double), stream));
  const long long tiles = (n + 256 * GMM_PER_THREAD - 1) / (256 * GMM_PER_THREAD);
  const int blocks = (int)(tiles < 148 * 4 ? tiles : 148 * 4);
  gmm_sums_kernel<<<blocks, 256, 0, stream>>>(x, n, P, sums);
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

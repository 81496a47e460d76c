// Training path of the strided (unfilled) classifier: fp32 forward / dgrad / wgrad convolutions on NHWC fp32
// activations with the reference's OIHW fp32 weights (gradients land directly in .grad layout), the fused
// GE-binomial loss + closed-form score gradient, and a fused flat-buffer Adam step.
// Replaces, for reference topaz/methods.py:98-165 (GE_binomial.step): the cuDNN fwd/bwd-data/bwd-filter
// calls under loss.backward() (:146), ~25 tiny ATen kernels + the CPU scipy binom.logpmf + .item() syncs of
// the loss (:103-136), and torch.optim.Adam.step()/zero_grad() (:159-160).
#include <stdlib.h>
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {

struct ConvGeom {
  int N, H, W, Ci;      // input  (x / dx)
  int Ho, Wo, Co;       // output (y / dy)
  int kh, kw, stride, dil, org;   // input coord = o*stride + tap*dil + org
};

constexpr int TM = 64, TN = 64, TK = 16;

// MODE 0: forward   M = output pixels, N = Co, K = (tap, ci)
// MODE 1: dgrad     M = input pixels,  N = Ci, K = (tap, co)
// MODE 2: wgrad     M = Co,            N = (tap, ci), K = output pixels (split over blockIdx.z, atomic accumulate)
template <int MODE>
__global__ void __launch_bounds__(256) conv_f32_kernel(ConvGeom g, const float* __restrict__ A0 /*x | dy | dy*/,
                                                       const float* __restrict__ Wt /*w | w | x*/,
                                                       const float* __restrict__ bias, const float* __restrict__ res,
                                                       int res_H, int res_W, int res_org, int res_stride,
                                                       const float* __restrict__ mask, float* __restrict__ out,
                                                       int relu, int accumulate, int k_per_split) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int taps = g.kh * g.kw;
  long long Mtot, Ntot, Ktot;
  if (MODE == 0) { Mtot = (long long)g.N * g.Ho * g.Wo; Ntot = g.Co; Ktot = (long long)taps * g.Ci; }
  else if (MODE == 1) { Mtot = (long long)g.N * g.H * g.W; Ntot = g.Ci; Ktot = (long long)taps * g.Co; }
  else { Mtot = g.Co; Ntot = (long long)taps * g.Ci; Ktot = (long long)g.N * g.Ho * g.Wo; }
  const long long m0 = (long long)blockIdx.x * TM;
  const long long n0 = (long long)blockIdx.y * TN;
  long long kbeg = 0, kend = Ktot;
  if (MODE == 2) { kbeg = (long long)blockIdx.z * k_per_split; kend = min(Ktot, kbeg + k_per_split); }

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += TK) {
    // ---- stage A (TM x TK) and B (TK x TN): 4 elements each per thread ----
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;          // 0..1023
      {
        const int kk = idx & (TK - 1), mm = idx >> 4;       // consecutive threads -> consecutive k (channel-contiguous)
        const long long m = m0 + mm, k = k0 + kk;
        float v = 0.f;
        if (m < Mtot && k < kend) {
          if (MODE == 0) {
            const int ci = k % g.Ci, tap = k / g.Ci;
            const int r = tap / g.kw, t = tap - r * g.kw;
            const int ox = m % g.Wo; const long long q = m / g.Wo; const int oy = q % g.Ho; const int n = q / g.Ho;
            const int iy = oy * g.stride + r * g.dil + g.org, ix = ox * g.stride + t * g.dil + g.org;
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = A0[(((long long)n * g.H + iy) * g.W + ix) * g.Ci + ci];
          } else if (MODE == 1) {
            const int co = k % g.Co, tap = k / g.Co;
            const int r = tap / g.kw, t = tap - r * g.kw;
            const int ix = m % g.W; const long long q = m / g.W; const int iy = q % g.H; const int n = q / g.H;
            const int ny = iy - g.org - r * g.dil, nx = ix - g.org - t * g.dil;
            if (ny >= 0 && nx >= 0 && ny % g.stride == 0 && nx % g.stride == 0) {
              const int oy = ny / g.stride, ox = nx / g.stride;
              if (oy < g.Ho && ox < g.Wo) v = A0[(((long long)n * g.Ho + oy) * g.Wo + ox) * g.Co + co];
            }
          } else {
            v = A0[k * g.Co + m];     // dy[p][co]
          }
        }
        As[kk][mm] = v;
      }
      {
        float v = 0.f;
        if (MODE == 2) {
          const int nn = idx & (TN - 1), kk = idx >> 6;     // consecutive threads -> consecutive (tap,ci)
          const long long n = n0 + nn, k = k0 + kk;
          if (n < Ntot && k < kend) {
            const int ci = n % g.Ci, tap = n / g.Ci;
            const int r = tap / g.kw, t = tap - r * g.kw;
            const int ox = k % g.Wo; const long long q = k / g.Wo; const int oy = q % g.Ho; const int b = q / g.Ho;
            const int iy = oy * g.stride + r * g.dil + g.org, ix = ox * g.stride + t * g.dil + g.org;
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = Wt[(((long long)b * g.H + iy) * g.W + ix) * g.Ci + ci];
          }
          Bs[kk][nn] = v;
        } else {
          const int kk = idx & (TK - 1), nn = idx >> 4;
          const long long n = n0 + nn, k = k0 + kk;
          if (n < Ntot && k < kend) {
            if (MODE == 0) {
              const int ci = k % g.Ci, tap = k / g.Ci;
              v = Wt[((long long)n * g.Ci + ci) * taps + tap];              // w[co=n][ci][tap]
            } else {
              const int co = k % g.Co, tap = k / g.Co;
              v = Wt[((long long)co * g.Ci + n) * taps + tap];              // w[co][ci=n][tap]
            }
          }
          Bs[kk][nn] = v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
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

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= Ntot) continue;
      float v = acc[i][j];
      if (MODE == 0) {
        if (bias) v += bias[n];
        if (res) {
          const int ox = m % g.Wo; const long long q = m / g.Wo; const int oy = q % g.Ho; const int b = q / g.Ho;
          v += res[(((long long)b * res_H + (oy * res_stride + res_org)) * res_W + (ox * res_stride + res_org)) * g.Co + n];
        }
        if (relu) v = fmaxf(v, 0.f);
        out[m * g.Co + n] = v;
      } else if (MODE == 1) {
        const long long o = m * g.Ci + n;
        if (accumulate) v += out[o];
        if (mask) v = mask[o] > 0.f ? v : 0.f;
        out[o] = v;
      } else {
        const int ci = n % g.Ci, tap = n / g.Ci;
        atomicAdd(&out[((long long)m * g.Ci + ci) * taps + tap], v);       // dw[co=m][ci][tap]
      }
    }
  }
}

__global__ void relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dy[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// dx[n, o*s+org, o*s+org, c] += g[n, o, o, c]   (identity / strided-identity skip connection backward)
__global__ void crop_add_kernel(float* __restrict__ dx, int H, int W, const float* __restrict__ g, int N, int Ho, int Wo,
                                int C, int org, int stride) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C; long long q = i / C;
    const int ox = q % Wo; q /= Wo;
    const int oy = q % Ho; const int n = q / Ho;
    dx[(((long long)n * H + (oy * stride + org)) * W + (ox * stride + org)) * C + c] += g[i];
  }
}

// db[c] += sum_p dy[p][c].  Block = 32 channels x 8 row-slices; 4 independent loads in flight per thread.
__global__ void bias_grad_kernel(const float* __restrict__ dy, long long P, int C, float* __restrict__ db) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;            // 8 slices
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    const long long step = (long long)gridDim.y * 8;
    long long p = (long long)blockIdx.y * 8 + slice;
    for (; p + 3 * step < P; p += 4 * step) {
      s0 += dy[p * C + c]; s1 += dy[(p + step) * C + c]; s2 += dy[(p + 2 * step) * C + c]; s3 += dy[(p + 3 * step) * C + c];
    }
    for (; p < P; p += step) s0 += dy[p * C + c];
  }
  float s = (s0 + s1) + (s2 + s3);
  __shared__ float sh[8][33];
  sh[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && c < C) {
    for (int k = 1; k < 8; ++k) s += sh[k][threadIdx.x & 31];
    atomicAdd(&db[c], s);
  }
}

// -------------------------------------------------------------------------------------------------
// Cin = 1 first layer of the training net (7x7 stride 2): dedicated forward and weight-gradient kernels
// (the generic implicit-GEMM tiles are wasteful for K = 49, Ci = 1).
// -------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(128) first_fwd_kernel(const float* __restrict__ x, int N, int H, int W,
                                                        const float* __restrict__ w, const float* __restrict__ bias, int k,
                                                        int stride, int relu, float* __restrict__ y, int Ho, int Wo) {
  extern __shared__ float s_wf[];                 // [taps][CO]
  const int taps = k * k;
  for (int i = threadIdx.x; i < taps * CO; i += blockDim.x) {
    const int t = i / CO, c = i - t * CO;
    s_wf[i] = w[c * taps + t];
  }
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)N * Ho * Wo;
  if (m >= M) return;
  const int ox = m % Wo; const long long q = m / Wo; const int oy = q % Ho; const int n = q / Ho;
  const float* xb = x + ((long long)n * H + oy * stride) * W + ox * stride;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = bias ? bias[c] : 0.f;
  for (int r = 0; r < k; ++r)
    for (int t = 0; t < k; ++t) {
      const float v = xb[r * W + t];
      const float4* wr = reinterpret_cast<const float4*>(s_wf + (r * k + t) * CO);
#pragma unroll
      for (int c4 = 0; c4 < CO / 4; ++c4) {
        const float4 wv = wr[c4];
        acc[c4 * 4 + 0] = fmaf(v, wv.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(v, wv.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(v, wv.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(v, wv.w, acc[c4 * 4 + 3]);
      }
    }
  float4* o = reinterpret_cast<float4*>(y + m * CO);
#pragma unroll
  for (int c4 = 0; c4 < CO / 4; ++c4) {
    float4 v = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    o[c4] = v;
  }
}

// dw[co][tap] += sum_p dy[p][co] * x[p*stride + tap]; thread = (co, tap group), block = a contiguous pixel range
template <int CO>
__global__ void __launch_bounds__(256) first_wgrad_kernel(const float* __restrict__ x, int N, int H, int W,
                                                          const float* __restrict__ dy, int Ho, int Wo, int k, int stride,
                                                          float* __restrict__ dw, int px_per_block) {
  constexpr int TG = 256 / CO;                    // tap groups
  constexpr int MAXT = 16;                        // taps per thread (k*k <= TG*MAXT)
  const int co = threadIdx.x % CO, tg = threadIdx.x / CO;
  const int taps = k * k;
  float acc[MAXT];
  int toff[MAXT];
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    acc[j] = 0.f;
    const int t = tg + j * TG;
    toff[j] = (t < taps) ? (t / k) * W + (t % k) : -1;
  }
  const long long M = (long long)N * Ho * Wo;
  const long long p0 = (long long)blockIdx.x * px_per_block, p1 = min(M, p0 + px_per_block);
  if (p0 >= p1) return;
  int ox = p0 % Wo; long long q = p0 / Wo; int oy = q % Ho; int n = q / Ho;      // advanced incrementally below
  const float* xb = x + ((long long)n * H + oy * stride) * W + ox * stride;
  for (long long p = p0; p < p1; ++p) {
    const float v = dy[p * CO + co];
#pragma unroll
    for (int j = 0; j < MAXT; ++j)
      if (toff[j] >= 0) acc[j] = fmaf(v, __ldg(xb + toff[j]), acc[j]);
    xb += stride;
    if (++ox == Wo) {
      ox = 0;
      if (++oy == Ho) { oy = 0; ++n; }
      xb = x + ((long long)n * H + oy * stride) * W;
    }
  }
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    const int t = tg + j * TG;
    if (t < taps) atomicAdd(&dw[co * taps + t], acc[j]);
  }
}

// Row-strip version of the same reduction.  A block walks over (image, output row) strips: it stages the k input rows
// of the strip and the strip's dy[Wo][CO] in shared memory (coalesced), then every thread accumulates a 4-channel x
// MAXT-tap register block over the strip's pixels: per pixel one 16-byte dy load + MAXT broadcast x loads feed
// 4*MAXT FMAs (the pixel-per-iteration kernel above issues one global load per FMA and waits on each).  Partial sums
// stay in registers across all strips of the block; one atomicAdd per (block, weight) at the end.
template <int CO>
__global__ void __launch_bounds__(256) first_wgrad_strip_kernel(const float* __restrict__ x, int N, int H, int W,
                                                                const float* __restrict__ dy, int Ho, int Wo, int k,
                                                                int stride, float* __restrict__ dw, int xw) {
  constexpr int CQ = CO / 4;                      // threads along the channel quads
  constexpr int TG = 256 / CQ;                    // tap groups
  constexpr int MAXT = 64 / TG;                   // taps per thread (k*k <= 64)
  extern __shared__ __align__(16) float s_fw[];   // dy strip [Wo][CO] | x strip [k][xw]
  float* s_dy = s_fw;
  float* s_x = s_fw + Wo * CO;
  const int cq = threadIdx.x % CQ, tg = threadIdx.x / CQ;
  const int taps = k * k;
  float acc[MAXT][4];
  int xoff[MAXT];
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const int t = tg + j * TG;
    xoff[j] = (t < taps) ? (t / k) * xw + (t % k) : 0;      // invalid taps read x[0] of the strip and are dropped at the end
  }
  const int rows = N * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, oy = row - n * Ho;
    __syncthreads();                                // previous strip fully consumed
    {
      const float4* src = reinterpret_cast<const float4*>(dy + (long long)row * Wo * CO);
      float4* dst = reinterpret_cast<float4*>(s_dy);
      for (int i = threadIdx.x; i < Wo * CQ; i += 256) dst[i] = src[i];
      const float* xr = x + ((long long)n * H + (long long)oy * stride) * W;
      for (int i = threadIdx.x; i < k * xw; i += 256) {
        const int r = i / xw, c = i - r * xw;
        s_x[i] = xr[(long long)r * W + c];
      }
    }
    __syncthreads();
    const float4* d4 = reinterpret_cast<const float4*>(s_dy) + cq;
#pragma unroll 3
    for (int ox = 0; ox < Wo; ++ox) {
      const float4 d = d4[ox * CQ];
      const float* xp = s_x + ox * stride;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        const float xv = xp[xoff[j]];
        acc[j][0] = fmaf(d.x, xv, acc[j][0]);
        acc[j][1] = fmaf(d.y, xv, acc[j][1]);
        acc[j][2] = fmaf(d.z, xv, acc[j][2]);
        acc[j][3] = fmaf(d.w, xv, acc[j][3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    const int t = tg + j * TG;
    if (t < taps) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&dw[(cq * 4 + e) * taps + t], acc[j][e]);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// GE-binomial loss (methods.py:103-151), one block.  scores: all B_total logits of the (global) minibatch,
// labels (fp64, as the reference's DataLoader collates them).  Writes dscore for [lo, hi) (the local shard)
// and out[5] = {classifier_loss, ge_penalty, precision, tpr, fpr}.
// closed form: dL/ds_i = slack*p_i(1-p_i)[G_mu + (1-2p_i)G_v]  (unlabeled),
//                                               (sigmoid(s_i) - 1)/|P|                   (positives)
// -------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < (blockDim.x >> 5); ++i) r += sh[i];
  __syncthreads();
  return r;
}
__device__ double block_max(double v, double* sh) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(~0u, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = -1e300;
  for (int i = 0; i < (blockDim.x >> 5); ++i) r = fmax(r, sh[i]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) ge_binomial_kernel(const float* __restrict__ score, const double* __restrict__ label,
                                                          int B, double pi, double slack, int lo, int hi,
                                                          float* __restrict__ dscore, float* __restrict__ out5) {
  __shared__ double sh[8];
  double s_p_pos = 0, s_p_all = 0, s_p_unl = 0, n_pos = 0, n_unl = 0, mu = 0, var = 0, bce = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const double s = score[i];
    const double p = (double)(1.0f / (1.0f + expf(-(float)s)));     // torch.sigmoid in fp32
    s_p_all += p;
    if (label[i] == 1.0) {
      n_pos += 1; s_p_pos += p;
      bce += fmax(s, 0.0) - s + log1p(exp(-fabs(s)));    // BCEWithLogits, target 1 (fp64 like the reference)
    } else if (label[i] == 0.0) {
      n_unl += 1; s_p_unl += p; mu += p; var += p * (1.0 - p);
    }
  }
  s_p_pos = block_sum(s_p_pos, sh); s_p_all = block_sum(s_p_all, sh); s_p_unl = block_sum(s_p_unl, sh);
  n_pos = block_sum(n_pos, sh); n_unl = block_sum(n_unl, sh);
  mu = (double)(float)block_sum(mu, sh); var = (double)(float)block_sum(var, sh); bce = block_sum(bce, sh);
  const int N = (int)n_unl;
  const double denom = var + 1e-10;
  // q_k = softmax_k(-0.5 (mu-k)^2 / denom);  c_k = -logBinom(k; N, pi)
  const double lgN = lgamma(N + 1.0), lp = log(pi), l1p = log1p(-pi);
  double lmax = -1e300;
  for (int k = threadIdx.x; k <= N; k += blockDim.x) lmax = fmax(lmax, -0.5 * (mu - k) * (mu - k) / denom);
  lmax = block_max(lmax, sh);
  double Z = 0, Sc = 0, Sc1 = 0, Sc2 = 0, S1 = 0, S2 = 0;
  for (int k = threadIdx.x; k <= N; k += blockDim.x) {
    const double e = exp(-0.5 * (mu - k) * (mu - k) / denom - lmax);
    const double c = -(double)(float)(lgN - lgamma(k + 1.0) - lgamma(N - k + 1.0) + k * lp + (N - k) * l1p);
    const double a1 = -(mu - k) / denom;                       // d logit_k / d mu
    const double a2 = 0.5 * (mu - k) * (mu - k) / (denom * denom);   // d logit_k / d var
    Z += e; Sc += e * c; Sc1 += e * c * a1; Sc2 += e * c * a2; S1 += e * a1; S2 += e * a2;
  }
  Z = block_sum(Z, sh); Sc = block_sum(Sc, sh); Sc1 = block_sum(Sc1, sh); Sc2 = block_sum(Sc2, sh);
  S1 = block_sum(S1, sh); S2 = block_sum(S2, sh);
  const double ge = Sc / Z;                                       // -sum_k logBinom_k q_k
  const double Gmu = Sc1 / Z - ge * (S1 / Z);                     // sum_k q_k (c_k - ge) a1_k
  const double Gv = Sc2 / Z - ge * (S2 / Z);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double s = score[i];
    const double p = 1.0 / (1.0 + exp(-s));
    double gsc = 0.0;
    if (label[i] == 1.0) gsc = (p - 1.0) / n_pos;
    else if (label[i] == 0.0) gsc = slack * p * (1.0 - p) * (Gmu + (1.0 - 2.0 * p) * Gv);
    dscore[i - lo] = (float)gsc;
  }
  if (threadIdx.x == 0) {
    out5[0] = (float)(bce / n_pos);
    out5[1] = (float)ge;
    out5[2] = (float)(s_p_pos / s_p_all);
    out5[3] = (float)(s_p_pos / n_pos);
    out5[4] = (float)(s_p_unl / n_unl);
  }
}

// -------------------------------------------------------------------------------------------------
// The other PU-learning objectives of topaz/methods.py on the same skeleton (loss value(s), metrics, d/dscore):
//   mode 0  PN     (methods.py:25-74)   pi <= 0: mean BCE over all;  pi > 0: pi*BCE_pos + (1-pi)*BCE_neg
//   mode 1  GE_KL  (methods.py:168-255) BCE on positives + slack/momentum * KL(pi || p_hat), p_hat = momentum*mean_unl
//                  sigmoid + (1-momentum)*running;   aux_in = running expectation, out[5] = new running expectation
//   mode 2  PU     (methods.py:258-322) non-negative PU risk with clipping at -beta (aux_in = beta)
// out6 = {loss (classifier loss for GE_KL), ge_penalty (GE_KL) or 0, precision, tpr, fpr, aux_out}
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_objective_kernel(const float* __restrict__ score, const double* __restrict__ label,
                                                           int B, int mode, double pi, double slack, double momentum,
                                                           double aux_in, int lo, int hi, float* __restrict__ dscore,
                                                           float* __restrict__ out6) {
  __shared__ double sh[8];
  double s_p_pos = 0, s_p_all = 0, s_p_unl = 0, n_pos = 0, n_unl = 0, bce_pos1 = 0, bce_pos0 = 0, bce_unl0 = 0, bce_all = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const double s = score[i];
    const double p = (double)(1.0f / (1.0f + expf(-(float)s)));
    const double sp = fmax(s, 0.0) + log1p(exp(-fabs(s)));      // softplus(s) = BCE(s, target 0)
    s_p_all += p;
    if (label[i] == 1.0) { n_pos += 1; s_p_pos += p; bce_pos1 += sp - s; bce_pos0 += sp; bce_all += sp - s; }
    else if (label[i] == 0.0) { n_unl += 1; s_p_unl += p; bce_unl0 += sp; bce_all += sp; }
    else { bce_all += sp - s * label[i]; }
  }
  s_p_pos = block_sum(s_p_pos, sh); s_p_all = block_sum(s_p_all, sh); s_p_unl = block_sum(s_p_unl, sh);
  n_pos = block_sum(n_pos, sh); n_unl = block_sum(n_unl, sh);
  bce_pos1 = block_sum(bce_pos1, sh); bce_pos0 = block_sum(bce_pos0, sh); bce_unl0 = block_sum(bce_unl0, sh);
  bce_all = block_sum(bce_all, sh);
  double loss = 0, ge = 0, aux = 0;
  double w_pos_a = 0, w_pos_b = 0, w_unl = 0;      // dscore = w_pos_a*(p-1) + w_pos_b*p  (positives),  w_unl*p (Y==0)
  double kl_coef = 0;                              // GE_KL: extra term kl_coef * p(1-p) on Y==0
  if (mode == 0) {
    if (pi > 0) {
      loss = pi * bce_pos1 / n_pos + (1 - pi) * bce_unl0 / n_unl;
      w_pos_a = pi / n_pos; w_unl = (1 - pi) / n_unl;
    } else {
      loss = bce_all / B;
      w_pos_a = 1.0 / B; w_unl = 1.0 / B;
    }
  } else if (mode == 1) {
    loss = bce_pos1 / n_pos;
    w_pos_a = 1.0 / n_pos;
    double p_hat = s_p_unl / n_unl;
    if (momentum < 1) p_hat = momentum * p_hat + (1 - momentum) * aux_in;
    aux = p_hat;
    const double entropy = pi * log(pi) + (1 - pi) * log1p(-pi);
    ge = (-log(p_hat) * pi - log1p(-p_hat) * (1 - pi) + entropy) * slack / momentum;
    kl_coef = (slack / momentum) * (-pi / p_hat + (1 - pi) / (1 - p_hat)) * momentum / n_unl;
  } else {
    const double loss_pp = bce_pos1 / n_pos, loss_pn = bce_pos0 / n_pos, loss_un = bce_unl0 / n_unl;
    const double loss_u = loss_un - loss_pn * pi;
    if (loss_u < -aux_in) {                  // clipped: step along -loss_u, report pi*loss_pp - beta
      loss = loss_pp * pi - aux_in;
      w_pos_b = pi / n_pos; w_unl = -1.0 / n_unl;
    } else {
      loss = loss_pp * pi + loss_u;
      w_pos_a = pi / n_pos; w_pos_b = -pi / n_pos; w_unl = 1.0 / n_unl;
    }
  }
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double s = score[i];
    const double p = 1.0 / (1.0 + exp(-s));
    double g = 0.0;
    if (label[i] == 1.0) g = w_pos_a * (p - 1.0) + w_pos_b * p;
    else if (label[i] == 0.0) g = w_unl * p + kl_coef * p * (1.0 - p);
    else if (mode == 0 && pi <= 0) g = (p - label[i]) / B;
    dscore[i - lo] = (float)g;
  }
  if (threadIdx.x == 0) {
    out6[0] = (float)loss; out6[1] = (float)ge;
    out6[2] = (float)(s_p_pos / s_p_all); out6[3] = (float)(s_p_pos / n_pos); out6[4] = (float)(s_p_unl / n_unl);
    out6[5] = (float)aux;
  }
}

// fused Adam on flat buffers (torch.optim.Adam defaults: no weight decay, no amsgrad) + gradient zeroing;
// optional L2 term l2*w added to the gradient (methods.py:153-157: d/dw of 0.5*l2*sum w^2)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float l2,
                            float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale + l2 * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    g[i] = 0.f;
  }
}

// Adam with the step count kept on the device (CUDA-graph replay of a whole training step: the bias corrections must not be
// baked into the captured launch).  step_inc_kernel advances the counter, adam_dev_kernel derives the corrections from it.
__global__ void step_inc_kernel(int* step) { step[0] += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step, float l2,
                                float gscale) {
  const float t = (float)step[0];
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale + l2 * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    g[i] = 0.f;
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static ConvGeom geom(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int kh, int kw, int stride, int dil, int org) {
  ConvGeom g; g.N = N; g.H = H; g.W = W; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.kh = kh; g.kw = kw;
  g.stride = stride; g.dil = dil; g.org = org; return g;
}

extern "C" int tpz_conv_fwd_f32(const float* x, int N, int H, int W, int Ci, const float* w, const float* bias, int Co,
                                int kh, int kw, int stride, int dil, int org, const float* res, int res_H, int res_W,
                                int res_org, int res_stride, int relu, float* y, int Ho, int Wo, void* stream) {
  const ConvGeom g = geom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  dim3 grid(tpz_div_up((long long)N * Ho * Wo, TM), tpz_div_up(Co, TN), 1);
  conv_f32_kernel<0><<<grid, 256, 0, ST(stream)>>>(g, x, w, bias, res, res_H, res_W, res_org, res_stride, nullptr, y,
                                                   relu, 0, 0);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_dgrad_f32(const float* dy, int N, int Ho, int Wo, int Co, const float* w, int Ci, int kh, int kw,
                                  int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx, int H,
                                  int W, void* stream) {
  const ConvGeom g = geom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  dim3 grid(tpz_div_up((long long)N * H * W, TM), tpz_div_up(Ci, TN), 1);
  conv_f32_kernel<1><<<grid, 256, 0, ST(stream)>>>(g, dy, w, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate, 0);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_wgrad_f32(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co,
                                  int kh, int kw, int stride, int dil, int org, float* dw, float* db, void* stream) {
  const ConvGeom g = geom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  const long long P = (long long)N * Ho * Wo;
  const int mt = tpz_div_up(Co, TM), nt = tpz_div_up((long long)kh * kw * Ci, TN);
  int splits = (148 * 4) / (mt * nt);
  if (splits < 1) splits = 1;
  long long kps = (P + splits - 1) / splits;
  kps = (kps + TK - 1) / TK * TK;
  splits = (int)((P + kps - 1) / kps);
  dim3 grid(mt, nt, splits);
  conv_f32_kernel<2><<<grid, 256, 0, ST(stream)>>>(g, dy, x, nullptr, nullptr, 0, 0, 0, 1, nullptr, dw, 0, 1, (int)kps);
  if (db) {
    dim3 bg(tpz_div_up(Co, 32), (unsigned)(P < 4096 ? 1 : (P < 65536 ? 64 : 296)));
    bias_grad_kernel<<<bg, 256, 0, ST(stream)>>>(dy, P, Co, db);
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_first_fwd_f32(const float* x, int N, int H, int W, const float* w, const float* bias, int Co, int k,
                                 int stride, int relu, float* y, int Ho, int Wo, void* stream) {
  TPZ_CHECK(Co == 32 || Co == 64, "tpz_first_fwd_f32: Co must be 32 or 64 (got %d)", Co);
  const long long M = (long long)N * Ho * Wo;
  const size_t smem = (size_t)k * k * Co * sizeof(float);
  if (Co == 32) first_fwd_kernel<32><<<tpz_div_up(M, 128), 128, smem, ST(stream)>>>(x, N, H, W, w, bias, k, stride, relu, y, Ho, Wo);
  else first_fwd_kernel<64><<<tpz_div_up(M, 128), 128, smem, ST(stream)>>>(x, N, H, W, w, bias, k, stride, relu, y, Ho, Wo);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_first_wgrad_f32(const float* x, int N, int H, int W, const float* dy, int Ho, int Wo, int Co, int k,
                                   int stride, float* dw, void* stream) {
  TPZ_CHECK(Co == 32 || Co == 64, "tpz_first_wgrad_f32: Co must be 32 or 64 (got %d)", Co);
  TPZ_CHECK(k * k <= (256 / Co) * 16, "tpz_first_wgrad_f32: kernel %dx%d too large", k, k);
  const long long M = (long long)N * Ho * Wo;
  {
    // row-strip kernel (default); TPZ_FIRST_WGRAD=v1 selects the pixel-per-iteration kernel (A/B switch)
    static const bool strip = []() { const char* e = getenv("TPZ_FIRST_WGRAD"); return !(e && strcmp(e, "v1") == 0); }();
    const int xw = (Wo - 1) * stride + k;
    const size_t smem = ((size_t)Wo * Co + (size_t)k * xw) * sizeof(float);
    if (strip && k * k <= 64 && xw <= W && (Ho - 1) * stride + k <= H && smem <= 48 * 1024) {
      const long long rows = (long long)N * Ho;
      const int grid = (int)(rows < 148 * 6 ? rows : 148 * 6);
      if (Co == 32) first_wgrad_strip_kernel<32><<<grid, 256, smem, ST(stream)>>>(x, N, H, W, dy, Ho, Wo, k, stride, dw, xw);
      else first_wgrad_strip_kernel<64><<<grid, 256, smem, ST(stream)>>>(x, N, H, W, dy, Ho, Wo, k, stride, dw, xw);
      TPZ_CUDA(cudaGetLastError());
      return 0;
    }
  }
  int ppb = (int)((M + 148 * 8 - 1) / (148 * 8));
  if (ppb < 64) ppb = 64;
  const int grid = tpz_div_up(M, ppb);
  if (Co == 32) first_wgrad_kernel<32><<<grid, 256, 0, ST(stream)>>>(x, N, H, W, dy, Ho, Wo, k, stride, dw, ppb);
  else first_wgrad_kernel<64><<<grid, 256, 0, ST(stream)>>>(x, N, H, W, dy, Ho, Wo, k, stride, dw, ppb);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// ---- classifier head of the training net: 1x1 conv C -> 1 on M pixels (classifier.py:29,65; M = the minibatch on training crops).
// The generic direct-conv kernels spend 36 + 24 us on this 64 KFLOP op (4 thread blocks); one warp per pixel / one fused backward.
namespace {
__global__ void cls_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, long long M, int C,
                               float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + m * C);
  float acc = 0.f;
  for (int c = lane; c < (C >> 2); c += 32) {
    const float4 a = __ldg(xr + c);
    // parameters live in one flat buffer (train_engine.FlatParams): a 1-element PReLU slope before them breaks 16-byte alignment
    const float4 k = make_float4(__ldg(w + 4 * c), __ldg(w + 4 * c + 1), __ldg(w + 4 * c + 2), __ldg(w + 4 * c + 3));
    acc = fmaf(a.x, k.x, acc); acc = fmaf(a.y, k.y, acc); acc = fmaf(a.z, k.z, acc); acc = fmaf(a.w, k.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[m] = acc + (b ? __ldg(b) : 0.f);
}
// dx[m][c] = (mask: x[m][c] > 0) ? g[m]*w[c] : 0;  dw[c] += sum_m g[m]*x[m][c];  db += sum_m g[m].  Block = 8 warps x 4 rows.
__global__ void cls_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ g, long long M, int C,
                               int masked, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float s_dw[];                      // [C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_dw[c] = 0.f;
  __syncthreads();
  const long long m0 = (long long)blockIdx.x * 32 + warp * 4;
  float gsum = 0.f;
  for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
    const float4 k = make_float4(__ldg(w + 4 * c4), __ldg(w + 4 * c4 + 1), __ldg(w + 4 * c4 + 2), __ldg(w + 4 * c4 + 3));   // see cls_fwd_kernel
    float4 dwv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + i;
      if (m < M) {
        const float gm = __ldg(g + m);
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + m * C) + c4);
        dwv.x = fmaf(gm, a.x, dwv.x); dwv.y = fmaf(gm, a.y, dwv.y); dwv.z = fmaf(gm, a.z, dwv.z); dwv.w = fmaf(gm, a.w, dwv.w);
        float4 d = make_float4(gm * k.x, gm * k.y, gm * k.z, gm * k.w);
        if (masked) { d.x = a.x > 0.f ? d.x : 0.f; d.y = a.y > 0.f ? d.y : 0.f; d.z = a.z > 0.f ? d.z : 0.f; d.w = a.w > 0.f ? d.w : 0.f; }
        if (dx) reinterpret_cast<float4*>(dx + m * C)[c4] = d;
      }
    }
    atomicAdd(&s_dw[4 * c4], dwv.x); atomicAdd(&s_dw[4 * c4 + 1], dwv.y); atomicAdd(&s_dw[4 * c4 + 2], dwv.z); atomicAdd(&s_dw[4 * c4 + 3], dwv.w);
  }
  if (lane < 4 && m0 + lane < M) gsum = __ldg(g + m0 + lane);
#pragma unroll
  for (int o = 2; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
  if (lane == 0 && db) atomicAdd(db, gsum);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dw + c, s_dw[c]);
}
}  // namespace

extern "C" int tpz_cls_fwd_f32(const float* x, long long M, int C, const float* w, const float* bias, float* y, void* stream) {
  TPZ_CHECK(C % 4 == 0 && C > 0, "tpz_cls_fwd_f32: C must be a multiple of 4 (C=%d)", C);
  cls_fwd_kernel<<<tpz_div_up(M, 8), 256, 0, ST(stream)>>>(x, w, bias, M, C, y);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_cls_bwd_f32(const float* x, long long M, int C, const float* w, const float* g, int masked, float* dx, float* dw,
                               float* db, void* stream) {
  TPZ_CHECK(C % 4 == 0 && C > 0 && C <= 8192, "tpz_cls_bwd_f32: C must be a multiple of 4, at most 8192 (C=%d)", C);
  cls_bwd_kernel<<<tpz_div_up(M, 32), 256, C * sizeof(float), ST(stream)>>>(x, w, g, M, C, masked, dx, dw, db);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_bias_grad_f32(const float* dy, long long P, int C, float* db, void* stream) {
  dim3 bg(tpz_div_up(C, 32), (unsigned)(P < 4096 ? 1 : (P < 65536 ? 64 : 296)));
  bias_grad_kernel<<<bg, 256, 0, ST(stream)>>>(dy, P, C, db);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_relu_bwd_f32(float* dy, const float* y, long long n, void* stream) {
  int grid = tpz_div_up(n, 256 * 4); if (grid > 148 * 8) grid = 148 * 8;
  relu_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(dy, y, n);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_crop_add_f32(float* dx, int N, int H, int W, int C, const float* g, int Ho, int Wo, int org, int stride,
                                void* stream) {
  const long long total = (long long)N * Ho * Wo * C;
  int grid = tpz_div_up(total, 256 * 4); if (grid > 148 * 8) grid = 148 * 8;
  crop_add_kernel<<<grid, 256, 0, ST(stream)>>>(dx, H, W, g, N, Ho, Wo, C, org, stride);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_ge_binomial_loss_grad(const float* scores, const double* labels, int B, double pi, double slack,
                                         int lo, int hi, float* dscores, float* out5, void* stream) {
  TPZ_CHECK(B > 0 && lo >= 0 && hi <= B && lo <= hi, "tpz_ge_binomial_loss_grad: bad shard [%d,%d) of %d", lo, hi, B);
  TPZ_CHECK(pi > 0.0 && pi < 1.0, "tpz_ge_binomial_loss_grad: pi=%g must be in (0,1)", pi);
  ge_binomial_kernel<<<1, 256, 0, ST(stream)>>>(scores, labels, B, pi, slack, lo, hi, dscores, out5);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_pu_objective_loss_grad(const float* scores, const double* labels, int B, int mode, double pi, double slack,
                                          double momentum, double aux_in, int lo, int hi, float* dscores, float* out6,
                                          void* stream) {
  TPZ_CHECK(B > 0 && lo >= 0 && hi <= B && lo <= hi, "tpz_pu_objective_loss_grad: bad shard [%d,%d) of %d", lo, hi, B);
  TPZ_CHECK(mode >= 0 && mode <= 2, "tpz_pu_objective_loss_grad: mode %d", mode);
  pu_objective_kernel<<<1, 256, 0, ST(stream)>>>(scores, labels, B, mode, pi, slack, momentum, aux_in, lo, hi, dscores, out6);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, int step, float l2, float grad_scale, void* stream) {
  TPZ_CHECK(step >= 1, "tpz_adam_step: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  int grid = tpz_div_up(n, 256 * 4); if (grid > 148 * 8) grid = 148 * 8;
  adam_kernel<<<grid, 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2, l2,
                                            grad_scale);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                 float beta1, float beta2, float eps, int* step_dev, float l2, float grad_scale, void* stream) {
  TPZ_CHECK(step_dev != nullptr, "tpz_adam_step_dev: null step counter");
  int grid = tpz_div_up(n, 256 * 4); if (grid > 148 * 8) grid = 148 * 8;
  step_inc_kernel<<<1, 1, 0, ST(stream)>>>(step_dev);
  adam_dev_kernel<<<grid, 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step_dev, l2,
                                                grad_scale);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// Tensor-core (mma.sync TF32, error-compensated "3xTF32") versions of the training convolutions for channel
// counts that are multiples of 16/32: forward and data-gradient as one gather-GEMM kernel, weight-gradient as a
// split-K GEMM.  fp32 in / fp32 out; the 3-term split (a = a_hi + a_lo, a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi)
// keeps fp32-level accuracy so gradients match the fp32 CPU reference to ~1e-6.
// The training step is latency-bound: these kernels matter for instruction count, not peak FLOPs;
// the legacy mma.sync path is used on purpose (tiny, ragged tiles; tcgen05 needs 128-row fp16 tiles).
// Same reference call sites as tpz_train.cu (topaz/methods.py:103,146).
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

#include <stdlib.h>
// X3 = 1: error-compensated 3xTF32 (fp32-level accuracy, default).  X3 = 0 (env TPZ_TRAIN_TF32=1): single-pass TF32,
// the precision of the reference's own cuDNN path (torch.backends.cudnn.allow_tf32 = True), ~1.3x faster step.
// X3 = 2 (default since round 2, validated on B200; TPZ_TRAIN_SPLIT=rna switches back to X3 = 1): 3xTF32 with a 3-instruction
// operand split.  On sm_100a `cvt.rna.tf32.f32` is emulated (FSETP inf/nan guard + predicated integer add of half an ulp + LOP3
// mask, `profiles/r01_sass_train_mma_s2.md`), so the default split costs 7 instructions per operand value and the hot loops
// issue 4.4-6.8 instructions per HMMA.  The fast split rounds hi with the same add+mask but without the guard (inf stays inf;
// NaN payloads are irrelevant here) and hands lo = x - hi to the MMA unrounded (the tensor core ignores the 13 low mantissa
// bits of a .tf32 operand, i.e. truncates: an extra error of < 2^-21 |x| per product).

namespace {

struct MGeom {
  int N, H, W, Ci;   // conv input  (x / dx)
  int Ho, Wo, Co;    // conv output (y / dy)
  int kh, kw, stride, dil, org;
};

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// split an fp32 fragment into TF32 hi (+ lo) parts once; reused by every MMA that consumes the fragment
template <int X3, int N>
__device__ __forceinline__ void split_tf32(const float (&v)[N], uint32_t (&hi)[N], uint32_t (&lo)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (X3 == 2) {
      hi[i] = (__float_as_uint(v[i]) + 0x1000u) & 0xffffe000u;          // round-half-away to 10 mantissa bits (finite inputs)
      lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));           // exact in fp32; truncated by the tensor core
    } else {
      hi[i] = to_tf32(v[i]);
      lo[i] = X3 ? to_tf32(v[i] - __uint_as_float(hi[i])) : 0u;
    }
  }
}
template <int X3>
__device__ __forceinline__ void mma_split(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  if (X3) {
    mma_tf32(c, al, bh);
    mma_tf32(c, ah, bl);
  }
  mma_tf32(c, ah, bh);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;      // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// -------------------------------------------------------------------------------------------------
// weight repack: OIHW [Co][Ci][taps] -> fwd layout [tap][ci][co] and dgrad layout [tap][co][ci]
// -------------------------------------------------------------------------------------------------
struct RepackDesc { long long src, dst_fwd, dst_dg; int Co, Ci, taps, pad; };

__global__ void repack_kernel(const float* __restrict__ flat, const RepackDesc* __restrict__ descs, float* __restrict__ packed) {
  const RepackDesc d = descs[blockIdx.y];
  const long long n = (long long)d.Co * d.Ci * d.taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = i % d.taps;
    const long long q = i / d.taps;
    const int ci = q % d.Ci, co = q / d.Ci;
    const float v = flat[d.src + i];
    packed[d.dst_fwd + ((long long)tap * d.Ci + ci) * d.Co + co] = v;
    packed[d.dst_dg + ((long long)tap * d.Co + co) * d.Ci + ci] = v;
  }
}

// -------------------------------------------------------------------------------------------------
// forward / dgrad gather-GEMM.  M = pixels of the OUTPUT tensor of this op (fwd: y pixels, dgrad: x pixels),
// K = (tap, source channel), N = output channels.  A(m,(tap,c)) = SRC[srcpix(m,tap)][c], B = packed weights.
// -------------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBK = 16, GAS = 20;   // A smem row stride (floats): conflict-free fragment reads

template <int BN, int MODE, int X3>   // MODE 0 fwd, 1 dgrad; X3: see the top of the file
__global__ void __launch_bounds__(256) conv_mma_kernel(MGeom g, const float* __restrict__ src, const float* __restrict__ wpk,
                                                       const float* __restrict__ bias, const float* __restrict__ res,
                                                       int res_H, int res_W, int res_org, int res_stride,
                                                       const float* __restrict__ mask, float* __restrict__ out, int relu,
                                                       int accumulate) {
  constexpr int BS = BN + 8;
  constexpr int WN = BN / 32;            // warps along N (1 or 2)
  constexpr int WM = 8 / WN;             // warps along M (8 or 4)
  constexpr int MT = GBM / WM / 16;      // m16 tiles per warp (1 or 2)
  constexpr int STG = 3;                 // cp.async stages: the step is latency-bound, keep 2 chunks in flight
  __shared__ __align__(16) float As[STG][GBM][GAS];
  __shared__ __align__(16) float Bs[STG][GBK][BS];

  const int taps = g.kh * g.kw;
  const int Cs = MODE == 0 ? g.Ci : g.Co;      // source channels (K per tap)
  const int Nn = MODE == 0 ? g.Co : g.Ci;      // output channels
  const int MH = MODE == 0 ? g.Ho : g.H, MW = MODE == 0 ? g.Wo : g.W;
  const int SH = MODE == 0 ? g.H : g.Ho, SW = MODE == 0 ? g.W : g.Wo;
  const long long Mtot = (long long)g.N * MH * MW;
  const long long m0 = (long long)blockIdx.x * GBM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = warp / WN, wn = warp % WN;

  // staging roles: A: 2 rows per thread (r, r+64), one float4 (quad) each; B: GBK*BN/4 float4 over 256 threads
  const int a_row = tid >> 2, a_quad = tid & 3;
  int an[2], ay[2], ax[2];
  bool arow_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long m = m0 + a_row + h * 64;
    arow_ok[h] = m < Mtot;
    const long long mm = arow_ok[h] ? m : 0;
    ax[h] = mm % MW;
    const long long q = mm / MW;
    ay[h] = q % MH;
    an[h] = q / MH;
  }
  const int cchunks = Cs / GBK;
  const int nk = taps * cchunks;

  // The tensor core's fp32 accumulate truncates (measured: a chain of ~430 mma ops drifts by 2e-5 relative, biased), which
  // is enough to flip ReLU masks of near-zero activations against the fp32 reference.  In the compensated mode the MMA
  // chain is therefore cut every FLUSH k-chunks and the partial sums are added with round-to-nearest FADDs.
  constexpr int FLUSH = 4;
  float acc[MT][4][4], tot[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc[i][j][e] = 0.f; tot[i][j][e] = 0.f; }

  // issue-side running state: chunks are issued in order, so (tap row, tap col, channel chunk) advance incrementally
  // and the per-row gather offsets are recomputed only when the tap changes (no divisions in the steady state)
  int i_r = 0, i_t = 0, i_c0 = 0;
  long long a_off[2] = {0, 0};
  bool a_ok[2] = {false, false};
  auto tap_setup = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      bool ok = arow_ok[h];
      int sy = 0, sx = 0;
      if (ok) {
        if (MODE == 0) {
          sy = ay[h] * g.stride + i_r * g.dil + g.org;
          sx = ax[h] * g.stride + i_t * g.dil + g.org;
          ok = sy >= 0 && sy < SH && sx >= 0 && sx < SW;
        } else {
          const int ny = ay[h] - g.org - i_r * g.dil, nx = ax[h] - g.org - i_t * g.dil;
          ok = ny >= 0 && nx >= 0;
          if (g.stride == 1) { sy = ny; sx = nx; }
          else { ok = ok && (ny % g.stride) == 0 && (nx % g.stride) == 0; sy = ny / g.stride; sx = nx / g.stride; }
          ok = ok && sy < SH && sx < SW;
        }
      }
      a_ok[h] = ok;
      a_off[h] = ok ? (((long long)an[h] * SH + sy) * SW + sx) * Cs + a_quad * 4 : 0;
    }
  };
  tap_setup();
  const int b_kk = tid / (BN / 4), b_nq = tid - b_kk * (BN / 4);     // B staging role (threads < GBK*BN/4)
  const bool b_role = tid < GBK * BN / 4;
  const bool b_ok = b_role && (n0 + b_nq * 4 < Nn);
  auto issue_chunk = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      cp_async16(&As[buf][a_row + h * 64][a_quad * 4], src + a_off[h] + i_c0, a_ok[h]);
    if (b_role) {
      const int tap = i_r * g.kw + i_t;
      const float* gp = b_ok ? wpk + ((long long)tap * Cs + i_c0 + b_kk) * Nn + n0 + b_nq * 4 : wpk;
      cp_async16(&Bs[buf][b_kk][b_nq * 4], gp, b_ok);
    }
    i_c0 += GBK;
    if (i_c0 == Cs) {
      i_c0 = 0;
      if (++i_t == g.kw) { i_t = 0; ++i_r; }
      tap_setup();
    }
  };

#pragma unroll
  for (int s2 = 0; s2 < STG - 1; ++s2) {
    if (s2 < nk) issue_chunk(s2);
    cp_async_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc % STG;
    cp_async_wait<STG - 2>();
    __syncthreads();
    if (kc + STG - 1 < nk) issue_chunk((kc + STG - 1) % STG);
    cp_async_commit();
#pragma unroll
    for (int k8 = 0; k8 < GBK; k8 += 8) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float bf[2] = {Bs[buf][k8 + tq][wn * 32 + j * 8 + gq], Bs[buf][k8 + tq + 4][wn * 32 + j * 8 + gq]};
        split_tf32<X3>(bf, bh[j], bl[j]);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int rb0 = wm * (GBM / WM) + i * 16;
        const float af[4] = {As[buf][rb0 + gq][k8 + tq], As[buf][rb0 + gq + 8][k8 + tq], As[buf][rb0 + gq][k8 + tq + 4],
                             As[buf][rb0 + gq + 8][k8 + tq + 4]};
        uint32_t ah[4], al[4];
        split_tf32<X3>(af, ah, al);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_split<X3>(acc[i][j], ah, al, bh[j], bl[j]);
      }
    }
    if (X3 && (kc % FLUSH) == FLUSH - 1) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) { tot[i][j][e] += acc[i][j][e]; acc[i][j][e] = 0.f; }
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] += tot[i][j][e];

  // epilogue: c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
#pragma unroll
  for (int i = 0; i < MT; ++i) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const long long m = m0 + wm * (GBM / WM) + i * 16 + gq + hh * 8;
      if (m >= Mtot) continue;
      long long rbase = 0;
      if (MODE == 0 && res) {
        const int ox = m % g.Wo; const long long q = m / g.Wo; const int oy = q % g.Ho; const int b = q / g.Ho;
        rbase = (((long long)b * res_H + (oy * res_stride + res_org)) * res_W + (ox * res_stride + res_org)) * g.Co;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn * 32 + j * 8 + tq * 2;
        if (n >= Nn) continue;
        float v0 = acc[i][j][hh * 2], v1 = acc[i][j][hh * 2 + 1];
        const long long o = m * Nn + n;
        if (MODE == 0) {
          if (bias) { v0 += bias[n]; v1 += bias[n + 1]; }
          if (res) { v0 += res[rbase + n]; v1 += res[rbase + n + 1]; }
          if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        } else {
          if (accumulate) { v0 += out[o]; v1 += out[o + 1]; }
          if (mask) { v0 = mask[o] > 0.f ? v0 : 0.f; v1 = mask[o + 1] > 0.f ? v1 : 0.f; }
        }
        *reinterpret_cast<float2*>(out + o) = make_float2(v0, v1);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// wgrad: dw[co][ci][tap] += sum_p dy[p][co] * x[p'(p,tap)][ci].  CTA = (64 co x 64 ci) tile of one tap over a
// K-split of the output pixels; fp32 atomics accumulate the splits (dw is zeroed by the Adam kernel).
// -------------------------------------------------------------------------------------------------
constexpr int WBK = 16;

template <int BT, int X3>    // square (BT co) x (BT ci) tile, BT = 64 or 32
__global__ void __launch_bounds__(256) wgrad_mma_kernel(MGeom g, const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dw, int k_per_split) {
  constexpr int STG = 4;
  constexpr int WSS = BT + 8;
  constexpr int WTM = BT / 2, WTN = BT / 4;        // warp tile (2 x 4 warps)
  constexpr int MI = WTM / 16, NJ = WTN / 8;
  __shared__ __align__(16) float As[STG][WBK][WSS];   // [pixel][co]
  __shared__ __align__(16) float Bs[STG][WBK][WSS];   // [pixel][ci]
  const int taps = g.kh * g.kw;
  const int tap = blockIdx.z % taps, split = blockIdx.z / taps;
  const int r = tap / g.kw, t = tap - r * g.kw;
  const int co0 = blockIdx.x * BT, ci0 = blockIdx.y * BT;
  const long long P = (long long)g.N * g.Ho * g.Wo;
  const long long pbeg = (long long)split * k_per_split;
  const long long pend = min(P, pbeg + k_per_split);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  // staging: 16 pixels x BT channels per operand.  BT=64: every thread copies one float4 of A and one of B;
  // BT=32: threads 0-127 copy A, threads 128-255 copy B.
  constexpr int QP = BT / 4;                        // float4 per pixel row
  const int role = (BT == 64) ? 2 : (tid >> 7);     // 0: A only, 1: B only, 2: both
  const int sidx = (BT == 64) ? tid : (tid & 127);
  const int s_p = sidx / QP, s_q = sidx % QP;

  constexpr int FLUSH = 4;          // see conv_mma_kernel: cut the truncating MMA accumulation chain every 4 chunks
  float acc[MI][NJ][4], tot[MI][NJ][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc[i][j][e] = 0.f; tot[i][j][e] = 0.f; }

  // running (n, oy, ox) of this thread's staging pixel; advanced by WBK pixels per issued chunk (no divisions)
  long long ip = pbeg + s_p;
  int i_ox, i_oy, i_n;
  {
    const long long pp = ip < P ? ip : 0;
    i_ox = pp % g.Wo; const long long q = pp / g.Wo; i_oy = q % g.Ho; i_n = q / g.Ho;
  }
  auto issue_chunk = [&](int buf) {
    const bool pv = ip < pend;
    if (role != 1) {
      const bool ok = pv && (co0 + s_q * 4 < g.Co);
      cp_async16(&As[buf][s_p][s_q * 4], ok ? dy + ip * g.Co + co0 + s_q * 4 : dy, ok);
    }
    if (role != 0) {
      const int iy = i_oy * g.stride + r * g.dil + g.org, ix = i_ox * g.stride + t * g.dil + g.org;
      const bool ok = pv && (ci0 + s_q * 4 < g.Ci) && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const float* gp = ok ? x + (((long long)i_n * g.H + iy) * g.W + ix) * g.Ci + ci0 + s_q * 4 : x;
      cp_async16(&Bs[buf][s_p][s_q * 4], gp, ok);
    }
    ip += WBK;
    i_ox += WBK;
    while (i_ox >= g.Wo) { i_ox -= g.Wo; if (++i_oy == g.Ho) { i_oy = 0; ++i_n; } }
  };
  const int nchunks = (int)((pend - pbeg + WBK - 1) / WBK);
#pragma unroll
  for (int s2 = 0; s2 < STG - 1; ++s2) {
    if (s2 < nchunks) issue_chunk(s2);
    cp_async_commit();
  }
  for (int it = 0; it < nchunks; ++it) {
    const int buf = it % STG;
    cp_async_wait<STG - 2>();
    __syncthreads();
    if (it + STG - 1 < nchunks) issue_chunk((it + STG - 1) % STG);
    cp_async_commit();
#pragma unroll
    for (int k8 = 0; k8 < WBK; k8 += 8) {
      uint32_t bh[NJ][2], bl[NJ][2];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float bf[2] = {Bs[buf][k8 + tq][wn * WTN + j * 8 + gq], Bs[buf][k8 + tq + 4][wn * WTN + j * 8 + gq]};
        split_tf32<X3>(bf, bh[j], bl[j]);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int mb = wm * WTM + i * 16;
        const float af[4] = {As[buf][k8 + tq][mb + gq], As[buf][k8 + tq][mb + gq + 8], As[buf][k8 + tq + 4][mb + gq],
                             As[buf][k8 + tq + 4][mb + gq + 8]};
        uint32_t ah[4], al[4];
        split_tf32<X3>(af, ah, al);
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma_split<X3>(acc[i][j], ah, al, bh[j], bl[j]);
      }
    }
    if (X3 && (it % FLUSH) == FLUSH - 1) {
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) { tot[i][j][e] += acc[i][j][e]; acc[i][j][e] = 0.f; }
    }
  }
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] += tot[i][j][e];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int co = co0 + wm * WTM + i * 16 + gq + hh * 8;
      if (co >= g.Co) continue;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = ci0 + wn * WTN + j * 8 + tq * 2 + e;
          if (ci < g.Ci) atomicAdd(&dw[((long long)co * g.Ci + ci) * taps + tap], acc[i][j][hh * 2 + e]);
        }
    }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int g_tf32_mode = -1;   // -1: read the environment on first use
static bool use_x3() {
  if (g_tf32_mode < 0) {
    const char* e = getenv("TPZ_TRAIN_TF32");
    g_tf32_mode = (e && e[0] == '1') ? 1 : 0;
  }
  return g_tf32_mode == 0;
}
// template argument X3 of the kernels: 0 single-pass TF32, 1 3xTF32 (default), 2 3xTF32 with the fast split (opt-in)
static int x3_mode() {
  // default since round 2: the 3-instruction split (validated on B200: the 47 training parity tests pass unchanged, step
  // 3.10 -> 2.65 ms without / 3.45 -> 3.01 ms with BatchNorm, profiles/r02_bench_split_ab.jsonl); TPZ_TRAIN_SPLIT=rna selects
  // the cvt.rna split again
  static const bool fast = []() { const char* e = getenv("TPZ_TRAIN_SPLIT"); return !(e && strcmp(e, "rna") == 0); }();
  return use_x3() ? (fast ? 2 : 1) : 0;
}
extern "C" int tpz_train_set_tf32(int single_pass) {
  int prev = use_x3() ? 0 : 1;
  g_tf32_mode = single_pass ? 1 : 0;
  return prev;
}

static MGeom mgeom(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int kh, int kw, int stride, int dil, int org) {
  MGeom g; g.N = N; g.H = H; g.W = W; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.kh = kh; g.kw = kw;
  g.stride = stride; g.dil = dil; g.org = org; return g;
}

// descs: device array of {src, dst_fwd, dst_dg (element offsets), Co, Ci, taps, pad}; one launch repacks all layers
extern "C" int tpz_train_repack(const float* flat_params, const void* descs, int ndesc, long long max_elems, float* packed,
                                void* stream) {
  if (ndesc == 0) return 0;
  dim3 grid(tpz_div_up(max_elems, 256 * 4), ndesc);
  repack_kernel<<<grid, 256, 0, ST(stream)>>>(flat_params, reinterpret_cast<const RepackDesc*>(descs), packed);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_fwd_mma(const float* x, int N, int H, int W, int Ci, const float* w_fwd_packed, const float* bias,
                                int Co, int kh, int kw, int stride, int dil, int org, const float* res, int res_H, int res_W,
                                int res_org, int res_stride, int relu, float* y, int Ho, int Wo, void* stream) {
  TPZ_CHECK(Ci % 16 == 0 && Co % 32 == 0, "tpz_conv_fwd_mma: needs Ci%%16==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  const MGeom g = mgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  const long long M = (long long)N * Ho * Wo;
  if (Co % 64 == 0) {
    dim3 grid(tpz_div_up(M, GBM), Co / 64);
    switch (x3_mode()) {
      case 1: conv_mma_kernel<64, 0, 1><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
      case 2: conv_mma_kernel<64, 0, 2><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
      default: conv_mma_kernel<64, 0, 0><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
    }
  } else {
    dim3 grid(tpz_div_up(M, GBM), Co / 32);
    switch (x3_mode()) {
      case 1: conv_mma_kernel<32, 0, 1><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
      case 2: conv_mma_kernel<32, 0, 2><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
      default: conv_mma_kernel<32, 0, 0><<<grid, 256, 0, ST(stream)>>>(g, x, w_fwd_packed, bias, res, res_H, res_W, res_org, res_stride, nullptr, y, relu, 0); break;
    }
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// Plain row-major fp32 product C[M][N] = A[M][K] * B[K][N] on the same gather-GEMM kernel (a 1x1 "convolution" over M
// pixels), always error-compensated 3xTF32.  Used by the Fourier-crop downsample (utils/image.py:38-61).
extern "C" int tpz_gemm_f32(const float* A, long long M, int K, const float* B, int N, float* Cout, void* stream) {
  TPZ_CHECK(K % 16 == 0 && N % 32 == 0 && M > 0 && M < (1ll << 31), "tpz_gemm_f32: needs K%%16==0, N%%32==0 (M=%lld K=%d N=%d)", M, K, N);
  const MGeom g = mgeom(1, 1, (int)M, K, 1, (int)M, N, 1, 1, 1, 1, 0);
  if (N % 64 == 0) {
    dim3 grid(tpz_div_up(M, GBM), N / 64);
    conv_mma_kernel<64, 0, true><<<grid, 256, 0, ST(stream)>>>(g, A, B, nullptr, nullptr, 0, 0, 0, 1, nullptr, Cout, 0, 0);
  } else {
    dim3 grid(tpz_div_up(M, GBM), N / 32);
    conv_mma_kernel<32, 0, true><<<grid, 256, 0, ST(stream)>>>(g, A, B, nullptr, nullptr, 0, 0, 0, 1, nullptr, Cout, 0, 0);
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_dgrad_mma(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh,
                                  int kw, int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx,
                                  int H, int W, void* stream) {
  TPZ_CHECK(Co % 16 == 0 && Ci % 32 == 0, "tpz_conv_dgrad_mma: needs Co%%16==0 and Ci%%32==0 (Ci=%d Co=%d)", Ci, Co);
  const MGeom g = mgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  const long long M = (long long)N * H * W;
  if (Ci % 64 == 0) {
    dim3 grid(tpz_div_up(M, GBM), Ci / 64);
    switch (x3_mode()) {
      case 1: conv_mma_kernel<64, 1, 1><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
      case 2: conv_mma_kernel<64, 1, 2><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
      default: conv_mma_kernel<64, 1, 0><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
    }
  } else {
    dim3 grid(tpz_div_up(M, GBM), Ci / 32);
    switch (x3_mode()) {
      case 1: conv_mma_kernel<32, 1, 1><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
      case 2: conv_mma_kernel<32, 1, 2><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
      default: conv_mma_kernel<32, 1, 0><<<grid, 256, 0, ST(stream)>>>(g, dy, w_dg_packed, nullptr, nullptr, 0, 0, 0, 1, relu_mask, dx, 0, accumulate); break;
    }
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_wgrad_mma(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co,
                                  int kh, int kw, int stride, int dil, int org, float* dw, void* stream) {
  TPZ_CHECK(Ci % 4 == 0 && Co % 4 == 0, "tpz_conv_wgrad_mma: needs Ci%%4==0 and Co%%4==0 (Ci=%d Co=%d)", Ci, Co);
  const MGeom g = mgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  const long long P = (long long)N * Ho * Wo;
  const int taps = kh * kw;
  const int BT = (Co <= 32 && Ci <= 32) ? 32 : 64;
  const int mt = tpz_div_up(Co, BT), nt = tpz_div_up(Ci, BT);
  int splits = (148 * 8) / (mt * nt * taps);
  if (splits < 1) splits = 1;
  long long kps = (P + splits - 1) / splits;
  kps = (kps + WBK - 1) / WBK * WBK;
  if (kps < 512) kps = 512;
  splits = (int)((P + kps - 1) / kps);
  dim3 grid(mt, nt, taps * splits);
  if (BT == 32) {
    switch (x3_mode()) {
      case 1: wgrad_mma_kernel<32, 1><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
      case 2: wgrad_mma_kernel<32, 2><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
      default: wgrad_mma_kernel<32, 0><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
    }
  } else {
    switch (x3_mode()) {
      case 1: wgrad_mma_kernel<64, 1><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
      case 2: wgrad_mma_kernel<64, 2><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
      default: wgrad_mma_kernel<64, 0><<<grid, 256, 0, ST(stream)>>>(g, x, dy, dw, (int)kps); break;
    }
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

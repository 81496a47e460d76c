// tcgen05 (5th-generation tensor core) versions of the TRAINING convolutions: forward, data-gradient and weight-gradient
// of the strided classifier (reference call sites: `score = self.model(X)` methods.py:103, `loss.backward()` :146).
//
// fp32 in / fp32 out at fp32-level accuracy: error-compensated 3xTF32 on `tcgen05.mma.kind::tf32`.  Every operand value
// x is split ONCE, by the thread that stages it, into hi = x rounded to 10 mantissa bits and lo = x - hi (exact in fp32;
// the tensor core reads the top 19 bits of lo), and the product is accumulated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi in the
// fp32 TMEM accumulator.  The mma.sync kernels (tpz_train_mma.cu) re-split every fragment in every warp that uses it and
// spend 3-7 issue slots per HMMA on it; here the split costs 3 ALU ops per staged value and the MMAs are issued by one thread.
//
//   forward / dgrad: gather-GEMM.  M = 128 pixels of the op's OUTPUT tensor (thread m stages row m: the 32 channels of one
//       tap of its source pixel = one 128-byte row, written hi / lo into the canonical K-major SWIZZLE_128B layout with
//       generic stores + fence.proxy.async -- the tpz_first_tc.cu pattern), N = 32 / 64 output channels, K = (tap, 32-channel
//       chunk).  Weights are pre-split per step by tpz_train_repack_tc into [tap][chunk][n][32] hi / lo planes and arrive by TMA.
//   wgrad: D[(tap, ci)][co] = sum over pixels of x[p'(p, tap)][ci] * dy[p][co].  M = 128 (tap, ci) pairs, N = co tile,
//       K = 32 output pixels per chunk; both operands are pixel-major in memory, so the staging warps transpose while they
//       write (lane = pixel: a warp's 32 four-byte stores fill one 128-byte row, conflict-free).  Split-K over the pixels,
//       fp32 atomics into the flat gradient (zeroed by the Adam kernel).
//   Accumulation: the tensor core's fp32 accumulate truncates (tpz_train_mma.cu: a chain of ~430 accumulations drifts by 2e-5,
//       enough to flip ReLU masks against the fp32 reference).  The MMA chain alternates between two TMEM accumulators every
//       `flush` chunks and the finished one is drained into registers with round-to-nearest FADDs while the other one fills.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <stdlib.h>

namespace {

struct TGeom {
  int N, H, W, Ci;   // conv input  (x / dx)
  int Ho, Wo, Co;    // conv output (y / dy)
  int kh, kw, stride, dil, org;
};

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol error traps after a few seconds instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins)
    if (spins > (1u << 27)) __trap();
}

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 operands (fp32 containers, low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::tf32: A/B = TF32 (2), D = F32 (1), both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// hi = x rounded (half away from zero) to 10 mantissa bits, lo = x - hi (exact); finite inputs
__device__ __forceinline__ void split1(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  lo = v - hi;
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  split1(v.x, hi.x, lo.x); split1(v.y, hi.y, lo.y); split1(v.z, hi.z, lo.z); split1(v.w, hi.w, lo.w);
}

constexpr int kStageA = 128 * 128;          // one operand plane of a stage: 128 rows x 128 B (32 fp32)

// -------------------------------------------------------------------------------------------------
// weight repack for the TMA-fed B operand: OIHW -> fwd [tap][ci/32][co][32] and dgrad [tap][co/32][ci][32], each as a hi plane
// followed by a lo plane
// -------------------------------------------------------------------------------------------------
struct RepackTcDesc { long long src, dst_fwd, dst_dg; int Co, Ci, taps, pad; };

__global__ void repack_tc_kernel(const float* __restrict__ flat, const RepackTcDesc* __restrict__ descs, float* __restrict__ packed) {
  const RepackTcDesc d = descs[blockIdx.y];
  const long long n = (long long)d.Co * d.Ci * d.taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = i % d.taps;
    const long long q = i / d.taps;
    const int ci = q % d.Ci, co = q / d.Ci;
    float hi, lo;
    split1(flat[d.src + i], hi, lo);
    const long long f = (((long long)tap * (d.Ci >> 5) + (ci >> 5)) * d.Co + co) * 32 + (ci & 31);
    const long long g = (((long long)tap * (d.Co >> 5) + (co >> 5)) * d.Ci + ci) * 32 + (co & 31);
    packed[d.dst_fwd + f] = hi; packed[d.dst_fwd + n + f] = lo;
    packed[d.dst_dg + g] = hi; packed[d.dst_dg + n + g] = lo;
  }
}

// -------------------------------------------------------------------------------------------------
// forward / dgrad gather-GEMM
// -------------------------------------------------------------------------------------------------
struct FwdArgs {
  TGeom g;
  const float* src;       // fwd: x [N][H][W][Ci]; dgrad: dy [N][Ho][Wo][Co]
  CUtensorMap tmB;        // packed weights: rows of 32 fp32 (declared to TMA as 64 fp16), box = BN rows
  long long lo_rows;      // row offset of the lo plane inside the packed weight tensor
  const float* bias; const float* res; int res_H, res_W, res_org, res_stride;
  const float* mask; float* out; int relu, accumulate, flush;
  int lat, lat0;          // dgrad with dil % stride == 0: only positions y = lat0 + i*lat (same for x) receive gradient; the M rows
                          // enumerate that sub-lattice (lat = 1: all positions) and the rest of dx is zero-filled by the caller
  int ksplit;             // forward, tiny grids (the last 5x5 layer: 4 tiles x 50 chunks): blockIdx.z takes `ksplit` consecutive
                          // (tap, chunk) blocks and adds its raw partial sums into the zero-filled output; bias / ReLU run afterwards
  int scatter;            // dgrad of a conv whose output is ONE pixel per image (the last 5x5 layer on a training crop): every dx
                          // pixel has exactly one valid tap, so blockIdx.y selects the tap, the rows are the images and K = Co
};

constexpr int kPF = 3;            // chunks of A prefetched into registers ahead of the one being staged (hides the gather latency)

template <int BN, int MODE>      // MODE 0 forward, 1 data gradient
__global__ void __launch_bounds__(256, 2) conv_tc_kernel(const __grid_constant__ FwdArgs a) {
  constexpr uint32_t IDESC = idesc_tf32(128, BN);
  constexpr int B_BYTES = BN * 128;
  constexpr int ASTAGE = 2 * kStageA;                       // A_hi | A_lo, two stages
  constexpr int BSTAGE = 2 * B_BYTES;                       // B_hi | B_lo, three stages: the TMA of chunk kc+1 is issued while
                                                            // chunk kc is staged, so its latency hides behind that work
  constexpr int TCOLS = 2 * BN < 32 ? 32 : 2 * BN;          // two accumulators
  constexpr int HN = BN / 2;                                // output channels per thread (two threads share a row)

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_free[2], bar_b[3], bar_acc[2];
  __shared__ uint32_t tmem_base_s;

  const TGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;    // thread (row, half) stages 16-byte chunks 4*half .. 4*half+3 of its row
  const int taps = g.kh * g.kw;
  const int Cs = MODE == 0 ? g.Ci : g.Co;        // source channels (K per tap)
  const int Nn = MODE == 0 ? g.Co : g.Ci;        // output channels
  const int lat = MODE == 1 ? a.lat : 1, lat0 = MODE == 1 ? a.lat0 : 0;
  const int MH = MODE == 0 ? g.Ho : (g.H - lat0 + lat - 1) / lat, MW = MODE == 0 ? g.Wo : (g.W - lat0 + lat - 1) / lat;
  const int SH = MODE == 0 ? g.H : g.Ho, SW = MODE == 0 ? g.W : g.Wo;
  const bool sc = MODE == 1 && a.scatter != 0;
  const long long Mtot = sc ? (long long)g.N : (long long)g.N * MH * MW;
  const long long m = (long long)blockIdx.x * 128 + row;
  const int n0 = sc ? (int)blockIdx.z * BN : (int)blockIdx.y * BN;     // scatter: blockIdx.y is the tap, blockIdx.z the N tile
  const int cchunks = Cs >> 5;
  const bool ks = MODE == 0 && a.ksplit > 0;
  // first weight block and block count of this CTA (scatter: the blocks of tap blockIdx.y; split-K: a slice of all blocks)
  const int kb0 = sc ? (int)blockIdx.y * cchunks : (ks ? (int)blockIdx.z * a.ksplit : 0);
  const int nk = sc ? cchunks : (ks ? min(a.ksplit, taps * cchunks - kb0) : taps * cchunks);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&bar_free[s], 1); ptx::mbar_init(&bar_acc[s], 1); }
    for (int s = 0; s < 3; ++s) ptx::mbar_init(&bar_b[s], 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tmB);
  }
  if (warp == 0) ptx::tmem_alloc<TCOLS>(&tmem_base_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  unsigned char* const bbase = base + 2 * ASTAGE;
  auto issue_b = [&](int kc) {                              // one elected thread asks TMA for the hi and lo weight blocks of chunk kc
    unsigned char* bst = bbase + (size_t)(kc % 3) * BSTAGE;
    ptx::mbar_expect_tx(&bar_b[kc % 3], 2u * B_BYTES);
    ptx::tma_load_2d(bst, &a.tmB, &bar_b[kc % 3], 0, (kb0 + kc) * Nn + n0);
    ptx::tma_load_2d(bst + B_BYTES, &a.tmB, &bar_b[kc % 3], 0, (int)a.lo_rows + (kb0 + kc) * Nn + n0);
  };
  if (tid == 0) issue_b(0);

  // this thread's output pixel
  const bool row_ok = m < Mtot;
  int px = 0, py = 0, pn = 0;
  if (row_ok) { px = (int)(m % MW); const long long q = m / MW; py = (int)(q % MH); pn = (int)(q / MH); }
  if (MODE == 1) { px = lat0 + px * lat; py = lat0 + py * lat; }
  int t_r = 0, t_t = 0, t_c = 0;
  if (sc) {                                                  // rows = images; the dx pixel is the one this tap reaches from dy (0, 0)
    t_r = (int)blockIdx.y / g.kw; t_t = (int)blockIdx.y - t_r * g.kw;
    pn = (int)m; py = g.org + t_r * g.dil; px = g.org + t_t * g.dil;
  }
  if (ks) { const int tap = kb0 / cchunks; t_c = (kb0 - tap * cchunks) * 32; t_r = tap / g.kw; t_t = tap - t_r * g.kw; }
  const long long mo = MODE == 1 ? ((long long)pn * g.H + py) * g.W + px : m;      // row of the output tensor

  // source pixel of (row, tap): recomputed when the tap changes
  long long a_off = 0;
  bool a_ok = false;
  auto tap_setup = [&]() {
    bool ok = row_ok;
    int sy = 0, sx = 0;
    if (ok) {
      if (MODE == 0) {
        sy = py * g.stride + t_r * g.dil + g.org;
        sx = px * g.stride + t_t * g.dil + g.org;
        ok = sy >= 0 && sy < SH && sx >= 0 && sx < SW;
      } else {
        const int ny = py - g.org - t_r * g.dil, nx = px - g.org - t_t * g.dil;
        ok = ny >= 0 && nx >= 0;
        if (g.stride == 1) { sy = ny; sx = nx; }
        else { ok = ok && (ny % g.stride) == 0 && (nx % g.stride) == 0; sy = ny / g.stride; sx = nx / g.stride; }
        ok = ok && sy < SH && sx < SW;
      }
    }
    a_ok = ok;
    a_off = ok ? (((long long)pn * SH + sy) * SW + sx) * Cs + half * 16 : 0;
  };
  tap_setup();

  // this thread's half of one chunk of its row: 16 fp32 = 4 x 16 B, prefetched into registers kPF chunks ahead
  float4 pre[kPF][4];
  auto load_chunk = [&](float4 (&dst)[4]) {
    const float4* p = reinterpret_cast<const float4*>(a.src + a_off + t_c);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = a_ok ? __ldg(p + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    t_c += 32;
    if (t_c == Cs) {
      t_c = 0;
      if (++t_t == g.kw) { t_t = 0; ++t_r; }
      tap_setup();
    }
  };
#pragma unroll
  for (int i = 0; i < kPF; ++i) {
    if (i < nk) load_chunk(pre[i]);
  }

  float tot[HN];
#pragma unroll
  for (int j = 0; j < HN; ++j) tot[j] = 0.f;
  const uint32_t d_hi = ptx::umma_desc_hi(1024, 2);         // SBO = 8 rows x 128 B, SWIZZLE_128B
  const int flush = a.flush > 0 ? a.flush : (1 << 30);
  const int sw = row & 7;
  uint32_t accf = 0;                                        // 0: the next MMA overwrites its accumulator
  int cur_acc = 0, in_group = 0, groups_done = 0;
  // TMEM accumulator -> registers, round-to-nearest adds.  Warp w reads lanes 32*(w%4) .. (its rows) and columns half*HN ..
  auto drain = [&](int acc) {
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * BN + half * HN;
    if (HN == 32) {
      uint32_t r[32];
      ptx::tmem_ld32(taddr, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) tot[j] += __uint_as_float(r[j]);
    } else {
      uint32_t r[16];
      ptx::tmem_ld16(taddr, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) tot[j] += __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
  };

  // One chunk: stage `cur` (this thread's part of its row for chunk kc, loaded kPF chunks ago) and refill `cur` with the
  // load for chunk kc + kPF.  The caller unrolls kPF chunks so that the prefetch queue is indexed at compile time: a
  // register-to-register rotation of the queue would make every iteration wait for the newest load.
  auto chunk_body = [&](const int kc, float4 (&cur)[4]) {
    const int s = kc & 1;
    unsigned char* stage = base + (size_t)s * ASTAGE;
    if (kc >= 2) mbar_wait(&bar_free[s], ((kc >> 1) - 1) & 1);       // the MMAs that read this stage (chunk kc-2) are done
    // B of the NEXT chunk: its ring slot (kc+1) % 3 was last read by chunk kc-2, which the wait above has seen complete
    if (tid == 0 && kc + 1 < nk) issue_b(kc + 1);
    // A: split this thread's part of the row and store it (16-byte chunk c of row p lands at chunk c ^ (p & 7))
    {
      unsigned char* rh = stage + row * 128;
      unsigned char* rl = rh + kStageA;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 hi, lo;
        split4(cur[c], hi, lo);
        const int cc = half * 4 + c;
        *reinterpret_cast<float4*>(rh + ((cc ^ sw) << 4)) = hi;
        *reinterpret_cast<float4*>(rl + ((cc ^ sw) << 4)) = lo;
      }
    }
    // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: the membar waits for every memory operation this thread has
    // in flight, so a prefetch issued BEFORE it is not a prefetch (measured: ~3 us per chunk, the full gather latency).  The next
    // loads are therefore issued after the fence and the barrier.
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (kc + kPF < nk) load_chunk(cur);                     // flies during the MMA issue and the next chunk's staging
    if (warp == 0) {
      ptx::tc_fence_after();
      mbar_wait(&bar_b[kc % 3], (kc / 3) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t ah = (ptx::smem_u32(stage) & 0x3FFFF) >> 4, al = ah + (kStageA >> 4);
        const uint32_t bh = (ptx::smem_u32(bbase + (size_t)(kc % 3) * BSTAGE) & 0x3FFFF) >> 4, bl = bh + (B_BYTES >> 4);
        const uint32_t d = tmem_base + cur_acc * BN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_tf32(d, al + 2 * k, d_hi, bh + 2 * k, d_hi, IDESC, accf | (uint32_t)k);      // small terms first
          umma_tf32(d, ah + 2 * k, d_hi, bl + 2 * k, d_hi, IDESC, 1u);
          umma_tf32(d, ah + 2 * k, d_hi, bh + 2 * k, d_hi, IDESC, 1u);
        }
        ptx::umma_commit(&bar_free[s]);
        if (in_group + 1 == flush || kc + 1 == nk) ptx::umma_commit(&bar_acc[cur_acc]);
      }
      __syncwarp();
    }
    accf = 1;
    if (++in_group == flush || kc + 1 == nk) {
      // this accumulator's group is issued; drain the PREVIOUS group's accumulator while these MMAs run
      if (groups_done >= 1) {
        const int prev = cur_acc ^ 1;
        mbar_wait(&bar_acc[prev], ((groups_done - 1) >> 1) & 1);
        ptx::tc_fence_after();
        drain(prev);
      }
      ++groups_done;
      cur_acc ^= 1; in_group = 0; accf = 0;
    }
  };
  for (int kc0 = 0; kc0 < nk; kc0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      if (kc0 + u < nk) chunk_body(kc0 + u, pre[u]);
    }
  }
  {                                                         // last group
    const int last = cur_acc ^ 1;
    mbar_wait(&bar_acc[last], ((groups_done - 1) >> 1) & 1);
    ptx::tc_fence_after();
    drain(last);
  }

  // epilogue: this thread owns output row m, channels n0 + half*HN .. + HN
  if (row_ok) {
    const int nb = n0 + half * HN;
    long long rbase = 0;
    if (MODE == 0 && a.res)
      rbase = (((long long)pn * a.res_H + (py * a.res_stride + a.res_org)) * a.res_W + (px * a.res_stride + a.res_org)) * g.Co;
    float* orow = a.out + mo * Nn + nb;
    if (ks) {                                                // split-K partial: raw sums, finished by bias_act_kernel
#pragma unroll
      for (int j = 0; j < HN; ++j) atomicAdd(orow + j, tot[j]);
    } else
#pragma unroll
    for (int c = 0; c < HN; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tot[c + j];
      if (MODE == 0) {
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += __ldg(a.bias + nb + c + j);
        }
        if (a.res) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb + c));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb + c + 4));
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
      } else {
        if (a.accumulate) {
          const float4 o0 = *reinterpret_cast<const float4*>(orow + c), o1 = *reinterpret_cast<const float4*>(orow + c + 4);
          v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w; v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
        }
        if (a.mask) {
          const float4 k0 = __ldg(reinterpret_cast<const float4*>(a.mask + mo * Nn + nb + c));
          const float4 k1 = __ldg(reinterpret_cast<const float4*>(a.mask + mo * Nn + nb + c + 4));
          v[0] = k0.x > 0.f ? v[0] : 0.f; v[1] = k0.y > 0.f ? v[1] : 0.f; v[2] = k0.z > 0.f ? v[2] : 0.f; v[3] = k0.w > 0.f ? v[3] : 0.f;
          v[4] = k1.x > 0.f ? v[4] : 0.f; v[5] = k1.y > 0.f ? v[5] : 0.f; v[6] = k1.z > 0.f ? v[6] : 0.f; v[7] = k1.w > 0.f ? v[7] : 0.f;
        }
      }
      ptx::st_global_256(orow + c, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]),
                         __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<TCOLS>(tmem_base);
}

// -------------------------------------------------------------------------------------------------
// wgrad
// -------------------------------------------------------------------------------------------------
struct WgArgs {
  TGeom g;
  const float* x; const float* dy; float* dw;
  int k_per_split;        // output pixels per split (multiple of 32)
  int flush;
};

constexpr int kPFW = 2;           // chunks prefetched ahead in the wgrad kernel

template <int BN>                // co tile
__global__ void __launch_bounds__(256, 2) wgrad_tc_kernel(const WgArgs a) {
  constexpr uint32_t IDESC = idesc_tf32(128, BN);
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = 2 * kStageA + 2 * B_BYTES;
  constexpr int TCOLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int BQ = BN / 8;                                  // dy channels staged per warp (8 or 4)
  constexpr int HN = BN / 2;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_free[2], bar_acc[2];
  __shared__ uint32_t tmem_base_s;

  const TGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int taps = g.kh * g.kw;
  const int rows_total = taps * g.Ci;                         // (tap, ci) pairs = M extent
  const int mt = blockIdx.x;                                  // M tile
  const int co0 = blockIdx.y * BN;
  const long long P = (long long)g.N * g.Ho * g.Wo;
  const long long pbeg = (long long)blockIdx.z * a.k_per_split;
  const long long pend = min(P, pbeg + a.k_per_split);
  const int nchunks = (int)((pend - pbeg + 31) / 32);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&bar_free[s], 1); ptx::mbar_init(&bar_acc[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<TCOLS>(&tmem_base_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // warp w stages A rows [16w, 16w + 16) of the tile = 16 consecutive ci of ONE tap (Ci % 32 == 0)
  const int row0 = mt * 128 + warp * 16;
  const bool warp_rows_ok = row0 < rows_total;
  const int tap = warp_rows_ok ? row0 / g.Ci : 0, ci0 = warp_rows_ok ? row0 % g.Ci : 0;
  const int tr = tap / g.kw, tt = tap - tr * g.kw;
  if (!warp_rows_ok) {                                        // rows beyond the (tap, ci) range stay zero in both stages
    for (int s = 0; s < 2; ++s)
      for (int j = 0; j < 16; ++j) {
        unsigned char* rh = base + (size_t)s * STAGE + (size_t)(warp * 16 + j) * 128;
        *reinterpret_cast<float*>(rh + lane * 4) = 0.f;
        *reinterpret_cast<float*>(rh + kStageA + lane * 4) = 0.f;
      }
  }

  // lane = pixel of the chunk; running (n, oy, ox) of that pixel
  long long ip = pbeg + lane;
  int i_ox, i_oy, i_n;
  {
    const long long pp = ip < P ? ip : 0;
    i_ox = (int)(pp % g.Wo); const long long q = pp / g.Wo; i_oy = (int)(q % g.Ho); i_n = (int)(q / g.Ho);
  }
  float4 prex[kPFW][4];          // x[p'(p, tap)][ci0 .. ci0+16)
  float4 prey[kPFW][BQ / 4];     // dy[p][co0 + warp*BQ .. + BQ)
  auto load_chunk = [&](float4 (&dx_)[4], float4 (&dy_)[BQ / 4]) {
    const bool pv = ip < pend;
    if (warp_rows_ok) {
      const int iy = i_oy * g.stride + tr * g.dil + g.org, ix = i_ox * g.stride + tt * g.dil + g.org;
      const bool ok = pv && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const float4* p = reinterpret_cast<const float4*>(a.x + (((long long)i_n * g.H + iy) * g.W + ix) * g.Ci + ci0);
#pragma unroll
      for (int c = 0; c < 4; ++c) dx_[c] = ok ? __ldg(p + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4* q = reinterpret_cast<const float4*>(a.dy + ip * g.Co + co0 + warp * BQ);
#pragma unroll
    for (int c = 0; c < BQ / 4; ++c) dy_[c] = pv ? __ldg(q + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ip += 32;
    i_ox += 32;
    while (i_ox >= g.Wo) { i_ox -= g.Wo; if (++i_oy == g.Ho) { i_oy = 0; ++i_n; } }
  };
#pragma unroll
  for (int i = 0; i < kPFW; ++i) {
    if (i < nchunks) load_chunk(prex[i], prey[i]);
  }

  float tot[HN];
#pragma unroll
  for (int j = 0; j < HN; ++j) tot[j] = 0.f;
  const uint32_t d_hi = ptx::umma_desc_hi(1024, 2);
  const int flush = a.flush > 0 ? a.flush : (1 << 30);
  uint32_t accf = 0;
  int cur_acc = 0, in_group = 0, groups_done = 0;
  auto drain = [&](int acc) {
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * BN + half * HN;
    if (HN == 32) {
      uint32_t r[32];
      ptx::tmem_ld32(taddr, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) tot[j] += __uint_as_float(r[j]);
    } else {
      uint32_t r[16];
      ptx::tmem_ld16(taddr, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) tot[j] += __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
  };
  // transposing store: element (row, pixel = lane) of a K-major SWIZZLE_128B tile
  auto put = [&](unsigned char* plane, int row, float v) {
    *reinterpret_cast<float*>(plane + (size_t)row * 128 + ((((lane >> 2) ^ (row & 7)) << 4) | ((lane & 3) << 2))) = v;
  };

  auto chunk_body = [&](const int kc, float4 (&curx)[4], float4 (&cury)[BQ / 4]) {
    const int s = kc & 1;
    unsigned char* stage = base + (size_t)s * STAGE;
    if (kc >= 2) mbar_wait(&bar_free[s], ((kc >> 1) - 1) & 1);
    if (warp_rows_ok) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 hi, lo;
        split4(curx[c], hi, lo);
        const int r = warp * 16 + c * 4;
        put(stage, r, hi.x); put(stage, r + 1, hi.y); put(stage, r + 2, hi.z); put(stage, r + 3, hi.w);
        put(stage + kStageA, r, lo.x); put(stage + kStageA, r + 1, lo.y); put(stage + kStageA, r + 2, lo.z); put(stage + kStageA, r + 3, lo.w);
      }
    }
    {
      unsigned char* bh = stage + 2 * kStageA;
#pragma unroll
      for (int c = 0; c < BQ / 4; ++c) {
        float4 hi, lo;
        split4(cury[c], hi, lo);
        const int r = warp * BQ + c * 4;
        put(bh, r, hi.x); put(bh, r + 1, hi.y); put(bh, r + 2, hi.z); put(bh, r + 3, hi.w);
        put(bh + B_BYTES, r, lo.x); put(bh + B_BYTES, r + 1, lo.y); put(bh + B_BYTES, r + 2, lo.z); put(bh + B_BYTES, r + 3, lo.w);
      }
    }
    ptx::fence_proxy_async();                               // see conv_tc_kernel: no loads may be in flight across this fence
    ptx::tc_fence_before();
    __syncthreads();
    if (kc + kPFW < nchunks) load_chunk(curx, cury);
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t ah = (ptx::smem_u32(stage) & 0x3FFFF) >> 4, al = ah + (kStageA >> 4);
        const uint32_t bh = al + (kStageA >> 4), bl = bh + (B_BYTES >> 4);
        const uint32_t d = tmem_base + cur_acc * BN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_tf32(d, al + 2 * k, d_hi, bh + 2 * k, d_hi, IDESC, accf | (uint32_t)k);
          umma_tf32(d, ah + 2 * k, d_hi, bl + 2 * k, d_hi, IDESC, 1u);
          umma_tf32(d, ah + 2 * k, d_hi, bh + 2 * k, d_hi, IDESC, 1u);
        }
        ptx::umma_commit(&bar_free[s]);
        if (in_group + 1 == flush || kc + 1 == nchunks) ptx::umma_commit(&bar_acc[cur_acc]);
      }
      __syncwarp();
    }
    accf = 1;
    if (++in_group == flush || kc + 1 == nchunks) {
      if (groups_done >= 1) {
        const int prev = cur_acc ^ 1;
        mbar_wait(&bar_acc[prev], ((groups_done - 1) >> 1) & 1);
        ptx::tc_fence_after();
        drain(prev);
      }
      ++groups_done;
      cur_acc ^= 1; in_group = 0; accf = 0;
    }
  };
  for (int kc0 = 0; kc0 < nchunks; kc0 += kPFW) {
#pragma unroll
    for (int u = 0; u < kPFW; ++u) {
      if (kc0 + u < nchunks) chunk_body(kc0 + u, prex[u], prey[u]);
    }
  }
  if (groups_done >= 1) {
    const int last = cur_acc ^ 1;
    mbar_wait(&bar_acc[last], ((groups_done - 1) >> 1) & 1);
    ptx::tc_fence_after();
    drain(last);
  }
  // thread (warp, lane) holds row (tap_r, ci_r) = tile row 32*(warp%4) + lane against co0 + half*HN .. + HN: dw[co][ci][tap] += tot
  const int row = mt * 128 + (warp & 3) * 32 + lane;
  if (row < rows_total && nchunks > 0) {
    const int tap_r = row / g.Ci, ci_r = row - tap_r * g.Ci;
#pragma unroll
    for (int j = 0; j < HN; ++j) {
      const int co = co0 + half * HN + j;
      if (co < g.Co) atomicAdd(a.dw + ((long long)co * g.Ci + ci_r) * taps + tap_r, tot[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<TCOLS>(tmem_base);
}

// -------------------------------------------------------------------------------------------------
// halo-resident forward / dgrad / wgrad for the wide, shallow layers (stride 1, 32 -> 32 channels: r1.conv0, r1.conv1,
// r2.conv0 of ResNet8-u32 and their gradients -- 8 of the 11 largest launches of a training step)
//
// The gather-GEMMs above stage one (tap, chunk) slice per block-wide barrier and re-read every source pixel once per tap
// (9x, from L1/L2).  Here a persistent CTA stages, per tile of 128 consecutive positions of the (zero-padded) source grid, the
// 128 + halo source rows ONCE (coalesced 16-byte loads, split into hi / lo planes, row = pixel = 128 bytes, 16-byte piece c
// of row j stored at piece c ^ (j & 7): the canonical SWIZZLE_128B tile); every tap is then the same tile read from a
// different start ROW (an operand may start at any row of a swizzled tile).
//   forward / dgrad: the tile is a K-major A operand (M = 128 pixels, K = 32 channels); per tap and K = 8 step
//       a_hi x [b_hi ; b_lo]  (N = 64: the hi and lo weight planes of a tap are adjacent in smem -> columns [main | small])
//       a_lo x b_hi           (N = 32, accumulated into the `small` columns);
//     all 9 x (32 x 32) weight blocks (72 KB with the lo planes) stay in shared memory for the life of the CTA.
//   wgrad: the tile is an MN-major A operand (M = channels, K = pixels; stored in the one layout 32-bit MN-major operands have,
//     SWIZZLE_128B_BASE32B: 4 pixel rows x 128 bytes per atom, 32-byte pieces XOR-ed with row & 3);
//     the M = 128 rows of one MMA are the four "taps" t = 0..3 of one kernel row r, i.e. four 32-channel atoms whose start
//     rows are dil apart (descriptor MN-atom stride = dil x 128 bytes -- overlapping atoms; t = 3 is a dummy whose result is
//     dropped), B = the dy tile, MN-major as well (N = 64: hi plane | lo plane).  No transposing stores at all.
// Work split: warps 0-7 load / split / store tiles and drain accumulators, warp 8 issues the MMAs; two smem stages and two
// TMEM stages, mbarriers only (no block-wide barrier in the loop): the MMAs of tile i run while tile i-1 is drained /
// written and tile i+1 is loaded.
// Truncating accumulate (see the header): the main term of every `group` taps (wgrad: every tile of 128 pixels) has its own
// TMEM accumulator; the accumulators are summed in registers (round-to-nearest).
// Positions are linear over the source grid Z (forward / wgrad: the input itself; dgrad: dy zero-padded by (k-1)*dil + org),
// so a tile's source rows are consecutive; outputs whose (u, v) fall outside the op's output are computed and dropped
// (6-12 % of the rows for the 25..33-pixel feature maps of the training crops).
// -------------------------------------------------------------------------------------------------
constexpr int kHaloRows = 288;                   // staged source rows per tile: 128 outputs + up to 160 rows of halo
constexpr int kHaloPlane = kHaloRows * 128;      // one plane (hi or lo) of a stage
constexpr int kHaloMaxTaps = 9;
constexpr int kHaloWBytes = kHaloMaxTaps * 8192; // per tap: hi 32 x 128 B | lo 32 x 128 B
constexpr int kHaloSmem = kHaloWBytes + 2 * 2 * kHaloPlane + 1024;
constexpr int kHaloMaxGroups = 4;                // accumulators per TMEM stage, 64 columns each: [main 32 | small 32]
constexpr int kHaloWorkers = 256;                // threads of warps 0-7 (64-channel and first-layer kernels)
constexpr int kHaloThreads = kHaloWorkers + 32;  // + the MMA warp
// The 32-channel kernels run 16 worker warps: their tiles are bounded by the serial latency chain of a worker (global load -> split ->
// swizzled store, then TMEM drain -> global store; tensor pipe 23-43 % busy with 8 workers, profiles/r02_ncu_train_halo.md), so the
// same tile is spread over twice the threads (half the chain per thread, twice the warps to overlap it)
constexpr int kHaloWorkersW = 512;
constexpr int kHaloThreadsW = kHaloWorkersW + 32;
constexpr int kHaloPre = (kHaloRows + 63) / 64;  // 16-byte pieces a wide worker stages per tile

struct HaloArgs {
  const float* src; int N, SH, SW;      // source tensor [N][SH][SW][32]
  int Hz, Wz, pad;                      // virtual source grid: Z[n][u][v] = src[n][u - pad][v - pad], zero outside
  int OH, OW;                           // output [N][OH][OW][32]; position (n, u, v) of Z is an output iff u < OH and v < OW
  int taps, rows, group, contiguous;    // rows staged per tile (multiple of 32); taps per accumulator; Z == src
  int roff[kHaloMaxTaps];               // source row of tap j relative to the output position (Z-linear)
  int wrow[kHaloMaxTaps];               // first row of tap j's 32 x 32 block in the packed weight tensor
  long long lo_rows, total;             // lo-plane row offset of the packed weights; N*Hz*Wz
  int ntiles;
  CUtensorMap tmB;                      // box = 32 rows
  const float* bias; const float* res; int res_H, res_W, res_org, res_stride;
  const float* mask; float* out; int relu, accumulate;
  // wgrad only
  const float* dy; float* dw; float* db; int dil, row0[3];   // row0[r]: source row of tap (r, 0)
  int x_ld4, dy_ld4, ci_blocks, Ci_tot;   // wider layers: blockIdx.y = (32-channel block of ci, 32-channel block of co); row strides in float4
};

// (n, u, v) of a Z-linear position, advanced incrementally (one division pair per tile and thread instead of one per row)
struct ZPos {
  int n, u, v;
  __device__ __forceinline__ void set(long long p, int Hz, int Wz) {
    const unsigned pu = (unsigned)p, hw = (unsigned)(Hz * Wz);
    const unsigned nn = pu / hw, rem = pu - nn * hw, uu = rem / (unsigned)Wz;
    n = (int)nn; u = (int)uu; v = (int)(rem - uu * (unsigned)Wz);
  }
  __device__ __forceinline__ void advance(int d, int Hz, int Wz) {
    v += d;
    while (v >= Wz) { v -= Wz; if (++u == Hz) { u = 0; ++n; } }
  }
};

// low half of an MN-major operand descriptor: start address and LBO (= byte stride between MN atoms)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr, uint32_t atom_stride_bytes) {
  return ((smem_addr & 0x3FFFF) >> 4) | (((atom_stride_bytes >> 4) & 0x3FFF) << 16);
}

template <int MODE>              // 0 forward (bias / residual / ReLU epilogue), 1 data gradient (accumulate / mask epilogue)
__global__ void __launch_bounds__(kHaloThreadsW, 1) conv_halo_tc_kernel(const __grid_constant__ HaloArgs a) {
  constexpr uint32_t IDESC64 = idesc_tf32(128, 64), IDESC32 = idesc_tf32(128, 32);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* const wsm = base;
  unsigned char* const stages = base + kHaloWBytes;
  __shared__ __align__(8) uint64_t bar_w, bar_full[2], bar_mma[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bar_w, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&bar_full[s], kHaloWorkersW); ptx::mbar_init(&bar_mma[s], 1); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tmB);
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int ngroups = (a.taps + a.group - 1) / a.group;
  const int stride_t = gridDim.x;

  if (warp == kHaloWorkersW / 32) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {                                         // the whole weight set, once
      ptx::mbar_expect_tx(&bar_w, (uint32_t)a.taps * 8192u);
      for (int j = 0; j < a.taps; ++j) {
        ptx::tma_load_2d(wsm + j * 8192, &a.tmB, &bar_w, 0, a.wrow[j]);
        ptx::tma_load_2d(wsm + j * 8192 + 4096, &a.tmB, &bar_w, 0, (int)a.lo_rows + a.wrow[j]);
      }
    }
    __syncwarp();
    mbar_wait(&bar_w, 0);
    const uint32_t d_hi = ptx::umma_desc_hi(1024, 2);        // SBO = 8 rows x 128 B, SWIZZLE_128B
    const uint32_t wb = ptx::smem_u32(wsm);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      // every worker has stored tile `it` AND finished reading TMEM stage s of tile it-2 (program order before its arrive)
      mbar_wait(&bar_full[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(stages + (size_t)s * 2 * kHaloPlane);
        const uint32_t d = tmem_base + (uint32_t)s * 256u;
        int g = 0, in_g = 0;
        for (int j = 0; j < a.taps; ++j) {
          const uint32_t ah = ((sa + (uint32_t)a.roff[j] * 128u) & 0x3FFFF) >> 4, al = ah + (kHaloPlane >> 4);
          const uint32_t bw = ((wb + (uint32_t)j * 8192u) & 0x3FFFF) >> 4;
          const uint32_t dg = d + (uint32_t)g * 64u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(dg, ah + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC64, (in_g | k) ? 1u : 0u);   // a_hi x [b_hi; b_lo] -> [main | small]
            umma_tf32(dg + 32u, al + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC32, 1u);               // a_lo x b_hi -> small
          }
          if (++in_g == a.group) { in_g = 0; ++g; }
        }
        ptx::umma_commit(&bar_mma[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ workers: stage tiles, drain and write outputs ------------------------------
    // staging: thread (jr, c) moves 16-byte piece c of rows jr, jr + 64, ...; a warp's load covers 4 rows = 512 contiguous bytes
    const int c = tid & 7, jr = tid >> 3;
    const float4* const src4 = reinterpret_cast<const float4*>(a.src);
    float4 pre[kHaloPre];
    auto load_tile = [&](int tile) {
      const long long p0 = (long long)tile * 128 + jr;
      if (a.contiguous) {
#pragma unroll
        for (int i = 0; i < kHaloPre; ++i) {
          const long long p = p0 + 64 * i;
          pre[i] = (jr + 64 * i < a.rows && p < a.total) ? __ldg(src4 + p * 8 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        ZPos z;
        z.set(p0 < a.total ? p0 : 0, a.Hz, a.Wz);
#pragma unroll
        for (int i = 0; i < kHaloPre; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (jr + 64 * i < a.rows && p0 + 64 * i < a.total) {
            const int sy = z.u - a.pad, sx = z.v - a.pad;
            if (sy >= 0 && sy < a.SH && sx >= 0 && sx < a.SW) v = __ldg(src4 + (((long long)z.n * a.SH + sy) * a.SW + sx) * 8 + c);
          }
          pre[i] = v;
          z.advance(64, a.Hz, a.Wz);
        }
      }
    };
    auto store_tile = [&](unsigned char* stage) {
#pragma unroll
      for (int i = 0; i < kHaloPre; ++i) {
        const int j = jr + 64 * i;
        if (j < a.rows) {
          float4 hi, lo;
          split4(pre[i], hi, lo);
          const int off = j * 128 + ((c ^ (j & 7)) << 4);
          *reinterpret_cast<float4*>(stage + off) = hi;
          *reinterpret_cast<float4*>(stage + kHaloPlane + off) = lo;
        }
      }
    };

    const int row = 32 * (warp & 3) + lane, part = warp >> 2;   // epilogue: thread (row, part) owns 8 channels of one output row
    const unsigned HWz = (unsigned)(a.Hz * a.Wz);
    auto epilogue = [&](int tile, int s, int it) {
      // output coordinates and the epilogue's global operands first: their latency hides behind the wait for the tile's MMAs
      const long long q = (long long)tile * 128 + row;
      bool valid = q < a.total;
      int n = 0, u = 0, v = 0;
      if (valid) {
        const unsigned qu = (unsigned)q;
        const unsigned nn = qu / HWz, rem = qu - nn * HWz;
        n = (int)nn; u = (int)(rem / (unsigned)a.Wz); v = (int)(rem - (unsigned)u * (unsigned)a.Wz);
        valid = u < a.OH && v < a.OW;
      }
      const long long m = ((long long)n * a.OH + u) * a.OW + v;
      const int nb = part * 8;
      float* orow = a.out + m * 32 + nb;
      float4 pr[2], pm[2], pa[2];                              // residual, mask, previous value: 8 channels each
      bool use_res = false;
      if (valid) {
        if (a.res) {
          long long rbase;
          if (MODE == 0) {
            rbase = (((long long)n * a.res_H + (u * a.res_stride + a.res_org)) * a.res_W + (v * a.res_stride + a.res_org)) * 32;
            use_res = true;
          } else {                                             // gradient of the cropped identity skip, embedded at res_org
            const int ry = u - a.res_org, rx = v - a.res_org;
            use_res = ry >= 0 && ry < a.res_H && rx >= 0 && rx < a.res_W;
            rbase = (((long long)n * a.res_H + ry) * a.res_W + rx) * 32;
          }
          if (use_res) {
            pr[0] = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb));
            pr[1] = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb) + 1);
          }
        }
        if (MODE == 1 && a.mask) {
          pm[0] = __ldg(reinterpret_cast<const float4*>(a.mask + m * 32 + nb));
          pm[1] = __ldg(reinterpret_cast<const float4*>(a.mask + m * 32 + nb) + 1);
        }
        if (MODE == 1 && a.accumulate) {
          pa[0] = *reinterpret_cast<const float4*>(orow);
          pa[1] = *(reinterpret_cast<const float4*>(orow) + 1);
        }
      }
      mbar_wait(&bar_mma[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)s * 256u + (uint32_t)part * 8u;
      float acc[8], mainsum[8];
      {
        uint32_t rs[8], rm[8];
        ptx::tmem_ld8(taddr + 32u, rs);                       // group 0: small, main
        ptx::tmem_ld8(taddr, rm);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j] = __uint_as_float(rs[j]); mainsum[j] = __uint_as_float(rm[j]); }
        for (int g = 1; g < ngroups; ++g) {
          ptx::tmem_ld8(taddr + (uint32_t)g * 64u + 32u, rs);
          ptx::tmem_ld8(taddr + (uint32_t)g * 64u, rm);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[j] += __uint_as_float(rs[j]); mainsum[j] += __uint_as_float(rm[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += mainsum[j];
      }
      ptx::tc_fence_before();
      if (!valid) return;
      if (MODE == 0) {
        if (a.bias) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += __ldg(a.bias + nb + e);           // scalar: parameter views need not be 16-byte aligned
        }
        if (use_res) {
          acc[0] += pr[0].x; acc[1] += pr[0].y; acc[2] += pr[0].z; acc[3] += pr[0].w; acc[4] += pr[1].x; acc[5] += pr[1].y; acc[6] += pr[1].z; acc[7] += pr[1].w;
        }
        if (a.relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.f);
        }
      } else {
        if (a.accumulate) {
          acc[0] += pa[0].x; acc[1] += pa[0].y; acc[2] += pa[0].z; acc[3] += pa[0].w; acc[4] += pa[1].x; acc[5] += pa[1].y; acc[6] += pa[1].z; acc[7] += pa[1].w;
        }
        if (use_res) {
          acc[0] += pr[0].x; acc[1] += pr[0].y; acc[2] += pr[0].z; acc[3] += pr[0].w; acc[4] += pr[1].x; acc[5] += pr[1].y; acc[6] += pr[1].z; acc[7] += pr[1].w;
        }
        if (a.mask) {
          acc[0] = pm[0].x > 0.f ? acc[0] : 0.f; acc[1] = pm[0].y > 0.f ? acc[1] : 0.f; acc[2] = pm[0].z > 0.f ? acc[2] : 0.f; acc[3] = pm[0].w > 0.f ? acc[3] : 0.f;
          acc[4] = pm[1].x > 0.f ? acc[4] : 0.f; acc[5] = pm[1].y > 0.f ? acc[5] : 0.f; acc[6] = pm[1].z > 0.f ? acc[6] : 0.f; acc[7] = pm[1].w > 0.f ? acc[7] : 0.f;
        }
      }
      ptx::st_global_256(orow, __float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]),
                         __float_as_uint(acc[4]), __float_as_uint(acc[5]), __float_as_uint(acc[6]), __float_as_uint(acc[7]));
    };

    int it = 0, tile = blockIdx.x;
    if (tile < a.ntiles) load_tile(tile);
    for (; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      // smem stage s was last read by the MMAs of tile it-2, which this thread saw complete in the epilogue of tile it-2
      store_tile(stages + (size_t)s * 2 * kHaloPlane);
      ptx::fence_proxy_async();                               // this thread's stores -> visible to the tensor core's (async-proxy) reads
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_full[s]);                         // every worker arrives for itself (count = kHaloWorkersW)
      if (tile + stride_t < a.ntiles) load_tile(tile + stride_t);     // in flight while the previous tile is written out
      if (it > 0) epilogue(tile - stride_t, s ^ 1, it - 1);
    }
    if (it > 0) epilogue(tile - stride_t, (it - 1) & 1, it - 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// 64 -> 64 channel variant of the forward / dgrad kernel (r3.conv0 / r3.conv1 of ResNet8-u32 and their data gradients: 160-240
// tiles, where the gather-GEMM spends ~3 us per (tap, chunk) on barrier and TMA latency with one CTA per SM).  Same tile scheme;
// a pixel row is two 32-channel chunks (two pairs of hi / lo planes), N = 64, so per tap, chunk and K = 8 step
//     a_hi x [b_hi ; b_lo] (N = 128 -> columns [main 64 | small 64]),  a_lo x b_hi (N = 64 -> small).
// The weights (9 taps x 32 KB with the lo planes) do not fit beside the tile: warp 9 streams them per tap through a 3-slot TMA
// ring, every tile.  One smem stage and one TMEM stage (a CTA sees one or two tiles).
constexpr int kH64Rows = 256;                      // 128 outputs + up to 128 rows of halo (31 x 31 maps with dilation 2, 33 x 33 with 1)
constexpr int kH64Plane = kH64Rows * 128;
constexpr int kH64Stage = 4 * kH64Plane;              // chunk 0 hi | chunk 0 lo | chunk 1 hi | chunk 1 lo
constexpr int kH64WSlot = 32768;                      // one tap: chunk 0 (hi 64 x 128 B | lo) | chunk 1 (hi | lo)
constexpr int kH64Smem = kH64Stage + 3 * kH64WSlot + 1024;
constexpr int kH64Threads = 256 + 64;

template <int MODE>
__global__ void __launch_bounds__(kH64Threads, 1) conv_halo64_tc_kernel(const __grid_constant__ HaloArgs a) {
  constexpr uint32_t IDESC128 = idesc_tf32(128, 128), IDESC64 = idesc_tf32(128, 64);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* const stage = base;
  unsigned char* const wring = base + kH64Stage;
  __shared__ __align__(8) uint64_t bar_full, bar_mma, bar_wfull[3], bar_wfree[3];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bar_full, 256); ptx::mbar_init(&bar_mma, 1);
    for (int i = 0; i < 3; ++i) { ptx::mbar_init(&bar_wfull[i], 1); ptx::mbar_init(&bar_wfree[i], 1); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tmB);
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int ngroups = (a.taps + a.group - 1) / a.group;
  const int stride_t = gridDim.x;
  const int my_tiles = (int)blockIdx.x < a.ntiles ? (a.ntiles - (int)blockIdx.x + stride_t - 1) / stride_t : 0;

  if (warp == 9) {
    // ------------------------------ weight producer: one tap (32 KB) per ring slot ------------------------------
    if (lane == 0) {
      const int total = my_tiles * a.taps;
      for (int q = 0; q < total; ++q) {
        const int slot = q % 3, j = q % a.taps;
        if (q >= 3) mbar_wait(&bar_wfree[slot], (uint32_t)((q / 3) - 1) & 1u);
        unsigned char* w = wring + (size_t)slot * kH64WSlot;
        ptx::mbar_expect_tx(&bar_wfull[slot], (uint32_t)kH64WSlot);
        for (int c = 0; c < 2; ++c) {
          ptx::tma_load_2d(w + c * 16384, &a.tmB, &bar_wfull[slot], 0, a.wrow[j] + c * 64);
          ptx::tma_load_2d(w + c * 16384 + 8192, &a.tmB, &bar_wfull[slot], 0, (int)a.lo_rows + a.wrow[j] + c * 64);
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t d_hi = ptx::umma_desc_hi(1024, 2);
    const uint32_t sa = ptx::smem_u32(stage), wb = ptx::smem_u32(wring);
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(&bar_full, (uint32_t)it & 1u);               // tile stored; every worker is done with the previous tile's TMEM
      ptx::tc_fence_after();
      int g = 0, in_g = 0;
      for (int j = 0; j < a.taps; ++j) {
        const int q = it * a.taps + j, slot = q % 3;
        mbar_wait(&bar_wfull[slot], (uint32_t)(q / 3) & 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t dg = tmem_base + (uint32_t)g * 128u;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t ah = ((sa + (uint32_t)c * 2u * kH64Plane + (uint32_t)a.roff[j] * 128u) & 0x3FFFF) >> 4, al = ah + (kH64Plane >> 4);
            const uint32_t bw = ((wb + (uint32_t)slot * kH64WSlot + (uint32_t)c * 16384u) & 0x3FFFF) >> 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_tf32(dg, ah + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC128, (in_g | c | k) ? 1u : 0u);   // a_hi x [b_hi; b_lo]
              umma_tf32(dg + 64u, al + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC64, 1u);                    // a_lo x b_hi
            }
          }
          ptx::umma_commit(&bar_wfree[slot]);
          if (j + 1 == a.taps) ptx::umma_commit(&bar_mma);
        }
        __syncwarp();
        if (++in_g == a.group) { in_g = 0; ++g; }
      }
    }
  } else {
    // ------------------------------ workers ------------------------------
    // staging: thread (jr, pc) moves 16-byte piece pc (0..15: chunk pc >> 3) of rows jr, jr + 16, ...
    const int pc = tid & 15, jr = tid >> 4;
    const float4* const src4 = reinterpret_cast<const float4*>(a.src);
    float4 pre[kH64Rows / 16];
    auto load_tile = [&](int tile) {
      const long long p0 = (long long)tile * 128 + jr;
      if (a.contiguous) {
#pragma unroll
        for (int i = 0; i < kH64Rows / 16; ++i) {
          const long long p = p0 + 16 * i;
          pre[i] = (16 * i < a.rows && p < a.total) ? __ldg(src4 + p * 16 + pc) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        ZPos z;
        z.set(p0 < a.total ? p0 : 0, a.Hz, a.Wz);
#pragma unroll
        for (int i = 0; i < kH64Rows / 16; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (16 * i < a.rows && p0 + 16 * i < a.total) {
            const int sy = z.u - a.pad, sx = z.v - a.pad;
            if (sy >= 0 && sy < a.SH && sx >= 0 && sx < a.SW) v = __ldg(src4 + (((long long)z.n * a.SH + sy) * a.SW + sx) * 16 + pc);
          }
          pre[i] = v;
          z.advance(16, a.Hz, a.Wz);
        }
      }
    };
    auto store_tile = [&]() {
      unsigned char* plane = stage + (size_t)(pc >> 3) * 2 * kH64Plane;
#pragma unroll
      for (int i = 0; i < kH64Rows / 16; ++i) {
        const int j = jr + 16 * i;
        if (16 * i < a.rows) {
          float4 hi, lo;
          split4(pre[i], hi, lo);
          const int off = j * 128 + (((pc & 7) ^ (j & 7)) << 4);
          *reinterpret_cast<float4*>(plane + off) = hi;
          *reinterpret_cast<float4*>(plane + kH64Plane + off) = lo;
        }
      }
    };
    const int row = 32 * (warp & 3) + lane, half = warp >> 2;   // epilogue: thread (row, half) owns 32 channels of one output row
    const unsigned HWz = (unsigned)(a.Hz * a.Wz);
    auto epilogue = [&](int tile, int it) {
      mbar_wait(&bar_mma, (uint32_t)it & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)half * 32u;
      float acc[32];
      {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + 64u, r);                        // small terms first
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
        for (int g = 1; g < ngroups; ++g) {
          ptx::tmem_ld32(taddr + (uint32_t)g * 128u + 64u, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        }
        for (int g = 0; g < ngroups; ++g) {
          ptx::tmem_ld32(taddr + (uint32_t)g * 128u, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        }
      }
      ptx::tc_fence_before();
      const long long q = (long long)tile * 128 + row;
      if (q >= a.total) return;
      const unsigned qu = (unsigned)q;
      const unsigned n = qu / HWz, rem = qu - n * HWz;
      const int u = (int)(rem / (unsigned)a.Wz), v = (int)(rem - (unsigned)u * (unsigned)a.Wz);
      if (u >= a.OH || v >= a.OW) return;
      const long long m = ((long long)n * a.OH + u) * a.OW + v;
      const int nb = half * 32;
      float* orow = a.out + m * 64 + nb;
      long long rbase = 0;
      bool use_res = false;
      if (a.res) {
        if (MODE == 0) {
          rbase = (((long long)n * a.res_H + (u * a.res_stride + a.res_org)) * a.res_W + (v * a.res_stride + a.res_org)) * 64;
          use_res = true;
        } else {
          const int ry = u - a.res_org, rx = v - a.res_org;
          use_res = ry >= 0 && ry < a.res_H && rx >= 0 && rx < a.res_W;
          rbase = (((long long)n * a.res_H + ry) * a.res_W + rx) * 64;
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < 32; c8 += 8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = acc[c8 + j];
        if (MODE == 0 && a.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += __ldg(a.bias + nb + c8 + j);
        }
        if (MODE == 1 && a.accumulate) {
          const float4 o0 = *reinterpret_cast<const float4*>(orow + c8), o1 = *reinterpret_cast<const float4*>(orow + c8 + 4);
          o[0] += o0.x; o[1] += o0.y; o[2] += o0.z; o[3] += o0.w; o[4] += o1.x; o[5] += o1.y; o[6] += o1.z; o[7] += o1.w;
        }
        if (use_res) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb + c8));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(a.res + rbase + nb + c8 + 4));
          o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w; o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
        }
        if (MODE == 0 && a.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (MODE == 1 && a.mask) {
          const float4 k0 = __ldg(reinterpret_cast<const float4*>(a.mask + m * 64 + nb + c8));
          const float4 k1 = __ldg(reinterpret_cast<const float4*>(a.mask + m * 64 + nb + c8 + 4));
          o[0] = k0.x > 0.f ? o[0] : 0.f; o[1] = k0.y > 0.f ? o[1] : 0.f; o[2] = k0.z > 0.f ? o[2] : 0.f; o[3] = k0.w > 0.f ? o[3] : 0.f;
          o[4] = k1.x > 0.f ? o[4] : 0.f; o[5] = k1.y > 0.f ? o[5] : 0.f; o[6] = k1.z > 0.f ? o[6] : 0.f; o[7] = k1.w > 0.f ? o[7] : 0.f;
        }
        ptx::st_global_256(orow + c8, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]),
                           __float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7]));
      }
    };

    int tile = blockIdx.x;
    if (my_tiles > 0) load_tile(tile);
    for (int it = 0; it < my_tiles; ++it, tile += stride_t) {
      if (it > 0) epilogue(tile - stride_t, it - 1);          // also: the previous tile's MMAs are done reading the smem stage
      store_tile();
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_full);
      if (it + 1 < my_tiles) load_tile(tile + stride_t);
    }
    if (my_tiles > 0) epilogue(tile - stride_t, my_tiles - 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// wgrad on the same tiles: dw[co][ci][r][t] += sum over positions p of x[p + roff(r, t)][ci] * dy[p][co]
constexpr int kWgDyPlane = 128 * 128;                          // dy tile: 128 positions x 32 channels, one plane
constexpr int kWgStage = 2 * kHaloPlane + 2 * kWgDyPlane;      // x hi | x lo | dy hi | dy lo
constexpr int kWgSmem = 2 * kWgStage + 1024;

__global__ void __launch_bounds__(kHaloThreadsW, 1) wgrad_halo_tc_kernel(const __grid_constant__ HaloArgs a) {
  // MN-major A and B (bits 15, 16)
  constexpr uint32_t IDESC64 = idesc_tf32(128, 64) | (1u << 15) | (1u << 16), IDESC32 = idesc_tf32(128, 32) | (1u << 15) | (1u << 16);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full[2], bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_db[32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 32) s_db[tid] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&bar_full[s], kHaloWorkersW); ptx::mbar_init(&bar_mma[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int stride_t = gridDim.x;

  if (warp == kHaloWorkersW / 32) {
    // MN-major 32-bit operands exist in ONE shared-memory layout, SWIZZLE_128B_BASE32B (layout type 1): atom = 4 K-rows (pixels) x
    // 128 bytes, 32-byte piece q of row p stored at piece q ^ (p & 3); LBO = stride between MN atoms, SBO = stride between the
    // two K atoms of a K = 8 step (4 rows = 512 bytes).  A's MN atoms are the taps t = 0..3: dil rows apart (overlapping atoms).
    const uint32_t d_hi_a = ptx::umma_desc_hi(512u, 1);
    const uint32_t d_hi_b = ptx::umma_desc_hi(512u, 1);                        // N atoms (LBO): hi plane, lo plane
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      mbar_wait(&bar_full[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t xs = ptx::smem_u32(base + (size_t)s * kWgStage);
        const uint32_t ys = xs + 2 * kHaloPlane;
        const uint32_t d = tmem_base + (uint32_t)s * 256u;
        for (int ks = 0; ks < 16; ++ks) {                     // 8 positions per MMA
          const uint32_t b_lo = desc_lo_mn(ys + (uint32_t)ks * 1024u, kWgDyPlane);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const uint32_t xr = xs + ((uint32_t)a.row0[r] + 8u * ks) * 128u;
            umma_tf32(d + r * 64u, desc_lo_mn(xr, (uint32_t)a.dil * 128u), d_hi_a, b_lo, d_hi_b, IDESC64, ks ? 1u : 0u);   // x_hi x [dy_hi | dy_lo]
            umma_tf32(d + r * 64u + 32u, desc_lo_mn(xr + kHaloPlane, (uint32_t)a.dil * 128u), d_hi_a, b_lo, d_hi_b, IDESC32, 1u);  // x_lo x dy_hi
          }
        }
        ptx::umma_commit(&bar_mma[s]);
      }
      __syncwarp();
    }
  } else {
    const int c = tid & 7, jr = tid >> 3;                     // 16 worker warps: piece c of rows jr, jr + 64, ...
    const int cib = (int)blockIdx.y % a.ci_blocks, cob = (int)blockIdx.y / a.ci_blocks;   // this CTA's 32 x 32 block of dw
    const float4* const x4 = reinterpret_cast<const float4*>(a.src) + cib * 8;
    const float4* const dy4 = reinterpret_cast<const float4*>(a.dy) + cob * 8;
    float4 prex[kHaloPre], prey[2];
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);            // bias gradient: column sums of the dy pieces this thread stages
    auto load_tile = [&](int tile) {
      const long long p0 = (long long)tile * 128 + jr;
#pragma unroll
      for (int i = 0; i < kHaloPre; ++i) {
        const long long p = p0 + 64 * i;
        prex[i] = (jr + 64 * i < a.rows && p < a.total) ? __ldg(x4 + p * a.x_ld4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ZPos z;
      z.set(p0 < a.total ? p0 : 0, a.Hz, a.Wz);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p0 + 64 * i < a.total && z.u < a.OH && z.v < a.OW) v = __ldg(dy4 + (((long long)z.n * a.OH + z.u) * a.OW + z.v) * a.dy_ld4 + c);
        prey[i] = v;
        z.advance(64, a.Hz, a.Wz);
      }
    };
    auto store_tile = [&](unsigned char* stage) {
#pragma unroll
      for (int i = 0; i < kHaloPre; ++i) {
        const int j = jr + 64 * i;
        if (j < a.rows) {
          float4 hi, lo;
          split4(prex[i], hi, lo);
          const int off = j * 128 + ((c ^ ((j & 3) << 1)) << 4);     // 32-byte piece (c >> 1) ^= j & 3
          *reinterpret_cast<float4*>(stage + off) = hi;
          *reinterpret_cast<float4*>(stage + kHaloPlane + off) = lo;
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int j = jr + 64 * i;
        float4 hi, lo;
        split4(prey[i], hi, lo);
        bsum.x += prey[i].x; bsum.y += prey[i].y; bsum.z += prey[i].z; bsum.w += prey[i].w;
        const int off = j * 128 + ((c ^ ((j & 3) << 1)) << 4);
        *reinterpret_cast<float4*>(stage + 2 * kHaloPlane + off) = hi;
        *reinterpret_cast<float4*>(stage + 2 * kHaloPlane + kWgDyPlane + off) = lo;
      }
    };
    // accumulator lane m = 32*t + ci (t = tap column, 3 = dummy), columns = co; thread (m, part) keeps co part*8 .. +8 of the 3 rows r
    const int part = warp >> 2;
    float tot[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) tot[r][j] = 0.f;
    auto drain = [&](int s, int it) {
      mbar_wait(&bar_mma[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)s * 256u + (uint32_t)part * 8u;
      uint32_t rs[3][8], rm[3][8];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        ptx::tmem_ld8(taddr + r * 64u + 32u, rs[r]);
        ptx::tmem_ld8(taddr + r * 64u, rm[r]);
      }
      ptx::tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[r][j] += __uint_as_float(rs[r][j]) + __uint_as_float(rm[r][j]);
      ptx::tc_fence_before();
    };
    int it = 0, tile = blockIdx.x;
    if (tile < a.ntiles) load_tile(tile);
    for (; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      store_tile(base + (size_t)s * kWgStage);
      ptx::fence_proxy_async();                               // this thread's stores -> visible to the tensor core's (async-proxy) reads
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_full[s]);                         // every worker arrives for itself (count = kHaloWorkersW)
      if (tile + stride_t < a.ntiles) load_tile(tile + stride_t);
      if (it > 0) drain(s ^ 1, it - 1);
    }
    if (it > 0) drain((it - 1) & 1, it - 1);
    const int m = 32 * (warp & 3) + lane, t = m >> 5, ci = m & 31;
    if (it > 0 && t < 3) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.dw + ((long long)(cob * 32 + part * 8 + j) * a.Ci_tot + cib * 32 + ci) * 9 + r * 3 + t, tot[r][j]);
    }
    if (a.db && cib == 0) {                                  // one ci block per co block reports the bias gradient
      atomicAdd(&s_db[4 * c], bsum.x); atomicAdd(&s_db[4 * c + 1], bsum.y); atomicAdd(&s_db[4 * c + 2], bsum.z); atomicAdd(&s_db[4 * c + 3], bsum.w);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (a.db && tid < 32 && (int)blockIdx.y % a.ci_blocks == 0) atomicAdd(a.db + ((int)blockIdx.y / a.ci_blocks) * 32 + tid, s_db[tid]);
  if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// -------------------------------------------------------------------------------------------------
// Cin = 1 first layer (7 x 7, stride 2 on the training crops; resnet.py:294, basic.py:47) on the tensor core.
// The im2col row of an output pixel (49 taps, padded to 64 = two 32-tap chunks) is built directly in shared memory, split hi / lo,
// in the operand layout the MMA wants -- K-major SWIZZLE_128B for the forward (M = pixels), MN-major SWIZZLE_128B_BASE32B for the
// weight gradient (K = pixels).  Tiles are 128 consecutive output pixels, so y / dy rows are contiguous.  Same pipeline as the
// halo kernels: warps 0-7 stage and drain, warp 8 issues, two smem and two TMEM stages.
//   forward: per 32-tap chunk and K = 8 step  x_hi x [w_hi ; w_lo] (N = 64 -> [main | small]) and x_lo x w_hi (N = 32 -> small);
//            14 MMAs per tile (the last K step of chunk 1 is all padding).
//   wgrad:   ONE MMA per 8 pixels: the four MN atoms of A are [col_hi c0 | col_hi c1 | col_lo c0 | col_lo c1] (M = 128) and
//            B = [dy_hi | dy_lo] (N = 64), so lanes 0-63 hold col_hi x dy_hi | col_hi x dy_lo and lanes 64-127 hold col_lo x dy_hi
//            (| col_lo x dy_lo, dropped): all three terms of the compensated product from one instruction.
// -------------------------------------------------------------------------------------------------
struct FirstArgs {
  const float* x; int N, H, W;          // [N][H][W], one channel
  int stride, Ho, Wo;
  const float* w; const float* bias; int relu; float* y;      // forward: w [32][49] (OIHW, Ci = 1)
  const float* dy; float* dw; float* db;                        // wgrad (db: bias gradient, may be NULL)
  long long M; int ntiles;
};
constexpr int kFirstK = 7, kFirstTaps = 49;
constexpr int kFirstPlane = 128 * 128;                          // 128 pixels x 32 taps
constexpr int kFirstFwdSmem = 16384 + 2 * 4 * kFirstPlane + 1024;
constexpr int kFirstWgSmem = 2 * 6 * kFirstPlane + 1024;

// the 32 taps of chunk C of one output pixel (taps beyond 49: zero)
template <int C>
__device__ __forceinline__ void first_load_taps(const float* __restrict__ xb, int W, bool ok, float (&v)[32]) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const int tap = C * 32 + e;
    v[e] = (ok && tap < kFirstTaps) ? __ldg(xb + (tap / kFirstK) * W + (tap % kFirstK)) : 0.f;
  }
}

template <int WGRAD>
__global__ void __launch_bounds__(kHaloThreads, 1) first_tc_train_kernel(const __grid_constant__ FirstArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // forward: [weights 16 KB: c0 hi | c0 lo | c1 hi | c1 lo (32 rows x 128 B each)] [stage: c0 hi | c0 lo | c1 hi | c1 lo] x 2
  // wgrad:   [stage: col_hi c0 | col_hi c1 | col_lo c0 | col_lo c1 | dy_hi | dy_lo] x 2
  unsigned char* const wsm = base;
  unsigned char* const stages = WGRAD ? base : base + 16384;
  constexpr int STAGE = WGRAD ? 6 * kFirstPlane : 4 * kFirstPlane;
  __shared__ __align__(8) uint64_t bar_full[2], bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_db[32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 32) s_db[tid] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&bar_full[s], kHaloWorkers); ptx::mbar_init(&bar_mma[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<128>(&tmem_base_s);
  if (!WGRAD) {                                               // weights -> K-major B operand, split hi / lo
    for (int i = tid; i < 32 * 64; i += kHaloThreads) {
      const int co = i >> 6, tap = i & 63;
      float hi = 0.f, lo = 0.f;
      if (tap < kFirstTaps) split1(__ldg(a.w + co * kFirstTaps + tap), hi, lo);
      const int off = (tap >> 5) * 8192 + co * 128 + (((((tap & 31) >> 2) ^ (co & 7)) << 4) | ((tap & 3) << 2));
      *reinterpret_cast<float*>(wsm + off) = hi;
      *reinterpret_cast<float*>(wsm + 4096 + off) = lo;
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int stride_t = gridDim.x;

  if (warp == 8) {
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      mbar_wait(&bar_full[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(stages + (size_t)s * STAGE);
        const uint32_t d = tmem_base + (uint32_t)s * 64u;
        if (WGRAD) {
          constexpr uint32_t IDESC = idesc_tf32(128, 64) | (1u << 15) | (1u << 16);
          const uint32_t d_hi = ptx::umma_desc_hi(512u, 1);
#pragma unroll 4
          for (int ks = 0; ks < 16; ++ks)
            umma_tf32(d, desc_lo_mn(sa + (uint32_t)ks * 1024u, kFirstPlane), d_hi, desc_lo_mn(sa + 4u * kFirstPlane + (uint32_t)ks * 1024u, kFirstPlane),
                      d_hi, IDESC, ks ? 1u : 0u);
        } else {
          constexpr uint32_t IDESC64 = idesc_tf32(128, 64), IDESC32 = idesc_tf32(128, 32);
          const uint32_t d_hi = ptx::umma_desc_hi(1024, 2);
          const uint32_t wb = ptx::smem_u32(wsm);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t ah = ((sa + (uint32_t)c * 2u * kFirstPlane) & 0x3FFFF) >> 4, al = ah + (kFirstPlane >> 4);
            const uint32_t bw = ((wb + (uint32_t)c * 8192u) & 0x3FFFF) >> 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (c == 1 && k == 3) continue;                 // taps 56..63: padding
              umma_tf32(d, ah + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC64, (c | k) ? 1u : 0u);
              umma_tf32(d + 32u, al + 2 * k, d_hi, bw + 2 * k, d_hi, IDESC32, 1u);
            }
          }
        }
        ptx::umma_commit(&bar_mma[s]);
      }
      __syncwarp();
    }
  } else {
    // im2col: thread (row, c) builds the 32 taps of chunk c of pixel `row`
    const int row = tid & 127, c = tid >> 7;
    float col[32];
    float4 prey[4];
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);            // bias gradient: column sums of the dy pieces this thread stages
    const int pcy = tid & 7, jry = tid >> 3;                  // dy staging (wgrad): 16-byte piece pcy of rows jry + 32 i
    auto load_tile = [&](int tile) {
      const long long m = (long long)tile * 128 + row;
      const bool ok = m < a.M;
      const long long mm = ok ? m : 0;
      const int ox = (int)(mm % a.Wo); const long long q = mm / a.Wo; const int oy = (int)(q % a.Ho); const int n = (int)(q / a.Ho);
      const float* xb = a.x + ((long long)n * a.H + oy * a.stride) * a.W + ox * a.stride;
      if (c == 0) first_load_taps<0>(xb, a.W, ok, col); else first_load_taps<1>(xb, a.W, ok, col);
      if (WGRAD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long p = (long long)tile * 128 + jry + 32 * i;
          prey[i] = p < a.M ? __ldg(reinterpret_cast<const float4*>(a.dy) + p * 8 + pcy) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto store_tile = [&](unsigned char* stage) {
      // forward: planes c0 hi, c0 lo, c1 hi, c1 lo (16-byte piece q at q ^ (row & 7)); wgrad: hi c0, hi c1, lo c0, lo c1 (32-byte piece
      // q >> 1 at (q >> 1) ^ (row & 3))
      unsigned char* ph = stage + (size_t)(WGRAD ? c : 2 * c) * kFirstPlane + row * 128;
      unsigned char* pl = ph + (size_t)(WGRAD ? 2 : 1) * kFirstPlane;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = make_float4(col[4 * q], col[4 * q + 1], col[4 * q + 2], col[4 * q + 3]), hi, lo;
        split4(v, hi, lo);
        const int piece = WGRAD ? (q ^ ((row & 3) << 1)) : (q ^ (row & 7));
        *reinterpret_cast<float4*>(ph + (piece << 4)) = hi;
        *reinterpret_cast<float4*>(pl + (piece << 4)) = lo;
      }
      if (WGRAD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = jry + 32 * i;
          float4 hi, lo;
          split4(prey[i], hi, lo);
          bsum.x += prey[i].x; bsum.y += prey[i].y; bsum.z += prey[i].z; bsum.w += prey[i].w;
          const int off = j * 128 + ((pcy ^ ((j & 3) << 1)) << 4);
          *reinterpret_cast<float4*>(stage + 4 * kFirstPlane + off) = hi;
          *reinterpret_cast<float4*>(stage + 5 * kFirstPlane + off) = lo;
        }
      }
    };
    const int erow = 32 * (warp & 3) + lane, half = warp >> 2;
    float tot[32];                                            // wgrad: running sums of this thread's 32 accumulator columns
#pragma unroll
    for (int j = 0; j < 32; ++j) tot[j] = 0.f;
    auto finish = [&](int tile, int s, int it) {
      mbar_wait(&bar_mma[s], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)s * 64u;
      if (WGRAD) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + (uint32_t)half * 32u, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) tot[j] += __uint_as_float(r[j]);
        ptx::tc_fence_before();
      } else {
        uint32_t rm[16], rs[16];
        ptx::tmem_ld16(taddr + 32u + (uint32_t)half * 16u, rs);
        ptx::tmem_ld16(taddr + (uint32_t)half * 16u, rm);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        const long long m = (long long)tile * 128 + erow;
        if (m >= a.M) return;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          o[j] = __uint_as_float(rs[j]) + __uint_as_float(rm[j]) + (a.bias ? __ldg(a.bias + half * 16 + j) : 0.f);
          if (a.relu) o[j] = fmaxf(o[j], 0.f);
        }
        float* orow = a.y + m * 32 + half * 16;
        ptx::st_global_256(orow, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]),
                           __float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7]));
        ptx::st_global_256(orow + 8, __float_as_uint(o[8]), __float_as_uint(o[9]), __float_as_uint(o[10]), __float_as_uint(o[11]),
                           __float_as_uint(o[12]), __float_as_uint(o[13]), __float_as_uint(o[14]), __float_as_uint(o[15]));
      }
    };
    int it = 0, tile = blockIdx.x;
    if (tile < a.ntiles) load_tile(tile);
    for (; tile < a.ntiles; tile += stride_t, ++it) {
      const int s = it & 1;
      store_tile(stages + (size_t)s * STAGE);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_full[s]);
      if (tile + stride_t < a.ntiles) load_tile(tile + stride_t);
      if (it > 0) finish(tile - stride_t, s ^ 1, it - 1);
    }
    if (it > 0) finish(tile - stride_t, (it - 1) & 1, it - 1);
    if (WGRAD && it > 0) {
      // lanes 0-63: tap = lane, columns [col_hi x dy_hi | col_hi x dy_lo]; lanes 64-127: tap = lane - 64, [col_lo x dy_hi | dropped]
      const int tap = erow & 63;
      if (tap < kFirstTaps && !(erow >= 64 && half == 1)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(a.dw + j * kFirstTaps + tap, tot[j]);
      }
    }
    if (WGRAD && a.db) {
      atomicAdd(&s_db[4 * pcy], bsum.x); atomicAdd(&s_db[4 * pcy + 1], bsum.y); atomicAdd(&s_db[4 * pcy + 2], bsum.z); atomicAdd(&s_db[4 * pcy + 3], bsum.w);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (WGRAD && a.db && tid < 32) atomicAdd(a.db + tid, s_db[tid]);
  if (warp == 0) ptx::tmem_dealloc<128>(tmem_base);
}

int g_flush = -1;
int flush_chunks() {
  if (g_flush < 0) {
    const char* e = getenv("TPZ_TRAIN_FLUSH");
    g_flush = e ? atoi(e) : 2;             // chunks (of 32 K-elements x 3 passes) per accumulator before the round-to-nearest drain
  }
  return g_flush;
}

TGeom tgeom(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int kh, int kw, int stride, int dil, int org) {
  TGeom g; g.N = N; g.H = H; g.W = W; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.kh = kh; g.kw = kw;
  g.stride = stride; g.dil = dil; g.org = org; return g;
}

template <int BN, int MODE>
int launch_conv_tc(const FwdArgs& a, long long M, int Nn, cudaStream_t stream, int grid_y = 0, int grid_z = 1) {
  const int smem = 2 * (2 * kStageA) + 3 * (2 * BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(tpz_div_up(M, 128), grid_y > 0 ? grid_y : Nn / BN, grid_z);
  conv_tc_kernel<BN, MODE><<<grid, 256, smem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// TMA view of a packed weight tensor: rows of 32 fp32, declared as 64 fp16 (same bytes, same 128-byte swizzle)
int weight_tmap(CUtensorMap* tm, const float* packed, long long rows, int box_rows) {
  uint64_t dims[2] = {64, (uint64_t)rows};
  uint64_t strides[1] = {128};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  uint32_t es[2] = {1, 1};
  return tpz_encode_tmap(tm, packed, 2, dims, strides, box, es, 128);
}


int g_halo = -1;
bool halo_enabled() {
  if (g_halo < 0) {
    const char* e = getenv("TPZ_TRAIN_HALO");
    g_halo = e ? atoi(e) : 1;
  }
  return g_halo != 0;
}
int g_halo_wg = -1;
bool halo_wgrad_enabled() {                                  // TPZ_TRAIN_HALO_WGRAD=0: keep the transposing wgrad kernel
  if (g_halo_wg < 0) {
    const char* e = getenv("TPZ_TRAIN_HALO_WGRAD");
    g_halo_wg = e ? atoi(e) : 1;
  }
  return g_halo_wg != 0;
}
int g_lat_dg = -1;
bool lattice_dgrad_enabled() {                               // TPZ_TRAIN_LATTICE_DGRAD=0: strided dgrad over every position
  if (g_lat_dg < 0) {
    const char* e = getenv("TPZ_TRAIN_LATTICE_DGRAD");
    g_lat_dg = e ? atoi(e) : 1;
  }
  return g_lat_dg != 0;
}
int g_sc_dg = -1;
bool scatter_dgrad_enabled() {                               // TPZ_TRAIN_SCATTER_DGRAD=0: generic dgrad for one-pixel outputs too
  if (g_sc_dg < 0) {
    const char* e = getenv("TPZ_TRAIN_SCATTER_DGRAD");
    g_sc_dg = e ? atoi(e) : 1;
  }
  return g_sc_dg != 0;
}
int g_ksplit = -1;
bool ksplit_enabled() {                                      // TPZ_TRAIN_KSPLIT=0: no split-K forward
  if (g_ksplit < 0) {
    const char* e = getenv("TPZ_TRAIN_KSPLIT");
    g_ksplit = e ? atoi(e) : 1;
  }
  return g_ksplit != 0;
}
// y[m][c] = act(y[m][c] + bias[c]) after a split-K forward
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n4, int C, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(y)[i];
    if (bias) {                                              // scalar loads: parameter views need not be 16-byte aligned
      const int c = (int)((i * 4) % C);
      v.x += __ldg(bias + c); v.y += __ldg(bias + c + 1); v.z += __ldg(bias + c + 2); v.w += __ldg(bias + c + 3);
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(y)[i] = v;
  }
}
int g_sms = 0;
int sm_count() {
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

// Fills the geometry of the halo-resident kernel; false when the layer does not fit it (caller falls back to the gather-GEMM).
// mode 0: forward of x [N][H][W][32] -> y [N][Ho][Wo][32]; mode 1: data gradient dy [N][Ho][Wo][32] -> dx [N][H][W][32].
bool halo_geometry(HaloArgs& a, int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int kh, int kw, int stride, int dil,
                   int org, int C = 32) {
  if (!halo_enabled() || stride != 1 || kh != kw || kh * kw > kHaloMaxTaps) return false;
  if (mode == 2 ? (Ci % 32 != 0 || Co % 32 != 0) : (Ci != C || Co != C)) return false;   // wgrad: any multiple of 32 (32 x 32 blocks of dw)
  const int chunks = C / 32, max_rows = C == 32 ? kHaloRows : kH64Rows;
  const int k = kh, span = (k - 1) * dil;
  a.N = N; a.taps = k * k;
  if (mode == 0 || mode == 2) {
    a.SH = H; a.SW = W; a.OH = Ho; a.OW = Wo;
    a.pad = org < 0 ? -org : 0;
    const int need_h = Ho + span + org + a.pad, need_w = Wo + span + org + a.pad;
    a.Hz = H + 2 * a.pad > need_h ? H + 2 * a.pad : need_h;
    a.Wz = W + 2 * a.pad > need_w ? W + 2 * a.pad : need_w;
    for (int r = 0; r < k; ++r)
      for (int t = 0; t < k; ++t) {
        a.roff[r * k + t] = (r * dil + org + a.pad) * a.Wz + (t * dil + org + a.pad);
        a.wrow[r * k + t] = (r * k + t) * chunks * C;
      }
  } else {
    a.SH = Ho; a.SW = Wo; a.OH = H; a.OW = W;
    a.pad = org + span;
    if (a.pad < 0) return false;
    a.Hz = H + span; a.Wz = W + span;
    for (int r = 0; r < k; ++r)
      for (int t = 0; t < k; ++t) {
        a.roff[r * k + t] = ((k - 1 - r) * dil) * a.Wz + (k - 1 - t) * dil;
        a.wrow[r * k + t] = (r * k + t) * chunks * C;
      }
  }
  int mx = 0;
  for (int j = 0; j < a.taps; ++j) {
    if (a.roff[j] < 0) return false;
    if (a.roff[j] > mx) mx = a.roff[j];
  }
  if (mode == 2) {                                          // wgrad: M = 128 rows = taps t = 0..3 (t = 3 dummy) of one kernel row
    if (k != 3) return false;
    a.dil = dil;
    for (int r = 0; r < 3; ++r) a.row0[r] = a.roff[r * 3];
    mx = a.row0[2] + 3 * dil;
    a.x_ld4 = Ci / 4; a.dy_ld4 = Co / 4; a.ci_blocks = Ci / 32; a.Ci_tot = Ci;
  }
  a.rows = (128 + mx + 31) / 32 * 32;
  if (a.rows > max_rows) return false;
  a.total = (long long)N * a.Hz * a.Wz;
  if (a.total >= (1ll << 31) - 512) return false;
  a.ntiles = (int)((a.total + 127) / 128);
  a.contiguous = a.pad == 0 && a.Hz == a.SH && a.Wz == a.SW;
  int group = flush_chunks() / chunks;                      // taps per main accumulator (`chunks` 32-channel chunks per tap)
  if (flush_chunks() <= 0) group = a.taps;
  if (group < 1) group = 1;
  while ((a.taps + group - 1) / group > (C == 32 ? kHaloMaxGroups : 3)) ++group;
  a.group = group;
  return true;
}

template <int MODE>
int launch_halo(const HaloArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(conv_halo_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem));
    configured = true;
  }
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  conv_halo_tc_kernel<MODE><<<grid, kHaloThreadsW, kHaloSmem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

template <int MODE>
int launch_halo64(const HaloArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(conv_halo64_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kH64Smem));
    configured = true;
  }
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  conv_halo64_tc_kernel<MODE><<<grid, kH64Threads, kH64Smem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

int g_halo64 = -1;
bool halo64_enabled() {                                      // TPZ_TRAIN_HALO64=0: 64-channel layers stay on the gather-GEMM
  if (g_halo64 < 0) {
    const char* e = getenv("TPZ_TRAIN_HALO64");
    g_halo64 = e ? atoi(e) : 1;
  }
  return g_halo64 != 0;
}

int launch_wgrad_halo(const HaloArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(wgrad_halo_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
    configured = true;
  }
  const int nby = a.ci_blocks * (a.dy_ld4 / 8);               // 32 x 32 blocks of dw
  int gx = sm_count() / nby;
  if (gx < 1) gx = 1;
  if (gx > a.ntiles) gx = a.ntiles;
  wgrad_halo_tc_kernel<<<dim3(gx, nby), kHaloThreadsW, kWgSmem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tpz_train_repack_tc(const float* flat_params, const void* descs, int ndesc, long long max_elems, float* packed,
                                   void* stream) {
  if (ndesc == 0) return 0;
  dim3 grid(tpz_div_up(max_elems, 256 * 4), ndesc);
  repack_tc_kernel<<<grid, 256, 0, ST(stream)>>>(flat_params, reinterpret_cast<const RepackTcDesc*>(descs), packed);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_fwd_tc(const float* x, int N, int H, int W, int Ci, const float* w_fwd_packed, const float* bias, int Co,
                               int kh, int kw, int stride, int dil, int org, const float* res, int res_H, int res_W, int res_org,
                               int res_stride, int relu, float* y, int Ho, int Wo, void* stream) {
  TPZ_CHECK(Ci % 32 == 0 && Co % 32 == 0, "tpz_conv_fwd_tc: needs Ci%%32==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  {
    HaloArgs h;
    memset(&h, 0, sizeof(h));
    if (halo_geometry(h, 0, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org)) {
      h.src = x; h.bias = bias; h.res = res; h.res_H = res_H; h.res_W = res_W; h.res_org = res_org; h.res_stride = res_stride;
      h.out = y; h.relu = relu;
      h.lo_rows = (long long)kh * kw * Co;
      int rc = weight_tmap(&h.tmB, w_fwd_packed, 2 * h.lo_rows, 32);
      if (rc) return rc;
      return launch_halo<0>(h, ST(stream));
    }
    memset(&h, 0, sizeof(h));
    if (halo64_enabled() && halo_geometry(h, 0, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org, 64)) {
      h.src = x; h.bias = bias; h.res = res; h.res_H = res_H; h.res_W = res_W; h.res_org = res_org; h.res_stride = res_stride;
      h.out = y; h.relu = relu;
      h.lo_rows = (long long)kh * kw * 2 * Co;
      int rc = weight_tmap(&h.tmB, w_fwd_packed, 2 * h.lo_rows, 64);
      if (rc) return rc;
      return launch_halo64<0>(h, ST(stream));
    }
  }
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.g = tgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  a.src = x; a.bias = bias; a.res = res; a.res_H = res_H; a.res_W = res_W; a.res_org = res_org; a.res_stride = res_stride;
  a.out = y; a.relu = relu; a.flush = flush_chunks();
  const long long rows = (long long)kh * kw * (Ci / 32) * Co;
  a.lo_rows = rows;
  const int BN = Co % 64 == 0 ? 64 : 32;
  int rc = weight_tmap(&a.tmB, w_fwd_packed, 2 * rows, BN);
  if (rc) return rc;
  const long long M = (long long)N * Ho * Wo;
  const int tiles = tpz_div_up(M, 128) * (Co / BN), nkb = kh * kw * (Ci / 32);
  if (res == nullptr && tiles * 8 <= sm_count() && nkb >= 16 && ksplit_enabled()) {
    // a handful of tiles with a long K loop (the last 5x5 layer on training crops: 4 tiles x 50 blocks): split K over the SMs
    int splits = 2 * sm_count() / tiles;
    if (splits > nkb / 4) splits = nkb / 4;
    a.ksplit = tpz_div_up(nkb, splits);
    splits = tpz_div_up(nkb, a.ksplit);
    TPZ_CUDA(cudaMemsetAsync(y, 0, (size_t)M * Co * sizeof(float), ST(stream)));
    rc = BN == 64 ? launch_conv_tc<64, 0>(a, M, Co, ST(stream), 0, splits) : launch_conv_tc<32, 0>(a, M, Co, ST(stream), 0, splits);
    if (rc) return rc;
    if (bias || relu) {
      const long long n4 = M * Co / 4;
      bias_act_kernel<<<tpz_div_up(n4, 256) < 1184 ? tpz_div_up(n4, 256) : 1184, 256, 0, ST(stream)>>>(y, bias, n4, Co, relu);
      TPZ_CUDA(cudaGetLastError());
    }
    return 0;
  }
  return BN == 64 ? launch_conv_tc<64, 0>(a, M, Co, ST(stream)) : launch_conv_tc<32, 0>(a, M, Co, ST(stream));
}

extern "C" int tpz_crop_add_f32(float* dx, int N, int H, int W, int C, const float* g, int Ho, int Wo, int org, int stride, void* stream);
extern "C" int tpz_relu_bwd_f32(float* dy, const float* y, long long n, void* stream);
extern "C" int tpz_bias_grad_f32(const float* dy, long long P, int C, float* db, void* stream);

extern "C" int tpz_conv_dgrad_tc(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw,
                                 int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx, int H, int W,
                                 void* stream);

// dx = mask(dgrad [+ dx] + res embedded at (res_org, res_org)): the data gradient of ResidA.conv0 together with the gradient of the
// cropped identity skip and the ReLU mask of the block input, in one pass when the halo-resident kernel takes the layer
extern "C" int tpz_conv_dgrad_tc_res(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw,
                                     int stride, int dil, int org, const float* relu_mask, int accumulate, const float* res, int res_H,
                                     int res_W, int res_org, float* dx, int H, int W, void* stream) {
  if (res == nullptr)
    return tpz_conv_dgrad_tc(dy, N, Ho, Wo, Co, w_dg_packed, Ci, kh, kw, stride, dil, org, relu_mask, accumulate, dx, H, W, stream);
  TPZ_CHECK(Ci % 32 == 0 && Co % 32 == 0, "tpz_conv_dgrad_tc_res: needs Ci%%32==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  HaloArgs h;
  memset(&h, 0, sizeof(h));
  if (halo_geometry(h, 1, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org)) {
    h.src = dy; h.mask = relu_mask; h.accumulate = accumulate; h.out = dx;
    h.res = res; h.res_H = res_H; h.res_W = res_W; h.res_org = res_org; h.res_stride = 1;
    h.lo_rows = (long long)kh * kw * Ci;
    int rc = weight_tmap(&h.tmB, w_dg_packed, 2 * h.lo_rows, 32);
    if (rc) return rc;
    return launch_halo<1>(h, ST(stream));
  }
  memset(&h, 0, sizeof(h));
  if (halo64_enabled() && halo_geometry(h, 1, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org, 64)) {
    h.src = dy; h.mask = relu_mask; h.accumulate = accumulate; h.out = dx;
    h.res = res; h.res_H = res_H; h.res_W = res_W; h.res_org = res_org; h.res_stride = 1;
    h.lo_rows = (long long)kh * kw * 2 * Ci;
    int rc = weight_tmap(&h.tmB, w_dg_packed, 2 * h.lo_rows, 64);
    if (rc) return rc;
    return launch_halo64<1>(h, ST(stream));
  }
  int rc = tpz_conv_dgrad_tc(dy, N, Ho, Wo, Co, w_dg_packed, Ci, kh, kw, stride, dil, org, nullptr, accumulate, dx, H, W, stream);
  if (rc) return rc;
  rc = tpz_crop_add_f32(dx, N, H, W, Ci, res, res_H, res_W, res_org, 1, stream);
  if (rc) return rc;
  return relu_mask ? tpz_relu_bwd_f32(dx, relu_mask, (long long)N * H * W * Ci, stream) : 0;
}

extern "C" int tpz_conv_dgrad_tc(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw,
                                 int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx, int H, int W,
                                 void* stream) {
  TPZ_CHECK(Ci % 32 == 0 && Co % 32 == 0, "tpz_conv_dgrad_tc: needs Ci%%32==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  {
    HaloArgs h;
    memset(&h, 0, sizeof(h));
    if (halo_geometry(h, 1, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org)) {
      h.src = dy; h.mask = relu_mask; h.accumulate = accumulate; h.out = dx;
      h.lo_rows = (long long)kh * kw * Ci;
      int rc = weight_tmap(&h.tmB, w_dg_packed, 2 * h.lo_rows, 32);
      if (rc) return rc;
      return launch_halo<1>(h, ST(stream));
    }
    memset(&h, 0, sizeof(h));
    if (halo64_enabled() && halo_geometry(h, 1, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org, 64)) {
      h.src = dy; h.mask = relu_mask; h.accumulate = accumulate; h.out = dx;
      h.lo_rows = (long long)kh * kw * 2 * Ci;
      int rc = weight_tmap(&h.tmB, w_dg_packed, 2 * h.lo_rows, 64);
      if (rc) return rc;
      return launch_halo64<1>(h, ST(stream));
    }
  }
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.g = tgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  a.src = dy; a.mask = relu_mask; a.accumulate = accumulate; a.out = dx; a.flush = flush_chunks();
  const long long rows = (long long)kh * kw * (Co / 32) * Ci;
  a.lo_rows = rows;
  const int BN = Ci % 64 == 0 ? 64 : 32;
  int rc = weight_tmap(&a.tmB, w_dg_packed, 2 * rows, BN);
  if (rc) return rc;
  a.lat = 1; a.lat0 = 0;
  long long M = (long long)N * H * W;
  if (stride > 1 && (dil % stride == 0 || (kh == 1 && kw == 1)) && org >= 0 && !(accumulate && relu_mask) && lattice_dgrad_enabled()) {
    // every tap reaches the same residue class of positions: run the GEMM on that sub-lattice only (1/stride^2 of the rows).
    // (accumulate AND mask together would have to mask the untouched positions too: left to the generic path)
    a.lat = stride; a.lat0 = org % stride;
    const int mh = (H - a.lat0 + stride - 1) / stride, mw = (W - a.lat0 + stride - 1) / stride;
    M = (long long)N * mh * mw;
    if (!accumulate) TPZ_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * H * W * Ci * sizeof(float), ST(stream)));
  }
  int grid_y = 0;
  const bool covered = org == 0 && dil == 1 && kh == H && kw == W;      // every dx pixel is reached by exactly one tap
  int grid_z = 1;
  if (Ho == 1 && Wo == 1 && stride == 1 && Ci % BN == 0 && org >= 0 && org + (kh - 1) * dil < H && org + (kw - 1) * dil < W &&
      (covered || !(accumulate && relu_mask)) && scatter_dgrad_enabled()) {
    // one source pixel per image: dx pixel (org + r*dil, org + t*dil) = dy x w[r][t]; a GEMM with M = images per tap
    a.scatter = 1; a.lat = 1; a.lat0 = 0;
    M = N; grid_y = kh * kw; grid_z = Ci / BN;
    if (!accumulate && !covered) TPZ_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * H * W * Ci * sizeof(float), ST(stream)));
  }
  return BN == 64 ? launch_conv_tc<64, 1>(a, M, Ci, ST(stream), grid_y, grid_z) : launch_conv_tc<32, 1>(a, M, Ci, ST(stream), grid_y, grid_z);
}

extern "C" int tpz_conv_wgrad_tc(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh,
                                 int kw, int stride, int dil, int org, float* dw, void* stream);

// weight gradient and bias gradient (db[c] += sum over pixels of dy[.][c]; db may be NULL) -- one kernel on the halo-resident path
extern "C" int tpz_conv_wgrad_tc_bias(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh,
                                      int kw, int stride, int dil, int org, float* dw, float* db, void* stream) {
  TPZ_CHECK(Ci % 32 == 0 && Co % 32 == 0, "tpz_conv_wgrad_tc_bias: needs Ci%%32==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  HaloArgs h;
  memset(&h, 0, sizeof(h));
  if (halo_wgrad_enabled() && halo_geometry(h, 2, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org)) {
    h.src = x; h.dy = dy; h.dw = dw; h.db = db;
    return launch_wgrad_halo(h, ST(stream));
  }
  int rc = tpz_conv_wgrad_tc(x, N, H, W, Ci, dy, Ho, Wo, Co, kh, kw, stride, dil, org, dw, stream);
  if (rc || !db) return rc;
  return tpz_bias_grad_f32(dy, (long long)N * Ho * Wo, Co, db, stream);
}

extern "C" int tpz_conv_wgrad_tc(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh,
                                 int kw, int stride, int dil, int org, float* dw, void* stream) {
  TPZ_CHECK(Ci % 32 == 0 && Co % 32 == 0, "tpz_conv_wgrad_tc: needs Ci%%32==0 and Co%%32==0 (Ci=%d Co=%d)", Ci, Co);
  {
    HaloArgs h;
    memset(&h, 0, sizeof(h));
    if (halo_wgrad_enabled() && halo_geometry(h, 2, N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org)) {
      h.src = x; h.dy = dy; h.dw = dw;
      return launch_wgrad_halo(h, ST(stream));
    }
  }
  WgArgs a;
  a.g = tgeom(N, H, W, Ci, Ho, Wo, Co, kh, kw, stride, dil, org);
  a.x = x; a.dy = dy; a.dw = dw; a.flush = flush_chunks();
  const long long P = (long long)N * Ho * Wo;
  const int BN = Co % 64 == 0 ? 64 : 32;
  const int mt = tpz_div_up((long long)kh * kw * Ci, 128), nt = Co / BN;
  // split the pixels so that ONE wave of 2 CTAs per SM covers the layer (a second wave costs a whole CTA lifetime -- TMEM allocation,
  // first-load latency, 8192 atomics -- on these latency-bound launches), at least 8 chunks of 32 pixels per split
  int splits = (sm_count() * 2) / (mt * nt);
  if (splits < 1) splits = 1;
  long long kps = (P + splits - 1) / splits;
  kps = (kps + 31) / 32 * 32;
  if (kps < 256) kps = 256;
  splits = (int)((P + kps - 1) / kps);
  a.k_per_split = (int)kps;
  dim3 grid(mt, nt, splits);
  if (BN == 64) {
    constexpr int smem = 2 * (2 * kStageA + 2 * 64 * 128) + 1024;
    static bool configured = false;
    if (!configured) { TPZ_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); configured = true; }
    wgrad_tc_kernel<64><<<grid, 256, smem, ST(stream)>>>(a);
  } else {
    constexpr int smem = 2 * (2 * kStageA + 2 * 32 * 128) + 1024;
    static bool configured = false;
    if (!configured) { TPZ_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); configured = true; }
    wgrad_tc_kernel<32><<<grid, 256, smem, ST(stream)>>>(a);
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// Cin = 1, Co = 32, 7 x 7 first layer on the tensor core (see first_tc_train_kernel); returns -1 when the geometry is not covered
// (the caller then uses the CUDA-core kernels of tpz_train.cu).
static int g_first_tc = -1;
static bool first_tc_enabled() {                              // TPZ_TRAIN_FIRST_TC=0: CUDA-core first layer
  if (g_first_tc < 0) {
    const char* e = getenv("TPZ_TRAIN_FIRST_TC");
    g_first_tc = e ? atoi(e) : 1;
  }
  return g_first_tc != 0;
}
extern "C" int tpz_first_fwd_tc(const float* x, int N, int H, int W, const float* w, const float* bias, int Co, int k, int stride,
                                int relu, float* y, int Ho, int Wo, void* stream) {
  if (!first_tc_enabled() || Co != 32 || k != kFirstK || (Ho - 1) * stride + k > H || (Wo - 1) * stride + k > W) return -1;
  FirstArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.N = N; a.H = H; a.W = W; a.stride = stride; a.Ho = Ho; a.Wo = Wo; a.w = w; a.bias = bias; a.relu = relu; a.y = y;
  a.M = (long long)N * Ho * Wo; a.ntiles = tpz_div_up(a.M, 128);
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(first_tc_train_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFirstFwdSmem));
    configured = true;
  }
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  first_tc_train_kernel<0><<<grid, kHaloThreads, kFirstFwdSmem, ST(stream)>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_first_wgrad_tc(const float* x, int N, int H, int W, const float* dy, int Ho, int Wo, int Co, int k, int stride,
                                  float* dw, float* db, void* stream) {
  if (!first_tc_enabled() || Co != 32 || k != kFirstK || (Ho - 1) * stride + k > H || (Wo - 1) * stride + k > W) return -1;
  FirstArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.N = N; a.H = H; a.W = W; a.stride = stride; a.Ho = Ho; a.Wo = Wo; a.dy = dy; a.dw = dw; a.db = db;
  a.M = (long long)N * Ho * Wo; a.ntiles = tpz_div_up(a.M, 128);
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(first_tc_train_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFirstWgSmem));
    configured = true;
  }
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  first_tc_train_kernel<1><<<grid, kHaloThreads, kFirstWgSmem, ST(stream)>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

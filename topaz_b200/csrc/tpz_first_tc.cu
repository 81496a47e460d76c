// Cin = 1 first convolution as ONE tensor-core kernel: the im2col tile is built in shared memory (never in HBM).
//
// Replaces the first BasicConv 7x7 of the classifiers (resnet.py:66,102; basic.py:47-52) and the 11x11 enc1 conv of
// UDenoiseNet (denoising/models.py:79) in dense (stride-1, dilation-1) mode.  The earlier path wrote the im2col matrix
// to HBM (2*ld B/px) and read it back in a 1-tap GEMM: 3 x 2.1 GB of traffic for a 4096^2 micrograph, 2.9 ms.  Here the
// only HBM traffic is the fp32 image in (4 B/px) and the fp16 NHWC feature map out (2*Cp B/px), which bounds the kernel.
//
// CTA = 128 threads = 128 output pixels (8 rows x 16 columns) = the 128 rows of the A operand.  Per tile:
//   1. the (8+k-1) x (16+k-1) input window is staged in smem (zero outside the image = the conv padding),
//   2. thread p writes row p of A: its k*k taps as fp16, K padded to 64/128, in the canonical K-major SWIZZLE_128B
//      layout (16-byte chunk c of row p lands at chunk c ^ (p & 7)), generic-proxy stores + fence.proxy.async,
//   3. one elected thread issues Kp/16 tcgen05.mma (M=128, N=Cp) against the smem-resident weights, commit -> mbarrier,
//   4. epilogue: tcgen05.ld -> +bias -> activation -> fp16 -> 16-byte global stores (a pixel's Cp channels are contiguous).
// A CTA is strictly sequential per tile; up to 8 CTAs are resident per SM (64 TMEM columns, ~27 KB smem, <= 64
// registers each) and overlap one another, and each CTA prefetches its next input window into registers.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {

constexpr int TH = 8, TW = 16;            // output tile: 8 rows x 16 columns = 128 pixels = 128 A rows / TMEM lanes

struct FirstArgs {
  const float* x; int B, H, W;
  const __half* w;        // [Kp/64][Cp][64] fp16 (k-block major), tap t = r*k+s, zero beyond k*k and beyond real Cout
  const float* bias;      // [Cp]
  __half* out;            // [B][Ho][Wo][Cp]
  int Ho, Wo, pad;
  float slope;
  int tiles_x, tiles_y;
  int pool;               // 1: fused 2x2 max-pool (floor), out is [B][Ho/2][Wo/2][Cp]
  const float* range;     // optional device float[2] = (s, 1/s): input and bias are multiplied by s (tpz_range_scale)
};

template <int KW>
struct FirstGeom {
  static constexpr int TAPS = KW * KW;
  static constexpr int KB = (TAPS + 63) / 64;                 // 64-wide k-blocks (7x7 -> 1, 11x11 -> 2)
  static constexpr int PH = TH + KW - 1, PW = TW + KW - 1;    // input window
  // window pitch == 16 (mod 32): a warp reads two tile rows of 16 pixels, the second row must land on the other 16 banks
  static constexpr int PWP = PW <= 16 ? 16 : (PW <= 48 ? 48 : 80);
  static constexpr int NPRE = (PH * PW + 127) / 128;          // window elements prefetched per thread
};

template <int KW, int CP, int MINB>
__global__ void __launch_bounds__(128, MINB) first_tc_kernel(const FirstArgs a) {
  using G = FirstGeom<KW>;
  constexpr int TAPS = G::TAPS, KB = G::KB, PH = G::PH, PW = G::PW, PWP = G::PWP, NPRE = G::NPRE;
  constexpr uint32_t A_BYTES = 128 * 128;            // one k-block of A: 128 rows x 128 B
  constexpr uint32_t B_BYTES = CP * 128;
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, CP);
  constexpr int TCOLS = CP < 32 ? 32 : CP;           // one accumulator; CTAs co-resident on the SM overlap each other

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;                                   // [KB][128 x 128 B]
  unsigned char* sB = base + KB * A_BYTES;                    // [KB][CP x 128 B]
  float* sImg = reinterpret_cast<float*>(sB + KB * B_BYTES);  // [PH][PWP]
  float* sBias = sImg + PH * PWP;                             // [CP]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<TCOLS>(&tmem_base_s);
  // weights -> smem in the same swizzled K-major layout (row n, 16-byte chunk c at chunk c ^ (n & 7))
  for (int i = tid; i < KB * CP * 8; i += 128) {
    const int c = i & 7, n = (i >> 3) % CP, kb = i / (8 * CP);
    const uint4 v = *reinterpret_cast<const uint4*>(a.w + ((size_t)(kb * CP + n) * 64 + c * 8));
    *reinterpret_cast<uint4*>(sB + kb * B_BYTES + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  const float rs = a.range ? a.range[0] : 1.f;         // range guard: a power of two, so x*rs and the later 1/rs are exact
  for (int i = tid; i < CP; i += 128) sBias[i] = a.bias[i] * rs;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int py = tid / TW, px = tid % TW;            // this thread's pixel inside the tile
  const long long tiles_per_img = (long long)a.tiles_x * a.tiles_y;
  const long long ntiles = tiles_per_img * a.B;
  const uint32_t a_hi = ptx::umma_desc_hi(1024, 2), b_hi = a_hi;       // SBO = 8 rows x 128 B, SWIZZLE_128B
  const uint32_t a_lo0 = (ptx::smem_u32(sA) & 0x3FFFF) >> 4, b_lo0 = (ptx::smem_u32(sB) & 0x3FFFF) >> 4;

  // input window of a tile -> registers (zero outside the image = the conv padding); issued one tile ahead so the
  // global-load latency hides behind the previous tile's MMA + epilogue
  float pre[NPRE];
  auto load_window = [&](long long tile) {
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int y0 = (tr / a.tiles_x) * TH - a.pad, x0 = (tr % a.tiles_x) * TW - a.pad;
    const float* img = a.x + (size_t)b * a.H * a.W;
#pragma unroll
    for (int e = 0; e < NPRE; ++e) {
      const int i = tid + e * 128;
      const int wy = i / PW, wx = i - wy * PW;
      const int iy = y0 + wy, ix = x0 + wx;
      pre[e] = (i < PH * PW && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) ? __ldg(img + (size_t)iy * a.W + ix) * rs : 0.f;
    }
  };

  long long tile = blockIdx.x;
  if (tile < ntiles) load_window(tile);
  uint32_t phase = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    // 1. window registers -> smem (sImg was last read before the second __syncthreads of the previous iteration)
#pragma unroll
    for (int e = 0; e < NPRE; ++e) {
      const int i = tid + e * 128;
      if (i < PH * PW) { const int wy = i / PW; sImg[wy * PWP + (i - wy * PW)] = pre[e]; }
    }
    __syncthreads();
    // 2. this pixel's im2col row (the A tile was last read by the previous tile's MMAs, whose completion every thread
    //    waited for before its epilogue)
    {
      unsigned char* rowp = sA + tid * 128;
      const float* win = sImg + py * PWP + px;
      const int sw = tid & 7;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (kb * 64 + c * 8 >= TAPS) {              // all-zero chunk (K padding)
            *reinterpret_cast<uint4*>(rowp + kb * A_BYTES + ((c ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
            continue;
          }
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int t0 = kb * 64 + c * 8 + 2 * j, t1 = t0 + 1;
            const float v0 = t0 < TAPS ? win[(t0 / KW) * PWP + (t0 % KW)] : 0.f;
            const float v1 = t1 < TAPS ? win[(t1 / KW) * PWP + (t1 % KW)] : 0.f;
            const __half2 h = __floats2half2_rn(v0, v1);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(rowp + kb * A_BYTES + ((c ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    ptx::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    // the next window is requested only now: fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS, and the membar would wait
    // for loads issued before it (they now fly during this tile's MMAs and epilogue instead)
    if (tile + gridDim.x < ntiles) load_window(tile + gridDim.x);
    // 3. MMAs of this tile
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kb * 64 + k * 16 >= TAPS) continue;   // K=16 slice that is entirely padding
            ptx::umma_f16_lohi(tmem_base, a_lo0 + kb * (A_BYTES >> 4) + 2 * k, a_hi, b_lo0 + kb * (B_BYTES >> 4) + 2 * k,
                               b_hi, IDESC, (kb | k) ? 1u : 0u);
          }
        }
        ptx::umma_commit(&bar);
      }
      __syncwarp();
    }
    // 4. epilogue (other CTAs resident on this SM fill the tensor-core / memory pipes meanwhile).  The finished MMAs no
    //    longer need the A tile, so its first CP*256 bytes stage the fp16 output tile: each thread deposits its pixel's
    //    channel vector (16-byte chunks, XOR-swizzled against bank conflicts), then the warp writes its 32 pixels back
    //    with every store instruction covering whole 128-byte lines (CP*2/32 lanes per pixel, 32 bytes per lane).
    ptx::mbar_wait(&bar, phase);
    phase ^= 1u;
    ptx::tc_fence_after();
    {
      constexpr int ROWB = CP * 2;                    // bytes per pixel
      constexpr int CH16 = ROWB / 16;                 // 16-byte chunks per pixel (4 or 8)
      unsigned char* stg = sA + (size_t)tid * ROWB;
      const int sw = tid & (CH16 - 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < CP; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = q * 8 + 2 * j;
            float v0 = __uint_as_float(r[ch]) + sBias[c0 + ch];
            float v1 = __uint_as_float(r[ch + 1]) + sBias[c0 + ch + 1];
            v0 = v0 > 0.f ? v0 : v0 * a.slope;
            v1 = v1 > 0.f ? v1 : v1 * a.slope;
            const __half2 h = __floats2half2_rn(v0, v1);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(stg + (((c0 / 8 + q) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      __syncwarp();
      constexpr int LPP = ROWB / 32;                  // lanes per pixel (2 or 4), 32 bytes each
      constexpr int PPI = 32 / LPP;                   // pixels per store instruction
      const int lane = tid & 31;
      const int b = (int)(tile / tiles_per_img);
      const int tr = (int)(tile % tiles_per_img);
      const int ty0 = (tr / a.tiles_x) * TH, tx0 = (tr % a.tiles_x) * TW;
      if (!a.pool) {
#pragma unroll
        for (int i = 0; i < 32 / PPI; ++i) {
          const int pix = warp * 32 + i * PPI + lane / LPP;        // pixel (row of the tile) this lane helps to write
          const int part = lane % LPP;                              // which 32-byte piece of its channel vector
          const int oy = ty0 + pix / TW, ox = tx0 + pix % TW;
          const unsigned char* src = sA + (size_t)pix * ROWB;
          const int psw = pix & (CH16 - 1);
          const uint4 lo = *reinterpret_cast<const uint4*>(src + (((2 * part) ^ psw) << 4));
          const uint4 hi = *reinterpret_cast<const uint4*>(src + (((2 * part + 1) ^ psw) << 4));
          if (oy < a.Ho && ox < a.Wo)
            ptx::st_global_256(a.out + (((size_t)b * a.Ho + oy) * a.Wo + ox) * CP + part * 16, lo.x, lo.y, lo.z, lo.w, hi.x,
                               hi.y, hi.z, hi.w);
        }
      } else {
        // fused MaxPool2d(2) (denoising/models.py:80: enc1 = conv, LeakyReLU, MaxPool): a warp owns tile rows 2w, 2w+1, i.e.
        // one row of 8 pooled pixels; lane -> (pooled pixel, 32-byte piece), max over the 2x2 staged vectors
        const int Hp = a.Ho >> 1, Wp = a.Wo >> 1;
#pragma unroll
        for (int i = 0; i < (8 * LPP + 31) / 32; ++i) {
          const int item = i * 32 + lane;
          if (item < 8 * LPP) {
            const int pxp = item / LPP, part = item % LPP;         // pooled column inside the tile, 32-byte piece
            const int oyp = (ty0 >> 1) + warp, oxp = (tx0 >> 1) + pxp;
            uint4 mlo = make_uint4(0, 0, 0, 0), mhi = mlo;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int pix = (2 * warp + (q >> 1)) * TW + 2 * pxp + (q & 1);
              const unsigned char* src = sA + (size_t)pix * ROWB;
              const int psw = pix & (CH16 - 1);
              const uint4 lo = *reinterpret_cast<const uint4*>(src + (((2 * part) ^ psw) << 4));
              const uint4 hi = *reinterpret_cast<const uint4*>(src + (((2 * part + 1) ^ psw) << 4));
              if (q == 0) { mlo = lo; mhi = hi; }
              else {
                __half2* a2 = reinterpret_cast<__half2*>(&mlo); const __half2* b2 = reinterpret_cast<const __half2*>(&lo);
                __half2* c2 = reinterpret_cast<__half2*>(&mhi); const __half2* d2 = reinterpret_cast<const __half2*>(&hi);
#pragma unroll
                for (int e = 0; e < 4; ++e) { a2[e] = __hmax2(a2[e], b2[e]); c2[e] = __hmax2(c2[e], d2[e]); }
              }
            }
            if (oyp < Hp && oxp < Wp)
              ptx::st_global_256(a.out + (((size_t)b * Hp + oyp) * Wp + oxp) * CP + part * 16, mlo.x, mlo.y, mlo.z, mlo.w,
                                 mhi.x, mhi.y, mhi.z, mhi.w);
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<TCOLS>(tmem_base);
}

template <int KW, int CP>
int launch_first(const FirstArgs& a, cudaStream_t stream) {
  using G = FirstGeom<KW>;
  constexpr int MINB = G::KB == 1 ? 8 : 4;
  size_t smem = 1024 + G::KB * 128 * 128 + G::KB * CP * 128 + (G::PH * G::PWP + CP) * sizeof(float);
  // resident CTAs per SM: TMEM allows 512 / max(32, CP); registers / threads allow MINB.  Ask for enough shared memory that
  // the hardware cannot co-schedule more than that (an extra CTA would spin in tcgen05.alloc until another one exits).
  int per_sm = 512 / (CP < 32 ? 32 : CP);
  if (per_sm > MINB) per_sm = MINB;
  while (per_sm > 1 && (size_t)per_sm * (smem + 1024) > 227 * 1024) --per_sm;
  const size_t floor_smem = (227 * 1024) / (per_sm + 1) + 1;
  if (smem < floor_smem) smem = floor_smem;
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(first_tc_kernel<KW, CP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (long long)a.tiles_x * a.tiles_y * a.B;
  const int grid = (int)(ntiles < (long long)sms * per_sm ? ntiles : (long long)sms * per_sm);
  first_tc_kernel<KW, CP, MINB><<<grid, 128, smem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int tpz_conv_first_tc(const float* x, int B, int H, int W, const tpz_half* w_packed, const float* bias, int Cp,
                                 int k, int pad, float neg_slope, int pool, tpz_half* out, const float* range, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FirstArgs a;
  a.x = x; a.B = B; a.H = H; a.W = W;
  a.w = reinterpret_cast<const __half*>(w_packed); a.bias = bias; a.out = reinterpret_cast<__half*>(out);
  a.Ho = H + 2 * pad - (k - 1); a.Wo = W + 2 * pad - (k - 1); a.pad = pad; a.slope = neg_slope; a.pool = pool ? 1 : 0; a.range = range;
  TPZ_CHECK(a.Ho > 0 && a.Wo > 0 && B > 0, "tpz_conv_first_tc: empty output (H=%d W=%d k=%d pad=%d)", H, W, k, pad);
  a.tiles_x = tpz_div_up(a.Wo, TW); a.tiles_y = tpz_div_up(a.Ho, TH);
#define TPZ_FIRST_CASE(KW_, CP_) if (k == KW_ && Cp == CP_) return launch_first<KW_, CP_>(a, stream);
  TPZ_FIRST_CASE(7, 32) TPZ_FIRST_CASE(7, 64) TPZ_FIRST_CASE(11, 64) TPZ_FIRST_CASE(11, 32)
  TPZ_FIRST_CASE(5, 32) TPZ_FIRST_CASE(5, 64) TPZ_FIRST_CASE(3, 32) TPZ_FIRST_CASE(3, 64)
#undef TPZ_FIRST_CASE
  TPZ_CHECK(false, "tpz_conv_first_tc: unsupported (k=%d, Cp=%d); supported k in {3,5,7,11}, Cp in {32,64}", k, Cp);
  return 1;
}

extern "C" int tpz_conv_first_tc_supported(int k, int Cp) {
  return (k == 3 || k == 5 || k == 7 || k == 11) && (Cp == 32 || Cp == 64);
}

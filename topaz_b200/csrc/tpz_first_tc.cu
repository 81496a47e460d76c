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
//   4. the epilogue of the PREVIOUS tile (other TMEM stage, other A buffer) runs while those MMAs execute:
//      tcgen05.ld -> +bias -> activation -> fp16 -> 16-byte global stores (each pixel's Cp channels are contiguous).
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
};

template <int KW, int CP>
__global__ void __launch_bounds__(128) first_tc_kernel(const FirstArgs a) {
  constexpr int TAPS = KW * KW;
  constexpr int KB = (TAPS + 63) / 64;               // 64-wide k-blocks (7x7 -> 1, 11x11 -> 2)
  constexpr int PH = TH + KW - 1, PW = TW + KW - 1;  // input window
  constexpr int PWP = PW | 1;                        // odd pitch: conflict-free column walks
  constexpr uint32_t A_BYTES = 128 * 128;            // one k-block of A: 128 rows x 128 B
  constexpr uint32_t B_BYTES = CP * 128;
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, CP);
  constexpr int TCOLS = 2 * CP < 32 ? 32 : 2 * CP;   // two accumulator stages

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;                                   // [2 stages][KB][128 x 128 B]
  unsigned char* sB = base + 2 * KB * A_BYTES;                // [KB][CP x 128 B]
  float* sImg = reinterpret_cast<float*>(sB + KB * B_BYTES);  // [PH][PWP]
  float* sBias = sImg + PH * PWP;                             // [CP]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<TCOLS>(&tmem_base_s);
  // weights -> smem in the same swizzled K-major layout (row n, 16-byte chunk c at chunk c ^ (n & 7))
  for (int i = tid; i < KB * CP * 8; i += 128) {
    const int c = i & 7, n = (i >> 3) % CP, kb = i / (8 * CP);
    const uint4 v = *reinterpret_cast<const uint4*>(a.w + ((size_t)(kb * CP + n) * 64 + c * 8));
    *reinterpret_cast<uint4*>(sB + kb * B_BYTES + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  for (int i = tid; i < CP; i += 128) sBias[i] = a.bias[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int py = tid / TW, px = tid % TW;            // this thread's pixel inside the tile
  const long long tiles_per_img = (long long)a.tiles_x * a.tiles_y;
  const long long ntiles = tiles_per_img * a.B;
  const uint32_t a_hi = ptx::umma_desc_hi(1024, 2), b_hi = a_hi;       // SBO = 8 rows x 128 B, SWIZZLE_128B
  const uint32_t sA_u32 = ptx::smem_u32(sA), sB_u32 = ptx::smem_u32(sB);

  long long prev_tile = -1;
  int it = 0;
  uint32_t phases = 0;                               // bit s = parity of the next completion of bar[s]
  auto epilogue = [&](long long tile, int stage) {
    ptx::mbar_wait(&bar[stage], (phases >> stage) & 1u);
    phases ^= 1u << stage;
    ptx::tc_fence_after();
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int oy = (tr / a.tiles_x) * TH + py, ox = (tr % a.tiles_x) * TW + px;
    const bool ok = oy < a.Ho && ox < a.Wo;
    __half* dst = a.out + (((size_t)b * a.Ho + oy) * a.Wo + ox) * CP;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + stage * CP;
#pragma unroll
    for (int c0 = 0; c0 < CP; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(taddr + c0, r);
      ptx::tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v0 = __uint_as_float(r[2 * j]) + sBias[c0 + 2 * j];
        float v1 = __uint_as_float(r[2 * j + 1]) + sBias[c0 + 2 * j + 1];
        v0 = v0 > 0.f ? v0 : v0 * a.slope;
        v1 = v1 > 0.f ? v1 : v1 * a.slope;
        const __half2 h = __floats2half2_rn(v0, v1);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      if (ok) {
        *reinterpret_cast<uint4*>(dst + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    ptx::tc_fence_before();
  };

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const int b = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int y0 = (tr / a.tiles_x) * TH - a.pad, x0 = (tr % a.tiles_x) * TW - a.pad;   // window origin in the image
    // 1. input window (sImg was last read before the previous __syncthreads)
    const float* img = a.x + (size_t)b * a.H * a.W;
    for (int i = tid; i < PH * PW; i += 128) {
      const int wy = i / PW, wx = i - wy * PW;
      const int iy = y0 + wy, ix = x0 + wx;
      sImg[wy * PWP + wx] = (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) ? img[(size_t)iy * a.W + ix] : 0.f;
    }
    __syncthreads();
    // 2. this pixel's im2col row (A buffer `stage` was last read by the MMAs of tile it-2, whose completion the
    //    epilogue of that tile waited for)
    {
      unsigned char* rowp = sA + stage * (KB * A_BYTES) + tid * 128;
      const float* win = sImg + py * PWP + px;
      const int sw = tid & 7;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
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
    // 3. MMAs of this tile
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d = tmem_base + stage * CP;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a_lo = ((sA_u32 + stage * (KB * A_BYTES) + kb * A_BYTES) & 0x3FFFF) >> 4;
          const uint32_t b_lo = ((sB_u32 + kb * B_BYTES) & 0x3FFFF) >> 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16_lohi(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, IDESC, (kb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&bar[stage]);
      }
      __syncwarp();
    }
    // 4. epilogue of the previous tile while the tensor core works
    if (prev_tile >= 0) epilogue(prev_tile, stage ^ 1);
    prev_tile = tile;
  }
  if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1);
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<TCOLS>(tmem_base);
}

template <int KW, int CP>
int launch_first(const FirstArgs& a, cudaStream_t stream) {
  constexpr int TAPS = KW * KW, KB = (TAPS + 63) / 64, PH = TH + KW - 1, PW = (TW + KW - 1) | 1;
  size_t smem = 1024 + 2 * KB * 128 * 128 + KB * CP * 128 + (PH * PW + CP) * sizeof(float);
  // resident CTAs per SM are bounded by TMEM (512 columns / 2*CP per CTA, at most 4 wanted); ask for enough shared
  // memory that the hardware cannot co-schedule more than that (an extra CTA would spin in tcgen05.alloc)
  int per_sm = 512 / (2 * CP < 32 ? 32 : 2 * CP);
  if (per_sm > 4) per_sm = 4;
  while (per_sm > 1 && (size_t)per_sm * (smem + 1024) > 227 * 1024) --per_sm;
  const size_t floor_smem = (227 * 1024) / (per_sm + 1) + 1;
  if (smem < floor_smem) smem = floor_smem;
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(first_tc_kernel<KW, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (long long)a.tiles_x * a.tiles_y * a.B;
  const int grid = (int)(ntiles < (long long)sms * per_sm ? ntiles : (long long)sms * per_sm);
  first_tc_kernel<KW, CP><<<grid, 128, smem, stream>>>(a);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int tpz_conv_first_tc(const float* x, int B, int H, int W, const tpz_half* w_packed, const float* bias, int Cp,
                                 int k, int pad, float neg_slope, tpz_half* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FirstArgs a;
  a.x = x; a.B = B; a.H = H; a.W = W;
  a.w = reinterpret_cast<const __half*>(w_packed); a.bias = bias; a.out = reinterpret_cast<__half*>(out);
  a.Ho = H + 2 * pad - (k - 1); a.Wo = W + 2 * pad - (k - 1); a.pad = pad; a.slope = neg_slope;
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

// tcgen05 implicit-GEMM convolution for sm_100a (B200).
//
// Replaces the cuDNN conv calls of the reference's dense ("filled") classifier forward and U-Net
// denoiser forward (reference: topaz/model/features/resnet.py:101-105,178-204, basic.py:101-111,
// denoising/models.py:130-175,508-564).  Activations are channels-last fp16 ([N][D][H][W][C]); weights are
// fp16, repacked per k-block as [kb][Cout][KC]; accumulation is fp32 in TMEM; bias/activation/residual/
// classifier-dot run in the epilogue in fp32.
//
// GEMM view: M = 128 output pixels (a TW x TH tile of one (n,z) plane), N = Cout, K = sum over
// (source, tap, channel chunk of KC).  Because every conv on this path is stride 1, the A operand of a
// tap is the input tile shifted by tap*dilation: one tiled TMA load per k-block, out-of-bounds = zero
// (this is also how "same" zero padding and the classifier's single input pad are realised).
//
// Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM
// allocator, warps4-7 = epilogue (TMEM lane quadrant = warp%4).  smem ring of S stages (full/empty
// mbarriers), two TMEM accumulator stages (tmem_full/tmem_empty) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Persistent grid, static round-robin tile schedule.
#include "tpz_common.cuh"
#include "tpz_tc_conv.h"

namespace {

constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) {
  }
}

template <int KC>
__global__ void __launch_bounds__(kThreads, 1) tc_conv_kernel(const __grid_constant__ TcConvParams p) {
  constexpr int ROWB = KC * 2;                       // bytes per operand row == swizzle span
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t SBO = 8 * ROWB;                 // 8-row core-matrix group stride
  constexpr int A_BYTES = 128 * ROWB;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int S = p.stages;
  const int b_bytes = p.Co * ROWB;
  const int stage_bytes = A_BYTES + b_bytes;
  uint8_t* tail = smem + (size_t)S * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_dotw = s_bias + 256;
  float* s_osc = s_dotw + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.Co; i += kThreads) {
    // range guard: activations are stored multiplied by range[0] (a power of two), so the bias is scaled with them
    s_bias[i] = p.bias ? p.bias[i] * (p.range ? p.range[0] : 1.f) : 0.f;
    s_dotw[i] = p.dot_w ? p.dot_w[i] : 0.f;
    s_osc[i] = p.oscale ? p.oscale[i] : 1.f;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA[0]);
    if (p.nsrc > 1) ptx::prefetch_tmap(&p.tmA[1]);
    ptx::prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_plane = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int plane = tile / tiles_per_plane;
        const int rem = tile - plane * tiles_per_plane;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int n = plane / p.Do, z = plane - n * p.Do;
        const int x0 = tx * p.TW, y0 = ty * p.TH;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + A_BYTES;
          ptx::mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
          const TcKBlock kbv = p.kb[kb];
          const int src = kbv.src;
          ptx::tma_load_5d(a_dst, &p.tmA[src], &full[s], kbv.c0, x0 + p.org[src][0] + kbv.dx,
                           y0 + p.org[src][1] + kbv.dy, z + p.org[src][2] + kbv.dz, n);
          ptx::tma_load_2d(b_dst, &p.tmB, &full[s], 0, kb * p.Co);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(128, p.Co);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = base + (uint32_t)s * stage_bytes;
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            const uint64_t da = ptx::umma_desc(a_addr + k * 32, SBO, LAYOUT);
            const uint64_t db = ptx::umma_desc(b_addr + k * 32, SBO, LAYOUT);
            ptx::umma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
        }
        ptx::umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int ew = warp - 4;
    const int m = ew * 32 + lane;
    const int ly = m / p.TW, lx = m - ly * p.TW;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int plane = tile / tiles_per_plane;
      const int rem = tile - plane * tiles_per_plane;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int n = plane / p.Do, z = plane - n * p.Do;
      const int gx = tx * p.TW + lx, gy = ty * p.TH + ly;
      const bool valid = (gx < p.Wo) && (gy < p.Ho);
      const uint32_t acc = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * kAccStride;

      const long long opix = (((long long)n * p.Do + z) * p.Ho + gy) * p.Wo + gx;
      __half* orow = p.out ? p.out + opix * p.out_ld + p.out_coff : nullptr;
      const __half* rrow = nullptr;
      if (p.res) {
        const long long rpix =
            (((long long)n * p.res_D + (z + p.res_org[2])) * p.res_H + (gy + p.res_org[1])) * p.res_W +
            (gx + p.res_org[0]);
        rrow = p.res + rpix * p.res_ld;
      }
      float dot = 0.f;
      for (int c = 0; c < p.Co; c += 32) {
        uint32_t r[32];
        if (p.Co - c >= 32) {
          ptx::tmem_ld32(taddr + c, r);
        } else {  // Co % 32 == 16 tail
          uint32_t r16[16];
          ptx::tmem_ld16(taddr + c, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { r[j] = r16[j]; r[16 + j] = 0; }
        }
        ptx::tmem_ld_wait();
        const int nc = min(32, p.Co - c);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), s_osc[min(c + j, 255)], s_bias[min(c + j, 255)]);
        if (rrow && valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 8 < nc) {
              const uint4 u = *reinterpret_cast<const uint4*>(rrow + c + q * 8);
              const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                const int j = q * 8 + e * 2;
                const float s0 = p.res_scale ? p.res_scale[c + j] : 1.f;
                const float s1 = p.res_scale ? p.res_scale[c + j + 1] : 1.f;
                v[j] += s0 * f.x;
                v[j + 1] += s1 * f.y;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.neg_slope;
        if (p.dot_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) dot = fmaf(v[j], s_dotw[c + j], dot);
        }
        if (orow && valid) {
          const bool wide = ((reinterpret_cast<uintptr_t>(orow + c) & 31) == 0);     // full-sector 32-byte stores
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q * 16 < nc) {
              uint32_t u[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const __half2 h = __floats2half2_rn(v[q * 16 + e * 2], v[q * 16 + e * 2 + 1]);
                u[e] = *reinterpret_cast<const uint32_t*>(&h);
              }
              if (wide && q * 16 + 8 < nc) {
                ptx::st_global_256(orow + c + q * 16, u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7]);
              } else {
                *reinterpret_cast<uint4*>(orow + c + q * 16) = make_uint4(u[0], u[1], u[2], u[3]);
                if (q * 16 + 8 < nc) *reinterpret_cast<uint4*>(orow + c + q * 16 + 8) = make_uint4(u[4], u[5], u[6], u[7]);
              }
              if (p.out_lo > 0) {          // strict mode: residual of the fp16 rounding, lo = fp16(v - hi)
                uint32_t ul[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&u[e]));
                  const __half2 l = __floats2half2_rn(v[q * 16 + e * 2] - hf.x, v[q * 16 + e * 2 + 1] - hf.y);
                  ul[e] = *reinterpret_cast<const uint32_t*>(&l);
                }
                *reinterpret_cast<uint4*>(orow + p.out_lo + c + q * 16) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
                if (q * 16 + 8 < nc) *reinterpret_cast<uint4*>(orow + p.out_lo + c + q * 16 + 8) = make_uint4(ul[4], ul[5], ul[6], ul[7]);
              }
            }
          }
        }
      }
      if (p.dot_out && valid) {
          float dv = fmaf(dot, p.range ? p.range[1] : 1.f, p.dot_b);     // undo the range scale (exact power of two)
          if (p.dot_affine) dv = dv * p.dot_affine[1] + p.dot_affine[0];
          p.dot_out[opix] = dv;
        }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// host launcher
// -------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

extern "C" int tpz_tc_conv_v1(const TpzTcConvArgs* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(a != nullptr, "tpz_tc_conv: null args");
  TPZ_CHECK(a->KC == 64 || a->KC == 32, "tpz_tc_conv: KC must be 32 or 64 (got %d)", a->KC);
  TPZ_CHECK(a->Co >= 16 && a->Co <= 256 && a->Co % 16 == 0, "tpz_tc_conv: Co=%d must be a multiple of 16 in [16,256]", a->Co);
  TPZ_CHECK(a->nsrc >= 1 && a->nsrc <= 2, "tpz_tc_conv: nsrc=%d", a->nsrc);
  TPZ_CHECK(a->nkb >= 1 && a->nkb <= TPZ_TC_MAX_KB, "tpz_tc_conv: nkb=%d exceeds %d", a->nkb, TPZ_TC_MAX_KB);
  TPZ_CHECK(a->TW * a->TH == 128 && a->TW % 8 == 0, "tpz_tc_conv: tile %dx%d must have 128 pixels, TW%%8==0", a->TW, a->TH);
  TPZ_CHECK(a->out != nullptr || a->dot_out != nullptr, "tpz_tc_conv: no output");
  TPZ_CHECK(a->out == nullptr || (a->out_ld % 8 == 0 && a->out_coff % 8 == 0), "tpz_tc_conv: output channel stride/offset must be multiples of 8");
  if (g_num_sms == 0) {
    int dev = 0;
    TPZ_CUDA(cudaGetDevice(&dev));
    TPZ_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }

  TPZ_CHECK(a->phase_sel == 0 && a->lattice_z <= 1, "tpz_tc_conv_v1: phase selection needs the halo-resident kernel");
  for (int s = 0; s < a->nsrc; ++s)
    TPZ_CHECK((a->src[s].lat == 0 || a->src[s].lat == a->lattice || a->lattice == 0) && !a->src[s].no_phase,
              "tpz_tc_conv_v1: mixed-resolution sources need the halo-resident kernel");
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < a->nsrc; ++s) {
    const TpzTcSrc& src = a->src[s];
    TPZ_CHECK(src.C % a->KC == 0, "tpz_tc_conv: source %d channels %d not a multiple of KC=%d", s, src.C, a->KC);
    TPZ_CHECK(src.ld % 8 == 0, "tpz_tc_conv: source %d channel stride %d must be a multiple of 8", s, src.ld);
    uint64_t dims[5] = {(uint64_t)src.C, (uint64_t)src.W, (uint64_t)src.H, (uint64_t)src.D, (uint64_t)src.N};
    uint64_t strides[4] = {(uint64_t)src.ld * 2, (uint64_t)src.ld * 2 * src.W, (uint64_t)src.ld * 2 * src.W * src.H,
                           (uint64_t)src.ld * 2 * src.W * src.H * src.D};
    uint32_t box[5] = {(uint32_t)a->KC, (uint32_t)a->TW, (uint32_t)a->TH, 1, 1};
    uint32_t es[5] = {1, 1, 1, 1, 1};
    int rc = tpz_encode_tmap(&p.tmA[s], src.ptr, 5, dims, strides, box, es, a->KC * 2);
    if (rc) return rc;
    p.org[s][0] = src.org[0]; p.org[s][1] = src.org[1]; p.org[s][2] = src.org[2];
  }
  {
    uint64_t dims[2] = {(uint64_t)a->KC, (uint64_t)a->nkb * a->Co};
    uint64_t strides[1] = {(uint64_t)a->KC * 2};
    uint32_t box[2] = {(uint32_t)a->KC, (uint32_t)a->Co};
    uint32_t es[2] = {1, 1};
    int rc = tpz_encode_tmap(&p.tmB, a->weights, 2, dims, strides, box, es, a->KC * 2);
    if (rc) return rc;
  }
  p.nsrc = a->nsrc;
  p.N = a->N; p.Do = a->Do; p.Ho = a->Ho; p.Wo = a->Wo; p.Co = a->Co;
  p.TW = a->TW; p.TH = a->TH;
  p.tiles_x = tpz_div_up(a->Wo, a->TW);
  p.tiles_y = tpz_div_up(a->Ho, a->TH);
  const long long nt = (long long)p.tiles_x * p.tiles_y * a->Do * a->N;
  TPZ_CHECK(nt > 0 && nt < (1ll << 31), "tpz_tc_conv: bad tile count %lld", nt);
  p.num_tiles = (int)nt;
  p.nkb = a->nkb;
  for (int i = 0; i < a->nkb; ++i) {
    p.kb[i] = a->kb[i];
    TPZ_CHECK(a->kb[i].src >= 0 && a->kb[i].src < a->nsrc, "tpz_tc_conv: k-block %d bad source", i);
  }
  p.bias = a->bias; p.neg_slope = a->neg_slope;
  p.res = reinterpret_cast<const __half*>(a->res); p.res_scale = a->res_scale; p.res_ld = a->res_ld;
  p.res_D = a->res_D; p.res_H = a->res_H; p.res_W = a->res_W;
  p.res_org[0] = a->res_org[0]; p.res_org[1] = a->res_org[1]; p.res_org[2] = a->res_org[2];
  TPZ_CHECK(a->res == nullptr || a->res_ld % 8 == 0, "tpz_tc_conv: residual channel stride must be a multiple of 8");
  p.out = reinterpret_cast<__half*>(a->out); p.out_ld = a->out_ld; p.out_coff = a->out_coff;
  p.dot_w = a->dot_w; p.dot_b = a->dot_b; p.dot_out = a->dot_out; p.dot_affine = a->dot_affine;
  p.oscale = a->oscale; p.range = a->range; p.out_lo = a->out_lo;
  TPZ_CHECK(a->out_lo % 8 == 0 && a->out_lo >= 0, "tpz_tc_conv: out_lo=%d must be a non-negative multiple of 8", a->out_lo);

  const int rowb = a->KC * 2;
  const int stage_bytes = 128 * rowb + a->Co * rowb;
  const int tail = 4096;
  const int budget = 227 * 1024 - 1024 - tail;
  int S = budget / stage_bytes;
  if (S > 8) S = 8;
  TPZ_CHECK(S >= 2, "tpz_tc_conv: stage too large");
  p.stages = S;
  int smem = S * stage_bytes + tail + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;  // force 1 CTA/SM: each CTA allocates all 512 TMEM columns

  int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  if (a->KC == 64) {
    TPZ_CUDA(cudaFuncSetAttribute(tc_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_conv_kernel<64><<<grid, kThreads, smem, stream>>>(p);
  } else {
    TPZ_CUDA(cudaFuncSetAttribute(tc_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_conv_kernel<32><<<grid, kThreads, smem, stream>>>(p);
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

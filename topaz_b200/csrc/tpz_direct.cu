// Direct (SIMT) kernels for the thin ends of the networks and for validation:
//   Cin=1 first convs, Cout=1 tails, a generic reference-quality conv, 2x max-pool, nearest upsample,
//   mean/std normalisation.  All fp32 math; activations fp16 channels-last.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {

__device__ __forceinline__ float act(float v, float slope) { return v > 0.f ? v : v * slope; }

// -------------------------------------------------------------------------------------------------
// Cin = 1 first conv.  Block = 256 threads = 16 x-groups (4 px each) x 16 rows -> 64x16 output tile of
// one (n,z) plane.  Input halo tile + all weights staged in smem; each thread keeps 4 px x 16 ch
// accumulators and walks the taps with a sliding register window along x.
// -------------------------------------------------------------------------------------------------
constexpr int FT_W = 64, FT_H = 16, FPX = 4, FCG = 16;

__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, int N, int D, int H, int W,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         int Co, int kd, int kh, int kw, int dil, int pad,
                                                         float slope, __half* __restrict__ out, int out_ld, int Do,
                                                         int Ho, int Wo, int CoPad, const float* __restrict__ range,
                                                         int out_lo) {
  extern __shared__ float sm[];
  const int tw = FT_W + (kw - 1) * dil, th = FT_H + (kh - 1) * dil;
  float* s_in = sm;                       // [kd][th][tw]
  float* s_w = sm + (size_t)kd * th * tw; // [kd*kh*kw][CoPad]
  const int ntaps = kd * kh * kw;
  const int plane = blockIdx.z;
  const int n = plane / Do, z = plane - n * Do;
  const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
  const float rs = range ? range[0] : 1.f;      // range guard (tpz_range_scale): input and bias scaled by a power of two

  for (int i = threadIdx.x; i < ntaps * CoPad; i += 256) {
    const int t = i / CoPad, c = i - t * CoPad;
    s_w[i] = (c < Co) ? w[(size_t)c * ntaps + t] : 0.f;
  }
  for (int i = threadIdx.x; i < kd * th * tw; i += 256) {
    const int q = i / (th * tw), r = i - q * th * tw;
    const int yy = r / tw, xx = r - yy * tw;
    const int gz = z - pad * (kd > 1) + q * dil, gy = y0 - pad + yy, gx = x0 - pad + xx;
    float v = 0.f;
    if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = x[(((size_t)n * D + gz) * H + gy) * W + gx] * rs;
    s_in[i] = v;
  }
  __syncthreads();

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lx = tx * FPX;
  for (int cg = 0; cg < CoPad; cg += FCG) {
    float acc[FPX][FCG];
#pragma unroll
    for (int pxi = 0; pxi < FPX; ++pxi)
#pragma unroll
      for (int c = 0; c < FCG; ++c) acc[pxi][c] = 0.f;
    for (int q = 0; q < kd; ++q) {
      for (int r = 0; r < kh; ++r) {
        const float* row = s_in + ((size_t)q * th + ty + r * dil) * tw + lx;
        const float* wr = s_w + (size_t)((q * kh + r) * kw) * CoPad + cg;
        for (int s = 0; s < kw; ++s) {
          float iv[FPX];
#pragma unroll
          for (int pxi = 0; pxi < FPX; ++pxi) iv[pxi] = row[pxi + s * dil];
          const float4* w4 = reinterpret_cast<const float4*>(wr + (size_t)s * CoPad);
#pragma unroll
          for (int c4 = 0; c4 < FCG / 4; ++c4) {
            const float4 wv = w4[c4];
#pragma unroll
            for (int pxi = 0; pxi < FPX; ++pxi) {
              acc[pxi][c4 * 4 + 0] = fmaf(iv[pxi], wv.x, acc[pxi][c4 * 4 + 0]);
              acc[pxi][c4 * 4 + 1] = fmaf(iv[pxi], wv.y, acc[pxi][c4 * 4 + 1]);
              acc[pxi][c4 * 4 + 2] = fmaf(iv[pxi], wv.z, acc[pxi][c4 * 4 + 2]);
              acc[pxi][c4 * 4 + 3] = fmaf(iv[pxi], wv.w, acc[pxi][c4 * 4 + 3]);
            }
          }
        }
      }
    }
    const int gy = y0 + ty;
    if (gy < Ho) {
#pragma unroll
      for (int pxi = 0; pxi < FPX; ++pxi) {
        const int gx = x0 + lx + pxi;
        if (gx < Wo) {
          __half* o = out + ((((size_t)n * Do + z) * Ho + gy) * Wo + gx) * out_ld + cg;
          uint4 u[2], ul[2];
          __half2* h = reinterpret_cast<__half2*>(u);
          __half2* hl = reinterpret_cast<__half2*>(ul);
#pragma unroll
          for (int c = 0; c < FCG; c += 2) {
            const float b0 = (cg + c < Co && bias) ? bias[cg + c] * rs : 0.f;
            const float b1 = (cg + c + 1 < Co && bias) ? bias[cg + c + 1] * rs : 0.f;
            float v0 = act(acc[pxi][c] + b0, slope), v1 = act(acc[pxi][c + 1] + b1, slope);
            if (cg + c >= Co) v0 = 0.f;
            if (cg + c + 1 >= Co) v1 = 0.f;
            h[c / 2] = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(h[c / 2]);
            hl[c / 2] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          }
          reinterpret_cast<uint4*>(o)[0] = u[0];
          reinterpret_cast<uint4*>(o)[1] = u[1];
          if (out_lo > 0) {                         // strict mode: fp16 rounding residuals at channel c + out_lo
            reinterpret_cast<uint4*>(o + out_lo)[0] = ul[0];
            reinterpret_cast<uint4*>(o + out_lo)[1] = ul[1];
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// im2col of a single-channel image for the tensor-core path of Cin = 1 convs: out[n,y,x,t] = x[n, y-pad+r, x-pad+s]
// for tap t = r*k+s (zero outside the image), channels [k*k, ld) = 0.  HBM-bound: reads 4 B, writes 2*ld B per pixel.
// Block = 32x8 output pixels; the (8+k-1) x (32+k-1) input patch is staged in shared memory.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ x, int N, int H, int W, int k, int pad,
                                                           __half* __restrict__ out, int ld, int Ho, int Wo,
                                                           const float* __restrict__ range, int out_lo) {
  extern __shared__ float sm_i2c[];
  const float rs = range ? range[0] : 1.f;
  const int old = out_lo > 0 ? 2 * ld : ld;       // channel stride of the output tensor (strict mode: hi | lo halves)
  const int tw = 32 + k - 1, th = 8 + k - 1;
  int* s_off = reinterpret_cast<int*>(sm_i2c + tw * th);      // tap -> offset inside the patch (or -1 for zero padding)
  const int n = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const int taps = k * k;
  for (int t = threadIdx.x; t < ld; t += 256) s_off[t] = (t < taps) ? (t / k) * tw + (t % k) : -1;
  for (int i = threadIdx.x; i < tw * th; i += 256) {
    const int yy = i / tw, xx = i - yy * tw;
    const int gy = y0 - pad + yy, gx = x0 - pad + xx;
    sm_i2c[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[((size_t)n * H + gy) * W + gx] * rs : 0.f;
  }
  __syncthreads();
  // consecutive threads write consecutive 16-byte chunks of a pixel's channel vector -> 512 B contiguous per warp store
  const int nch = ld >> 3;                    // 16-byte chunks per pixel (power of two for ld = 32/64/128)
  const int sh = 31 - __clz(nch);
  for (int idx = threadIdx.x; idx < 256 * nch; idx += 256) {
    const int pix = idx >> sh, c = (idx & (nch - 1)) * 8;
    const int lx = pix & 31, ly = pix >> 5;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gx >= Wo || gy >= Ho) continue;
    const float* bp = sm_i2c + ly * tw + lx;
    uint4 u, ul;
    __half2* h = reinterpret_cast<__half2*>(&u);
    __half2* hl = reinterpret_cast<__half2*>(&ul);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o0 = s_off[c + 2 * e], o1 = s_off[c + 2 * e + 1];
      const float f0 = o0 >= 0 ? bp[o0] : 0.f, f1 = o1 >= 0 ? bp[o1] : 0.f;
      h[e] = __floats2half2_rn(f0, f1);
      const float2 hf = __half22float2(h[e]);
      hl[e] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
    }
    __half* op = out + (((size_t)n * Ho + gy) * Wo + gx) * old + c;
    *reinterpret_cast<uint4*>(op) = u;
    if (out_lo > 0) *reinterpret_cast<uint4*>(op + out_lo) = ul;
  }
}

// -------------------------------------------------------------------------------------------------
// 3-D im2col of a single-channel volume ('same' zero padding): out[n][z][y][x][t] = x[n][z+dz-p][y+dy-p][x+dx-p],
// t = (dz*k + dy)*k + dx, channels >= k^3 zero.  Feeds the raw-volume slice of UDenoiseNet3D's dec1.0 (the concat
// `[upsampled, x]` of denoising/models.py:555) to the tensor-core conv as a second source; it replaces a one-hot fp32
// convolution that cost 0.55 ms per 192^3 patch.  One thread per (voxel, 16-channel piece): 32-byte stores, the
// neighbourhood reads hit L1/L2.
// -------------------------------------------------------------------------------------------------
template <int KT>   // KT > 0: compile-time kernel size (tap offsets become constants); KT = 0: runtime k
__global__ void im2col3d_first_kernel(const float* __restrict__ x, int N, int D, int H, int W, int k_rt, int pad,
                                      __half* __restrict__ out, int ld, const float* __restrict__ range, int out_lo) {
  const int k = KT > 0 ? KT : k_rt;
  const float rs = range ? range[0] : 1.f;
  const int old = out_lo > 0 ? 2 * ld : ld;
  const int pieces = ld >> 4;
  const size_t total = (size_t)N * D * H * W * pieces;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int piece = (int)(idx % pieces);
  size_t v = idx / pieces;
  const int gx = v % W; v /= W;
  const int gy = v % H; v /= H;
  const int gz = v % D;
  const int n = (int)(v / D);
  const int taps = k * k * k;
  const float* vol = x + (size_t)n * D * H * W;
  uint32_t pk[8], pl[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float f[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = piece * 16 + 2 * e + h;
      float val = 0.f;
      if (t < taps) {
        const int dx = t % k, dy = (t / k) % k, dz = t / (k * k);
        const int ix = gx + dx - pad, iy = gy + dy - pad, iz = gz + dz - pad;
        if (ix >= 0 && ix < W && iy >= 0 && iy < H && iz >= 0 && iz < D) val = __ldg(vol + ((size_t)iz * H + iy) * W + ix) * rs;
      }
      f[h] = val;
    }
    const __half2 hh = __floats2half2_rn(f[0], f[1]);
    pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(f[0] - hf.x, f[1] - hf.y);
    pl[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  __half* op = out + ((((size_t)n * D + gz) * H + gy) * W + gx) * old + piece * 16;
  ptx::st_global_256(op, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
  if (out_lo > 0) ptx::st_global_256(op + out_lo, pl[0], pl[1], pl[2], pl[3], pl[4], pl[5], pl[6], pl[7]);
}

// -------------------------------------------------------------------------------------------------
// Cout = 1 tail conv: one thread per output pixel, 8-channel (16 B) vector loads.
// -------------------------------------------------------------------------------------------------
__global__ void conv_last_kernel(const __half* __restrict__ x, int N, int D, int H, int W, int C, int ld,
                                 const float* __restrict__ w /*[taps][C]*/, float bias, int kd, int kh, int kw,
                                 int dil, int pad, float oscale, float oshift, const float* __restrict__ stats,
                                 float* __restrict__ out, const float* __restrict__ range) {
  const size_t total = (size_t)N * D * H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gx = idx % W;
  size_t r = idx / W;
  const int gy = r % H; r /= H;
  const int gz = r % D;
  const int n = r / D;
  float acc = 0.f;
  const int pz = (kd > 1) ? pad : 0;
  for (int q = 0; q < kd; ++q) {
    const int iz = gz - pz + q * dil;
    if (iz < 0 || iz >= D) continue;
    for (int rr = 0; rr < kh; ++rr) {
      const int iy = gy - pad + rr * dil;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < kw; ++s) {
        const int ix = gx - pad + s * dil;
        if (ix < 0 || ix >= W) continue;
        const __half* px = x + ((((size_t)n * D + iz) * H + iy) * W + ix) * ld;
        const float* wt = w + (size_t)((q * kh + rr) * kw + s) * C;
        for (int c = 0; c < C; c += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(px + c);
          const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            acc = fmaf(f.x, __ldg(wt + c + 2 * e), acc);
            acc = fmaf(f.y, __ldg(wt + c + 2 * e + 1), acc);
          }
        }
      }
    }
  }
  float v = (fmaf(acc, range ? range[1] : 1.f, bias)) * oscale + oshift;      // undo the range scale (power of two: exact)
  if (stats) v = v * stats[1] + stats[0];
  out[idx] = v;
}

// -------------------------------------------------------------------------------------------------
// Cout = 1 tail conv, 2-D, tiled: the U-Net's dec1.4 (32 -> 1, 5x5) has 1600 FLOP/px -- far too little for the tensor
// core path, whose issue loop and 4 KB-per-MMA A-operand reads made it cost as much as a 64-channel layer (0.58 ms per
// 2048^2 patch, 15 % of the network).  Here a block stages a (TY+K-1) x (TX+K-1) x C fp16 window in shared memory (pixel
// stride C*2+16 bytes: conflict-free 16-byte loads across a warp), each thread owns 4 vertically adjacent outputs of one
// column and walks the (s, channel-chunk) pairs with the 5 x 8 weights of that pair held in registers, so every staged
// value is loaded once per s and used by up to 4 outputs.  fp32 accumulation, fused bias + de-normalisation.
// -------------------------------------------------------------------------------------------------
template <int C, int K>
__global__ void __launch_bounds__(128) conv_last_tiled_kernel(const __half* __restrict__ x, int N, int H, int W, int ld,
                                                              const float* __restrict__ w /*[K*K][C]*/, float bias,
                                                              float oscale, float oshift, const float* __restrict__ stats,
                                                              float* __restrict__ out, int tiles_x, int tiles_y,
                                                              const float* __restrict__ range) {
  constexpr int TX = 32, TY = 16, PX = TX + K - 1, PY = TY + K - 1;
  constexpr int PSTRIDE = C * 2 + 16;          // bytes per staged pixel
  constexpr int CH = C / 8;                    // 16-byte chunks per pixel
  extern __shared__ __align__(16) unsigned char smem_last[];
  unsigned char* tile = smem_last;                                            // [PY][PX][PSTRIDE]
  float* sw = reinterpret_cast<float*>(smem_last + PY * PX * PSTRIDE);        // [K*K][C]
  const int tid = threadIdx.x;
  for (int i = tid; i < K * K * C; i += 128) sw[i] = w[i];
  const int tx = tid & 31, tg = tid >> 5;      // column inside the tile, group of 4 rows
  const long long ntiles = (long long)tiles_x * tiles_y * N;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int n = (int)(t / ((long long)tiles_x * tiles_y));
    const int tr = (int)(t % ((long long)tiles_x * tiles_y));
    const int x0 = (tr % tiles_x) * TX, y0 = (tr / tiles_x) * TY;
    __syncthreads();                            // previous tile fully consumed (also orders the weight staging)
    for (int i = tid; i < PY * PX * CH; i += 128) {
      const int c = i % CH, p = i / CH;
      const int wx = p % PX, wy = p / PX;
      const int ix = x0 + wx - K / 2, iy = y0 + wy - K / 2;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ix >= 0 && ix < W && iy >= 0 && iy < H) v = *reinterpret_cast<const uint4*>(x + (((size_t)n * H + iy) * W + ix) * ld + c * 8);
      *reinterpret_cast<uint4*>(tile + (size_t)(wy * PX + wx) * PSTRIDE + c * 16) = v;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < K; ++s) {
#pragma unroll 1
      for (int c = 0; c < CH; ++c) {
        float wr[K][8];                         // weights of taps (r, s), channels 8c..8c+7
#pragma unroll
        for (int r = 0; r < K; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(sw + (r * K + s) * C + c * 8);
          const float4 b = *reinterpret_cast<const float4*>(sw + (r * K + s) * C + c * 8 + 4);
          wr[r][0] = a.x; wr[r][1] = a.y; wr[r][2] = a.z; wr[r][3] = a.w;
          wr[r][4] = b.x; wr[r][5] = b.y; wr[r][6] = b.z; wr[r][7] = b.w;
        }
        const unsigned char* col = tile + (size_t)((tg * 4) * PX + tx + s) * PSTRIDE + c * 16;
#pragma unroll
        for (int row = 0; row < 4 + K - 1; ++row) {
          const uint4 u = *reinterpret_cast<const uint4*>(col + (size_t)row * PX * PSTRIDE);
          const __half2* h = reinterpret_cast<const __half2*>(&u);
          float f[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 q = __half22float2(h[e]); f[2 * e] = q.x; f[2 * e + 1] = q.y; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = row - j;              // output j (row tg*4 + j) sees this input row through tap row r
            if (r >= 0 && r < K) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[j] = fmaf(f[e], wr[r][e], acc[j]);
            }
          }
        }
      }
    }
    const int ox = x0 + tx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int oy = y0 + tg * 4 + j;
      if (ox < W && oy < H) {
        float v = (fmaf(acc[j], range ? range[1] : 1.f, bias)) * oscale + oshift;
        if (stats) v = v * stats[1] + stats[0];
        out[((size_t)n * H + oy) * W + ox] = v;
      }
    }
  }
}

template <int C, int K>
static int launch_conv_last_tiled(const __half* x, int N, int H, int W, int ld, const float* w, float bias, float oscale,
                                  float oshift, const float* stats, float* out, const float* range, cudaStream_t stream) {
  constexpr int TX = 32, TY = 16;
  const int smem = (TY + K - 1) * (TX + K - 1) * (C * 2 + 16) + K * K * C * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(conv_last_tiled_kernel<C, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_x = tpz_div_up(W, TX), tiles_y = tpz_div_up(H, TY);
  const long long ntiles = (long long)tiles_x * tiles_y * N;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = (220 * 1024) / (smem + 1024) > 8 ? 8 : (220 * 1024) / (smem + 1024);
  const long long cap = (long long)sms * (per_sm < 1 ? 1 : per_sm);
  conv_last_tiled_kernel<C, K><<<(int)(ntiles < cap ? ntiles : cap), 128, smem, stream>>>(x, N, H, W, ld, w, bias, oscale,
                                                                                       oshift, stats, out, tiles_x, tiles_y, range);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Cout = 1 tail conv, 3-D (UDenoiseNet3D dec1.4: 32 -> 1, 3x3x3, denoising/models.py:505).  On the tensor-core path this layer
// is an N = 16 GEMM: 54 MMAs per 128 voxels for 0.06 TFLOP (1.1 ms per 192^3 patch, 3.0 ms with split operands).  Here: fp32
// CUDA-core math on fp16 inputs -- exact products, so split (hi, lo) inputs are simply 2*C channels with the weights
// repeated, and no weight split is needed.  A block owns a 16 x 32 (y, x) tile and marches along z: every input plane is
// staged ONCE (16 channels at a time, 48-byte voxel stride = conflict-free 16-byte loads) and feeds the three output planes
// it touches; a thread owns 4 y-adjacent outputs x 3 z-planes in flight = 12 accumulators, the 27 x 8 weights of one
// (x-tap, 8-channel) slice live in registers, so each 16-byte shared load feeds up to 72 FMAs.
// -------------------------------------------------------------------------------------------------
template <int C>       // stored input channels that carry data (32: fp16; 64: hi | lo halves, weights repeated)
__global__ void __launch_bounds__(128, 3) conv_last3d_tiled_kernel(const __half* __restrict__ x, int N, int D, int H, int W, int ld,
                                                                   const float* __restrict__ w /*[27][C]*/, float bias,
                                                                   const float* __restrict__ stats, const float* __restrict__ range,
                                                                   float* __restrict__ out, int tiles_x, int tiles_y, int zchunks, int TZ) {
  constexpr int TX = 32, TY = 16, PX = TX + 2, PY = TY + 2, PST = 48;     // staged voxel: 16 channels (32 B) + 16 B pad
  extern __shared__ __align__(16) unsigned char smem_l3[];
  unsigned char* tile = smem_l3;                                          // [PY][PX][PST]
  float* sw = reinterpret_cast<float*>(smem_l3 + PY * PX * PST);          // [27][C]
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * C; i += 128) sw[i] = w[i];
  const int tx = tid & 31, tg = tid >> 5;
  long long b = blockIdx.x;
  const int zc = (int)(b % zchunks); b /= zchunks;
  const int tr = (int)(b % ((long long)tiles_x * tiles_y));
  const int n = (int)(b / ((long long)tiles_x * tiles_y));
  const int x0 = (tr % tiles_x) * TX, y0 = (tr / tiles_x) * TY;
  const int z0 = zc * TZ, z1 = min(D, z0 + TZ);
  const float inv_s = range ? range[1] : 1.f;
  float acc[3][4];        // acc[k][j]: output plane (zi - 1 + k) for the input plane zi being processed, rows tg*4 + j
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
  for (int zi = z0 - 1; zi <= z1; ++zi) {
    if (zi >= 0 && zi < D) {
#pragma unroll 1
      for (int c16 = 0; c16 < C; c16 += 16) {
        __syncthreads();                                   // previous slice consumed (first pass: orders the weight staging)
        for (int i = tid; i < PY * PX * 2; i += 128) {
          const int half = i & 1, p = i >> 1;
          const int wx = p % PX, wy = p / PX;
          const int ix = x0 + wx - 1, iy = y0 + wy - 1;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (ix >= 0 && ix < W && iy >= 0 && iy < H)
            v = *reinterpret_cast<const uint4*>(x + ((((size_t)n * D + zi) * H + iy) * W + ix) * ld + c16 + half * 8);
          *reinterpret_cast<uint4*>(tile + (size_t)p * PST + half * 16) = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            float wr[3][3][8];                             // [z-tap q][y-tap r][channel]
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const float* wp = sw + ((q * 3 + r) * 3 + s) * C + c16 + sub * 8;
                const float4 a = *reinterpret_cast<const float4*>(wp);
                const float4 bq = *reinterpret_cast<const float4*>(wp + 4);
                wr[q][r][0] = a.x; wr[q][r][1] = a.y; wr[q][r][2] = a.z; wr[q][r][3] = a.w;
                wr[q][r][4] = bq.x; wr[q][r][5] = bq.y; wr[q][r][6] = bq.z; wr[q][r][7] = bq.w;
              }
            const unsigned char* col = tile + (size_t)((tg * 4) * PX + tx + s) * PST + sub * 16;
#pragma unroll
            for (int row = 0; row < 6; ++row) {
              const uint4 u = *reinterpret_cast<const uint4*>(col + (size_t)row * PX * PST);
              const __half2* h = reinterpret_cast<const __half2*>(&u);
              float f[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 t2 = __half22float2(h[e]); f[2 * e] = t2.x; f[2 * e + 1] = t2.y; }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int r = row - j;                     // output row j sees this input row through y-tap r
                if (r >= 0 && r < 3) {
#pragma unroll
                  for (int k = 0; k < 3; ++k) {            // output plane zi - 1 + k sees input plane zi through z-tap q = 2 - k
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[k][j] = fmaf(f[e], wr[2 - k][r][e], acc[k][j]);
                  }
                }
              }
            }
          }
        }
      }
    }
    // output plane zi - 1 is complete once input plane zi has been added
    const int zo = zi - 1;
    if (zo >= z0 && zo < z1) {
      const int ox = x0 + tx;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int oy = y0 + tg * 4 + j;
        if (ox < W && oy < H) {
          float v = fmaf(acc[0][j], inv_s, bias);
          if (stats) v = v * stats[1] + stats[0];
          out[(((size_t)n * D + zo) * H + oy) * W + ox] = v;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][j] = acc[1][j]; acc[1][j] = acc[2][j]; acc[2][j] = 0.f; }
  }
}

template <int C>
static int launch_conv_last3d(const __half* x, int N, int D, int H, int W, int ld, const float* w, float bias, const float* stats,
                              const float* range, float* out, cudaStream_t stream) {
  constexpr int TX = 32, TY = 16;
  const int smem = (TY + 2) * (TX + 2) * 48 + 27 * C * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(conv_last3d_tiled_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_x = tpz_div_up(W, TX), tiles_y = tpz_div_up(H, TY);
  // z chunks: enough blocks for ~2 waves of 3 blocks per SM, at most 1/8 redundant halo planes
  int TZ = D;
  while (TZ > 16 && (long long)tiles_x * tiles_y * N * tpz_div_up(D, TZ) < 148 * 6) TZ = (TZ + 1) / 2;
  const int zchunks = tpz_div_up(D, TZ);
  const long long blocks = (long long)tiles_x * tiles_y * N * zchunks;
  TPZ_CHECK(blocks > 0 && blocks < (1ll << 31), "tpz_conv_last: bad block count %lld", blocks);
  conv_last3d_tiled_kernel<C><<<(unsigned)blocks, 128, smem, stream>>>(x, N, D, H, W, ld, w, bias, stats, range, out, tiles_x,
                                                                         tiles_y, zchunks, TZ);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Generic conv (validation / uncovered shapes): one thread per (output pixel, 8 output channels).
// -------------------------------------------------------------------------------------------------
__global__ void conv_generic_kernel(const __half* __restrict__ x0, int C0, int ld0, const __half* __restrict__ x1,
                                    int C1, int ld1, int N, int D, int H, int W, const float* __restrict__ w,
                                    const float* __restrict__ bias, int Co, int kd, int kh, int kw, int stride,
                                    int dil, int pad, float slope, const __half* __restrict__ res, int res_ld,
                                    int res_org, __half* __restrict__ out, int out_ld, int Do, int Ho, int Wo) {
  const int cgs = (Co + 7) / 8;
  const size_t total = (size_t)N * Do * Ho * Wo * cgs;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = idx % cgs;
  size_t r = idx / cgs;
  const int ox = r % Wo; r /= Wo;
  const int oy = r % Ho; r /= Ho;
  const int oz = r % Do;
  const int n = r / Do;
  const int Ci = C0 + C1, taps = kd * kh * kw;
  const int pz = (kd > 1) ? pad : 0, sz = (kd > 1) ? stride : 1;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int q = 0; q < kd; ++q) {
    const int iz = oz * sz - pz + q * dil;
    if (iz < 0 || iz >= D) continue;
    for (int rr = 0; rr < kh; ++rr) {
      const int iy = oy * stride - pad + rr * dil;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < kw; ++s) {
        const int ix = ox * stride - pad + s * dil;
        if (ix < 0 || ix >= W) continue;
        const size_t pix = (((size_t)n * D + iz) * H + iy) * W + ix;
        const int t = (q * kh + rr) * kw + s;
        for (int c = 0; c < Ci; ++c) {
          const float v = (c < C0) ? __half2float(x0[pix * ld0 + c]) : __half2float(x1[pix * ld1 + (c - C0)]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int co = cg * 8 + j;
            if (co < Co) acc[j] = fmaf(v, __ldg(w + ((size_t)co * Ci + c) * taps + t), acc[j]);
          }
        }
      }
    }
  }
  const size_t opix = (((size_t)n * Do + oz) * Ho + oy) * Wo + ox;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = cg * 8 + j;
    if (co >= Co) break;
    float v = acc[j] + (bias ? bias[co] : 0.f);
    if (res) {
      const size_t rp = (((size_t)n * D + (oz * sz + (kd > 1 ? res_org : 0))) * H + (oy * stride + res_org)) * W +
                        (ox * stride + res_org);
      v += __half2float(res[rp * res_ld + co]);
    }
    out[opix * out_ld + co] = __float2half_rn(act(v, slope));
  }
}

// -------------------------------------------------------------------------------------------------
// 2x max pool (floor) and nearest upsample on fp16 NDHWC, 8-channel vectors.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __half2* ra = reinterpret_cast<__half2*>(&a);
  __half2* rb = reinterpret_cast<__half2*>(&b);
  __half2* rr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) rr[e] = __hmax2(ra[e], rb[e]);
  return r;
}

__global__ void maxpool2_kernel(const __half* __restrict__ x, int N, int D, int H, int W, int C, int ld, int dims,
                                __half* __restrict__ out, int out_ld, int Do, int Ho, int Wo, int lo_off) {
  const int cv = C / 8;
  const size_t total = (size_t)N * Do * Ho * Wo * cv;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (idx % cv) * 8;
  size_t r = idx / cv;
  const int ox = r % Wo; r /= Wo;
  const int oy = r % Ho; r /= Ho;
  const int oz = r % Do;
  const int n = r / Do;
  const int nz = (dims == 3) ? 2 : 1;
  if (lo_off > 0) {
    // strict mode: values are (hi, lo) fp16 pairs; the maximum is taken over hi + lo (exact in fp32) and re-split
    float mv[8];
    bool first = true;
    for (int q = 0; q < nz; ++q)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int iz = (dims == 3) ? oz * 2 + q : oz;
          const __half* px = x + ((((size_t)n * D + iz) * H + (oy * 2 + a)) * W + (ox * 2 + b)) * ld + c;
          const uint4 vh = *reinterpret_cast<const uint4*>(px);
          const uint4 vl = *reinterpret_cast<const uint4*>(px + lo_off);
          const __half2* hh = reinterpret_cast<const __half2*>(&vh);
          const __half2* hl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fh = __half22float2(hh[e]), fl = __half22float2(hl[e]);
            const float v0 = fh.x + fl.x, v1 = fh.y + fl.y;
            mv[2 * e] = first ? v0 : fmaxf(mv[2 * e], v0);
            mv[2 * e + 1] = first ? v1 : fmaxf(mv[2 * e + 1], v1);
          }
          first = false;
        }
    uint4 oh, ol;
    __half2* ph = reinterpret_cast<__half2*>(&oh);
    __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ph[e] = __floats2half2_rn(mv[2 * e], mv[2 * e + 1]);
      const float2 hf = __half22float2(ph[e]);
      pl[e] = __floats2half2_rn(mv[2 * e] - hf.x, mv[2 * e + 1] - hf.y);
    }
    __half* op = out + ((((size_t)n * Do + oz) * Ho + oy) * Wo + ox) * out_ld + c;
    *reinterpret_cast<uint4*>(op) = oh;
    *reinterpret_cast<uint4*>(op + lo_off) = ol;
    return;
  }
  uint4 m;
  bool first = true;
  for (int q = 0; q < nz; ++q)
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int iz = (dims == 3) ? oz * 2 + q : oz;
        const uint4 v = *reinterpret_cast<const uint4*>(
            x + ((((size_t)n * D + iz) * H + (oy * 2 + a)) * W + (ox * 2 + b)) * ld + c);
        m = first ? v : hmax8(m, v);
        first = false;
      }
  *reinterpret_cast<uint4*>(out + ((((size_t)n * Do + oz) * Ho + oy) * Wo + ox) * out_ld + c) = m;
}

// torch nearest: src = min(int(floorf(dst * (float)in / out)), in - 1)   (F.interpolate(size=...), ATen
// UpSample.h nearest_neighbor_compute_source_index)
__device__ __forceinline__ int nn_src(int dst, int in, int out) {
  const float scale = (float)in / (float)out;
  const int s = (int)floorf(dst * scale);
  return s < in - 1 ? s : in - 1;
}

__global__ void upsample_kernel(const __half* __restrict__ x, int N, int D, int H, int W, int C, int ld, int Do,
                                int Ho, int Wo, __half* __restrict__ out, int out_ld, int out_coff) {
  const int cv = C / 8;
  const size_t total = (size_t)N * Do * Ho * Wo * cv;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (idx % cv) * 8;
  size_t r = idx / cv;
  const int ox = r % Wo; r /= Wo;
  const int oy = r % Ho; r /= Ho;
  const int oz = r % Do;
  const int n = r / Do;
  const int iz = nn_src(oz, D, Do), iy = nn_src(oy, H, Ho), ix = nn_src(ox, W, Wo);
  const uint4 v = *reinterpret_cast<const uint4*>(x + ((((size_t)n * D + iz) * H + iy) * W + ix) * ld + c);
  *reinterpret_cast<uint4*>(out + ((((size_t)n * Do + oz) * Ho + oy) * Wo + ox) * out_ld + out_coff + c) = v;
}

// -------------------------------------------------------------------------------------------------
// mean / std (fp64 accumulation) and affine (de)normalisation
// -------------------------------------------------------------------------------------------------
__global__ void zero2_kernel(double* w) { if (threadIdx.x < 2) w[threadIdx.x] = 0.0; }

__global__ void sums_kernel(const float* __restrict__ x, long long n, double* __restrict__ work) {
  double s = 0.0, ss = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    s += v; ss += v * v;
  }
  __shared__ double sh[2][32];
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(~0u, s, o); ss += __shfl_xor_sync(~0u, ss, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.0;
    ss = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(~0u, s, o); ss += __shfl_xor_sync(~0u, ss, o); }
    if (threadIdx.x == 0) { atomicAdd(&work[0], s); atomicAdd(&work[1], ss); }
  }
}

// second pass for the variance around the mean (numerically safe for large offsets, e.g. raw micrographs)
__global__ void var_kernel(const float* __restrict__ x, long long n, double* __restrict__ work) {
  const double mean = work[0] / (double)n;
  double ss = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = (double)x[i] - mean;
    ss += d * d;
  }
  __shared__ double sh[32];
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(~0u, ss, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    ss = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(~0u, ss, o);
    if (threadIdx.x == 0) atomicAdd(&work[2], ss);
  }
}

__global__ void finalize_kernel(const double* work, long long n, int unbiased, float* stats) {
  const double mean = work[0] / (double)n;
  const double var = work[2] / (double)(unbiased ? n - 1 : n);
  stats[0] = (float)mean;
  stats[1] = (float)sqrt(var);
}

__global__ void affine_kernel(const float* __restrict__ x, long long n, const float* __restrict__ stats, int inverse,
                              float* __restrict__ y) {
  const float mu = stats[0], sd = stats[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = inverse ? x[i] * sd + mu : (x[i] - mu) / sd;
}

// dense 1->1 "same" convolution with a small fp32 filter (GaussianDenoise, topaz/filters.py:62-79): one thread per
// output voxel, filter in shared memory, zero padding.  Memory/L1-bound; optional pre/post filter of the denoise path.
__global__ void filter_f32_kernel(const float* __restrict__ x, int N, int D, int H, int W, const float* __restrict__ f,
                                  int kd, int kh, int kw, float bias, float* __restrict__ y) {
  extern __shared__ float s_f[];
  const int taps = kd * kh * kw;
  for (int i = threadIdx.x; i < taps; i += blockDim.x) s_f[i] = f[i];
  __syncthreads();
  const size_t total = (size_t)N * D * H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gx = idx % W; size_t r = idx / W;
  const int gy = r % H; r /= H;
  const int gz = r % D; const int n = r / D;
  float acc = bias;
  for (int q = 0; q < kd; ++q) {
    const int iz = gz + q - kd / 2;
    if (iz < 0 || iz >= D) continue;
    for (int a = 0; a < kh; ++a) {
      const int iy = gy + a - kh / 2;
      if (iy < 0 || iy >= H) continue;
      const float* row = x + (((size_t)n * D + iz) * H + iy) * W;
      const float* fr = s_f + (q * kh + a) * kw;
      for (int b = 0; b < kw; ++b) {
        const int ix = gx + b - kw / 2;
        if (ix >= 0 && ix < W) acc = fmaf(row[ix], fr[b], acc);
      }
    }
  }
  y[idx] = acc;
}

// range guard: max|x| -> power-of-two scale (see tpz_range_scale in the header)
__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ work) {
  unsigned m = 0;            // bit pattern of |x|: ordered like the value for finite x; inf / NaN compare above every finite value
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(x[i]) & 0x7fffffffu);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(work, m);
}
__global__ void range_finalize_kernel(unsigned* work, float* range) {
  const unsigned bits = *work;
  *work = 0;                                        // ready for the next call on the same scratch word
  float s = 1.f;
  if (bits != 0 && bits < 0x7f800000u) {            // finite, non-zero maximum
    const float amax = __uint_as_float(bits);
    if (amax > 64.f) {                               // only ever scale DOWN: biases are scaled with the activations, so scaling a
      int e;                                         // tiny input up would blow the (then dominant) bias terms out of range
      frexpf(amax, &e);                              // amax = m * 2^e, m in [0.5, 1)  ->  amax * 2^(3-e) in [4, 8)
      int k = 3 - e;
      k = k < -100 ? -100 : k;
      s = ldexpf(1.f, k);
    }
  }
  range[0] = s;
  range[1] = 1.f / s;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ x, long long n, __half* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(x[i]);
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define HP(p) reinterpret_cast<__half*>(p)
#define HCP(p) reinterpret_cast<const __half*>(p)

extern "C" int tpz_conv_first(const float* x, int N, int D, int H, int W, const float* w, const float* bias, int Co,
                              int kd, int kh, int kw, int dil, int pad, float neg_slope, int pool, tpz_half* out,
                              int out_ld, const float* range, int out_lo, void* stream) {
  TPZ_CHECK(pool == 1, "tpz_conv_first: fused pooling not available (pool=%d)", pool);
  TPZ_CHECK(out_ld % 16 == 0, "tpz_conv_first: out_ld=%d must be a multiple of 16", out_ld);
  TPZ_CHECK(out_lo >= 0 && (out_lo == 0 || 2 * out_lo == out_ld), "tpz_conv_first: out_lo=%d must be 0 or out_ld/2", out_lo);
  const int CoPad = out_lo > 0 ? out_lo : out_ld;  // channels [Co, CoPad) are written as zeros (channel padding of the fp16 layout)
  TPZ_CHECK(Co <= CoPad, "tpz_conv_first: Co=%d exceeds the stored channels %d", Co, CoPad);
  const int Do = (kd > 1) ? D + 2 * pad - (kd - 1) * dil : D;
  const int Ho = H + 2 * pad - (kh - 1) * dil, Wo = W + 2 * pad - (kw - 1) * dil;
  TPZ_CHECK(Do > 0 && Ho > 0 && Wo > 0, "tpz_conv_first: empty output");
  const int tw = FT_W + (kw - 1) * dil, th = FT_H + (kh - 1) * dil;
  const size_t smem = ((size_t)kd * th * tw + (size_t)kd * kh * kw * CoPad) * sizeof(float);
  TPZ_CHECK(smem <= 220 * 1024, "tpz_conv_first: filter too large for shared memory (%zu B)", smem);
  TPZ_CUDA(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(tpz_div_up(Wo, FT_W), tpz_div_up(Ho, FT_H), N * Do);
  conv_first_kernel<<<grid, 256, smem, ST(stream)>>>(x, N, D, H, W, w, bias, Co, kd, kh, kw, dil, pad, neg_slope,
                                                     HP(out), out_ld, Do, Ho, Wo, CoPad, range, out_lo);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_im2col_first(const float* x, int N, int H, int W, int k, int pad, tpz_half* out, int ld, const float* range,
                                int out_lo, void* stream) {
  TPZ_CHECK(out_lo == 0 || out_lo == ld, "tpz_im2col_first: out_lo=%d must be 0 or ld", out_lo);
  TPZ_CHECK(k >= 1 && k * k <= ld && ld % 8 == 0, "tpz_im2col_first: k=%d taps do not fit ld=%d", k, ld);
  const int Ho = H + 2 * pad - (k - 1), Wo = W + 2 * pad - (k - 1);
  TPZ_CHECK(Ho > 0 && Wo > 0, "tpz_im2col_first: empty output");
  TPZ_CHECK((ld & (ld - 1)) == 0, "tpz_im2col_first: ld=%d must be a power of two", ld);
  const size_t smem = (size_t)(32 + k - 1) * (8 + k - 1) * sizeof(float) + (size_t)ld * sizeof(int);
  dim3 grid(tpz_div_up(Wo, 32), tpz_div_up(Ho, 8), N);
  im2col_first_kernel<<<grid, 256, smem, ST(stream)>>>(x, N, H, W, k, pad, HP(out), ld, Ho, Wo, range, out_lo);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_im2col3d_first(const float* x, int N, int D, int H, int W, int k, int pad, tpz_half* out, int ld,
                                  const float* range, int out_lo, void* stream) {
  TPZ_CHECK(out_lo == 0 || out_lo == ld, "tpz_im2col3d_first: out_lo=%d must be 0 or ld", out_lo);
  TPZ_CHECK(k >= 1 && k * k * k <= ld && ld % 16 == 0 && pad == k / 2, "tpz_im2col3d_first: k=%d pad=%d ld=%d", k, pad, ld);
  const size_t total = (size_t)N * D * H * W * (ld / 16);
  if (total == 0) return 0;
  if (k == 3) im2col3d_first_kernel<3><<<tpz_div_up(total, 256), 256, 0, ST(stream)>>>(x, N, D, H, W, k, pad, HP(out), ld, range, out_lo);
  else im2col3d_first_kernel<0><<<tpz_div_up(total, 256), 256, 0, ST(stream)>>>(x, N, D, H, W, k, pad, HP(out), ld, range, out_lo);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_last(const tpz_half* x, int N, int D, int H, int W, int C, int ld, const float* w, float bias,
                             int kd, int kh, int kw, int dil, int pad, float out_scale, float out_shift,
                             const float* affine_stats, float* out, const float* range, void* stream) {
  TPZ_CHECK(C % 8 == 0 && ld % 8 == 0, "tpz_conv_last: C=%d / ld=%d must be multiples of 8", C, ld);
  if (kd == 1 && dil == 1 && kh == kw && pad == kh / 2 && C == 32 && (kh == 5 || kh == 3)) {     // tiled 2-D kernel
    const __half* xh = HCP(x);
    if (kh == 5) return launch_conv_last_tiled<32, 5>(xh, N * D, H, W, ld, w, bias, out_scale, out_shift, affine_stats, out, range, ST(stream));
    return launch_conv_last_tiled<32, 3>(xh, N * D, H, W, ld, w, bias, out_scale, out_shift, affine_stats, out, range, ST(stream));
  }
  if (kd == 3 && kh == 3 && kw == 3 && dil == 1 && pad == 1 && (C == 32 || C == 64) && out_scale == 1.f && out_shift == 0.f) {
    if (C == 32) return launch_conv_last3d<32>(HCP(x), N, D, H, W, ld, w, bias, affine_stats, range, out, ST(stream));
    return launch_conv_last3d<64>(HCP(x), N, D, H, W, ld, w, bias, affine_stats, range, out, ST(stream));
  }
  const size_t total = (size_t)N * D * H * W;
  conv_last_kernel<<<tpz_div_up(total, 128), 128, 0, ST(stream)>>>(HCP(x), N, D, H, W, C, ld, w, bias, kd, kh, kw,
                                                                   dil, pad, out_scale, out_shift, affine_stats, out, range);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_conv_generic(const tpz_half* x0, int C0, int ld0, const tpz_half* x1, int C1, int ld1, int N, int D,
                                int H, int W, const float* w, const float* bias, int Co, int kd, int kh, int kw,
                                int stride, int dil, int pad, float neg_slope, const tpz_half* res, int res_ld,
                                int res_org, tpz_half* out, int out_ld, int Do, int Ho, int Wo, void* stream) {
  const size_t total = (size_t)N * Do * Ho * Wo * ((Co + 7) / 8);
  conv_generic_kernel<<<tpz_div_up(total, 128), 128, 0, ST(stream)>>>(
      HCP(x0), C0, ld0, HCP(x1), C1, ld1, N, D, H, W, w, bias, Co, kd, kh, kw, stride, dil, pad, neg_slope, HCP(res),
      res_ld, res_org, HP(out), out_ld, Do, Ho, Wo);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_maxpool2(const tpz_half* x, int N, int D, int H, int W, int C, int ld, int dims, tpz_half* out,
                            int out_ld, int lo_off, void* stream) {
  TPZ_CHECK(lo_off >= 0 && lo_off % 8 == 0, "tpz_maxpool2: lo_off=%d must be a non-negative multiple of 8", lo_off);
  TPZ_CHECK(C % 8 == 0 && ld % 8 == 0 && out_ld % 8 == 0, "tpz_maxpool2: channel counts must be multiples of 8");
  const int Do = dims == 3 ? D / 2 : D, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Do * Ho * Wo * (C / 8);
  if (total == 0) return 0;
  maxpool2_kernel<<<tpz_div_up(total, 256), 256, 0, ST(stream)>>>(HCP(x), N, D, H, W, C, ld, dims, HP(out), out_ld,
                                                                  Do, Ho, Wo, lo_off);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_upsample_nearest(const tpz_half* x, int N, int D, int H, int W, int C, int ld, int Do, int Ho,
                                    int Wo, tpz_half* out, int out_ld, int out_coff, void* stream) {
  TPZ_CHECK(C % 8 == 0 && ld % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0,
            "tpz_upsample_nearest: channel counts must be multiples of 8");
  const size_t total = (size_t)N * Do * Ho * Wo * (C / 8);
  upsample_kernel<<<tpz_div_up(total, 256), 256, 0, ST(stream)>>>(HCP(x), N, D, H, W, C, ld, Do, Ho, Wo, HP(out),
                                                                  out_ld, out_coff);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_meanstd(const float* x, long long n, int unbiased, float* stats, double* work, void* stream) {
  TPZ_CHECK(n > 1, "tpz_meanstd: need n > 1");
  int grid = tpz_div_up(n, 256 * 8);
  if (grid > 148 * 8) grid = 148 * 8;
  zero2_kernel<<<1, 32, 0, ST(stream)>>>(work);
  zero2_kernel<<<1, 32, 0, ST(stream)>>>(work + 2);
  sums_kernel<<<grid, 256, 0, ST(stream)>>>(x, n, work);
  var_kernel<<<grid, 256, 0, ST(stream)>>>(x, n, work);
  finalize_kernel<<<1, 1, 0, ST(stream)>>>(work, n, unbiased, stats);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_affine(const float* x, long long n, const float* stats, int inverse, float* y, void* stream) {
  int grid = tpz_div_up(n, 256 * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  affine_kernel<<<grid, 256, 0, ST(stream)>>>(x, n, stats, inverse, y);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_filter_f32(const float* x, int N, int D, int H, int W, const float* f, int kd, int kh, int kw, float bias,
                              float* y, void* stream) {
  TPZ_CHECK(kd % 2 == 1 && kh % 2 == 1 && kw % 2 == 1, "tpz_filter_f32: filter sizes must be odd");
  const size_t smem = (size_t)kd * kh * kw * sizeof(float);
  TPZ_CHECK(smem <= 200 * 1024, "tpz_filter_f32: filter too large");
  TPZ_CUDA(cudaFuncSetAttribute(filter_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t total = (size_t)N * D * H * W;
  filter_f32_kernel<<<tpz_div_up(total, 256), 256, smem, ST(stream)>>>(x, N, D, H, W, f, kd, kh, kw, bias, y);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_range_scale(const float* x, long long n, float* range, unsigned* work, void* stream) {
  TPZ_CHECK(n > 0, "tpz_range_scale: empty input");
  int grid = tpz_div_up(n, 256 * 8);
  if (grid > 148 * 8) grid = 148 * 8;
  absmax_kernel<<<grid, 256, 0, ST(stream)>>>(x, n, work);
  range_finalize_kernel<<<1, 1, 0, ST(stream)>>>(work, range);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tpz_f32_to_f16(const float* x, long long n, tpz_half* y, void* stream) {
  int grid = tpz_div_up(n, 256 * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  f32_to_f16_kernel<<<grid, 256, 0, ST(stream)>>>(x, n, HP(y));
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// tcgen05 implicit-GEMM convolution, halo-resident variant ("v2") for sm_100a.
//
// Same math and epilogue as tpz_tc_conv.cu (see there for the reference call sites), different data movement:
//   * poly-phase lattice: a conv with in-plane dilation L touches, for one output pixel, only input pixels
//     on the same residue class mod L.  A CTA tile is therefore 8 x 32 points of ONE lattice (pixels L apart),
//     fetched with TMA element strides (1, L, L); on that lattice every tap is a shift by ONE point.
//   * halo-resident A: per (source, channel chunk, z-tap) the (8+kw-1) x (32+kh-1) lattice halo is loaded
//     ONCE; each (r,s) tap is an MMA whose A descriptor starts at row r*HX+s of that tile (128B-swizzled
//     operands may start at any row; 8-row groups are HX rows apart -> SBO = HX*row_bytes).  Verified on
//     B200 by tools/bringup.py (lab): arbitrary start row / SBO with base_offset = 0.
//   * two M=128 accumulators per CTA (rows 0-15 / 16-31 of the tile) share every B (weight) block, halving
//     weight traffic; small weight sets stay resident in shared memory for the whole persistent kernel.
// Warp roles: warp0 = A (activation) TMA producer, warp3 = B (weight) TMA producer, warp1 = MMA issuer,
// warp2 = TMEM allocator, warps4-7 = epilogue.
#include "tpz_common.cuh"
#include "tpz_tc_conv.h"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int T2W = 8, T2H = 32;
constexpr int kMaxGroups = 64;
constexpr int kMaxAStages = 4, kMaxBStages = 8;

struct Tc2Group {
  int16_t src, c0, dz, ntaps;
  int16_t tap_begin, gox, goy, pad;   // (gox, goy): residue class of the group's taps modulo the source lattice
};
struct Tc2Tap {
  uint16_t kb, row_off;
};

struct Tc2Params {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nsrc;
  int org[2][3];
  int hx[2], rows_loaded[2], split_rows[2];
  int slat[2], sphase[2];   // per-source lattice spacing / whether the output phase shifts the source
  int slatz[2], Lz, phz, Dq; // z lattice: output plane = zq*Lz + phz (zq < Dq), source plane = zq*slatz + org_z + dz
  int L, phase_fixed;
  int N, Do, Ho, Wo, Co, CS;
  int tiles_x, tiles_y, num_tiles;
  int ngroups, nkb;
  int a_stages, b_stages, b_resident, acc_stages, ntile;
  int a_stage_bytes, b_block_bytes;
  const float* bias;
  float neg_slope;
  const __half* res;
  const float* res_scale;
  int res_ld, res_D, res_H, res_W, res_org[3];
  __half* out;
  int out_ld, out_coff;
  const float* dot_w;
  float dot_b;
  float* dot_out;
  const float* dot_affine;
  const float* oscale;
  const float* range;
  int out_lo;
  Tc2Group groups[kMaxGroups];
  Tc2Tap taps[TPZ_TC_MAX_KB];
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
// Bounded spin: a protocol error (wrong arrival count, lost TMA credit) traps after ~10 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins)
    if (spins > (1u << 28)) __trap();
}

struct TileCoord {
  int n, z, qx, qy, phx, phy;   // lattice index of the tile origin and the phase; output pixel = (q + i) * L + ph
};
__device__ __forceinline__ TileCoord decode_tile(const Tc2Params& p, int tile) {
  const int LL = p.L * p.L;
  int ph, rest;
  if (p.phase_fixed >= 0) { ph = p.phase_fixed; rest = tile; }
  else { ph = tile % LL; rest = tile / LL; }
  const int txq = rest % p.tiles_x;
  rest /= p.tiles_x;
  const int tyq = rest % p.tiles_y;
  const int plane = rest / p.tiles_y;
  TileCoord t;
  t.n = plane / p.Dq;
  t.z = plane - t.n * p.Dq;      // z lattice index
  t.qx = txq * T2W;
  t.qy = tyq * (16 * p.ntile);
  t.phx = ph % p.L;
  t.phy = ph / p.L;
  return t;
}

// PAIR = true: the kernel is launched in clusters of two CTAs (one TPC) that issue M = 256 MMAs together
// (tcgen05 cta_group::2).  CTA `rank` works on tile 2T + rank of super-tile T with its own halo tiles and accumulators
// and holds HALF of every weight block (rows rank*Co/2 ...), so the shared-memory operand traffic per MMA drops from
// A + B to A + B/2 per SM -- the N <= 128 layers are bound by exactly that traffic.  The even CTA (leader) issues all
// MMAs; TMA loads of both CTAs credit the leader's "full" barriers, tcgen05.commit multicasts the "empty" / "accumulator
// full" arrivals to both CTAs, and both epilogues release an accumulator stage on the leader's barrier.
template <int KC, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) tc_conv2_kernel(const __grid_constant__ Tc2Params p) {
  constexpr int ROWB = KC * 2;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cta_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // first super-tile of this CTA (pair)
  const int cta_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int nsuper = PAIR ? (p.num_tiles + 1) >> 1 : p.num_tiles;
  const int last_tile = p.num_tiles - 1;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int AS = p.a_stages, BS = p.b_stages;
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)AS * p.a_stage_bytes;
  const int b_region = p.b_resident ? p.nkb * p.b_block_bytes : BS * p.b_block_bytes;
  uint8_t* tail = smem + (size_t)AS * p.a_stage_bytes + b_region;
  uint64_t* afull = reinterpret_cast<uint64_t*>(tail);
  uint64_t* aempty = afull + kMaxAStages;
  uint64_t* bfull = aempty + kMaxAStages;
  uint64_t* bempty = bfull + kMaxBStages;
  uint64_t* bres = bempty + kMaxBStages;
  uint64_t* tfull = bres + 1;
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
    for (int s = 0; s < kMaxAStages; ++s) { ptx::mbar_init(&afull[s], 1); ptx::mbar_init(&aempty[s], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { ptx::mbar_init(&bfull[s], 1); ptx::mbar_init(&bempty[s], 1); }
    ptx::mbar_init(bres, 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], PAIR ? 8 : 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) ptx::tmem_alloc_pair<kTmemCols>(tmem_slot); else ptx::tmem_alloc<kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();        // the peer's barriers are initialised before any remote arrive / TMA credit
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ A producer: one halo tile per (tile, group) ------------------------------
    // warp-uniform loop; the TMA instructions are issued by one elected lane (keeps operands in uniform registers)
    {
      uint32_t s = 0, ph = 0;
      for (int T = cta_first; T < nsuper; T += cta_step) {
        const int tile = PAIR ? min(2 * T + (int)rank, last_tile) : T;      // an odd tile count leaves the peer a repeat
        const TileCoord tc = decode_tile(p, tile);
        for (int g = 0; g < p.ngroups; ++g) {
          mbar_wait(&aempty[s], ph ^ 1);
          const Tc2Group G = p.groups[g];
          const int src = G.src;
          const int hx = p.hx[src];
          const int sr = p.split_rows[src];
          uint8_t* dst = smem + (size_t)s * p.a_stage_bytes;
          if (ptx::elect_one()) {
            const uint32_t bytes = (uint32_t)(hx * p.rows_loaded[src] * ROWB);
            if (!PAIR) ptx::mbar_expect_tx(&afull[s], bytes);
            else if (leader) ptx::mbar_expect_tx(&afull[s], 2 * bytes);
            for (int row0 = 0; row0 < p.rows_loaded[src]; row0 += sr) {
              const int cx = tc.qx * p.slat[src] + p.sphase[src] * tc.phx + p.org[src][0] + G.gox;
              const int cy = (tc.qy + row0) * p.slat[src] + p.sphase[src] * tc.phy + p.org[src][1] + G.goy;
              const int cz = tc.z * p.slatz[src] + p.sphase[src] * p.phz + p.org[src][2] + G.dz;
              if (PAIR) ptx::tma_load_5d_pair(dst + (size_t)row0 * hx * ROWB, &p.tmA[src], &afull[s], G.c0, cx, cy, cz, tc.n);
              else ptx::tma_load_5d(dst + (size_t)row0 * hx * ROWB, &p.tmA[src], &afull[s], G.c0, cx, cy, cz, tc.n);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)AS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------ B producer: weights, resident or streamed per tap ------------------------------
    if (p.b_resident) {
      if (ptx::elect_one()) {
        const uint32_t bytes = (uint32_t)(p.nkb * p.b_block_bytes);
        if (!PAIR) ptx::mbar_expect_tx(bres, bytes);
        else if (leader) ptx::mbar_expect_tx(bres, 2 * bytes);
        for (int kb = 0; kb < p.nkb; ++kb) {
          uint8_t* dst = smem + (size_t)AS * p.a_stage_bytes + (size_t)kb * p.b_block_bytes;
          if (PAIR) ptx::tma_load_2d_pair(dst, &p.tmB, bres, 0, kb * p.Co + (int)rank * (p.Co >> 1));
          else ptx::tma_load_2d(dst, &p.tmB, bres, 0, kb * p.Co);
        }
      }
    } else {
      uint32_t s = 0, ph = 0;
      for (int T = cta_first; T < nsuper; T += cta_step) {
        for (int g = 0; g < p.ngroups; ++g) {
          const Tc2Group G = p.groups[g];
          for (int t = 0; t < G.ntaps; ++t) {
            mbar_wait(&bempty[s], ph ^ 1);
            const int kb = (int)p.taps[G.tap_begin + t].kb;
            if (ptx::elect_one()) {
              uint8_t* dst = smem + (size_t)AS * p.a_stage_bytes + (size_t)s * p.b_block_bytes;
              if (!PAIR) {
                ptx::mbar_expect_tx(&bfull[s], (uint32_t)p.b_block_bytes);
                ptx::tma_load_2d(dst, &p.tmB, &bfull[s], 0, kb * p.Co);
              } else {
                if (leader) ptx::mbar_expect_tx(&bfull[s], 2u * (uint32_t)p.b_block_bytes);
                ptx::tma_load_2d_pair(dst, &p.tmB, &bfull[s], 0, kb * p.Co + (int)rank * (p.Co >> 1));
              }
            }
            __syncwarp();
            if (++s == (uint32_t)BS) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------ MMA issuer (pair mode: the leader CTA only) ------------------------------
    // The whole warp walks the (warp-uniform) loops so addresses/descriptors live in uniform registers; lane 0
    // issues the tcgen05 instructions.  Measured on B200 (tools/layer_bench.py): back-to-back MMAs into ONE
    // accumulator are latency-chained (N=64: 93 cycles each), alternating two accumulators reaches 48 (N=64) /
    // 64 (N=128) / 128 (N=256) cycles, so the issue loop must stay well below ~40 instructions per MMA pair.
    const uint32_t idesc = PAIR ? ptx::umma_idesc_f16_pair(p.Co) : ptx::umma_idesc_f16(128, p.Co);
    const uint32_t b_hi = ptx::umma_desc_hi(8 * ROWB, LAYOUT);
    if (p.b_resident) {
      mbar_wait(bres, 0);
      ptx::tc_fence_after();
    }
    const bool two = p.ntile == 2;
    const bool resident = p.b_resident != 0;
    uint32_t as = 0, aph = 0, bs = 0, bph = 0, st = 0, tph = 0;
    for (int T = cta_first; T < nsuper; T += cta_step) {
      mbar_wait(&tempty[st], tph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + (st * p.ntile) * p.CS;
      const uint32_t d1 = d0 + p.CS;
      uint32_t accf = 0;   // first MMA of the tile overwrites the accumulators
      for (int g = 0; g < p.ngroups; ++g) {
        mbar_wait(&afull[as], aph);
        ptx::tc_fence_after();
        const Tc2Group G = p.groups[g];
        const uint32_t hxb = (uint32_t)p.hx[G.src] * ROWB;          // bytes between consecutive tile rows (y)
        const uint32_t a_hi = ptx::umma_desc_hi(hxb, LAYOUT);
        const uint32_t a_tile_lo = ((a_base + as * (uint32_t)p.a_stage_bytes) & 0x3FFFFu) >> 4;
        const uint32_t a1_off_lo = (16u * hxb) >> 4;
        for (int t = 0; t < G.ntaps; ++t) {
          const Tc2Tap T = p.taps[G.tap_begin + t];
          uint32_t b_lo;
          if (resident) {
            b_lo = ((b_base + (uint32_t)T.kb * p.b_block_bytes) & 0x3FFFFu) >> 4;
          } else {
            mbar_wait(&bfull[bs], bph);
            ptx::tc_fence_after();
            b_lo = ((b_base + bs * (uint32_t)p.b_block_bytes) & 0x3FFFFu) >> 4;
          }
          const uint32_t a0_lo = a_tile_lo + (((uint32_t)T.row_off * ROWB) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              if (PAIR) {
                ptx::umma_f16_pair(d0, a0_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (k == 0) ? accf : 1u);
                if (two) ptx::umma_f16_pair(d1, a0_lo + a1_off_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (k == 0) ? accf : 1u);
              } else {
                ptx::umma_f16_lohi(d0, a0_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (k == 0) ? accf : 1u);
                if (two) ptx::umma_f16_lohi(d1, a0_lo + a1_off_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (k == 0) ? accf : 1u);
              }
            }
            if (!resident) { if (PAIR) ptx::umma_commit_pair(&bempty[bs]); else ptx::umma_commit(&bempty[bs]); }
          }
          accf = 1;
          if (!resident) {
            if (++bs == (uint32_t)BS) { bs = 0; bph ^= 1; }
          }
        }
        if (ptx::elect_one()) { if (PAIR) ptx::umma_commit_pair(&aempty[as]); else ptx::umma_commit(&aempty[as]); }
        if (++as == (uint32_t)AS) { as = 0; aph ^= 1; }
      }
      if (ptx::elect_one()) { if (PAIR) ptx::umma_commit_pair(&tfull[st]); else ptx::umma_commit(&tfull[st]); }
      if (p.acc_stages == 2) { st ^= 1; if (st == 0) tph ^= 1; } else { tph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int ew = warp - 4;
    const int m = ew * 32 + lane;
    const int li = m & 7, lj = m >> 3;   // lattice point (li, lj) of accumulator 0; accumulator 1 is 16 rows lower
    uint32_t tcount = 0;
    for (int T = cta_first; T < nsuper; T += cta_step, ++tcount) {
      const int tile_raw = PAIR ? 2 * T + (int)rank : T;
      const bool tile_ok = tile_raw <= last_tile;                       // pair mode, odd tile count: nothing to write
      const TileCoord tc = decode_tile(p, tile_ok ? tile_raw : last_tile);
      const uint32_t st = (p.acc_stages == 2) ? (tcount & 1) : 0;
      const uint32_t aph = (p.acc_stages == 2) ? ((tcount >> 1) & 1) : (tcount & 1);
      mbar_wait(&tfull[st], aph);
      ptx::tc_fence_after();
      for (int a = 0; a < p.ntile; ++a) {
        const int gx = (tc.qx + li) * p.L + tc.phx;
        const int gy = (tc.qy + lj + 16 * a) * p.L + tc.phy;
        const bool valid = tile_ok && (gx < p.Wo) && (gy < p.Ho);
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (st * p.ntile + a) * p.CS;
        const int gz = tc.z * p.Lz + p.phz;
        const long long opix = (((long long)tc.n * p.Do + gz) * p.Ho + gy) * p.Wo + gx;
        __half* orow = p.out ? p.out + opix * p.out_ld + p.out_coff : nullptr;
        const __half* rrow = nullptr;
        if (p.res) {
          const long long rpix =
              (((long long)tc.n * p.res_D + (gz + p.res_org[2])) * p.res_H + (gy + p.res_org[1])) * p.res_W +
              (gx + p.res_org[0]);
          rrow = p.res + rpix * p.res_ld;
        }
        float dot = 0.f;
        for (int c = 0; c < p.Co; c += 32) {
          uint32_t r[32];
          if (p.Co - c >= 32) {
            ptx::tmem_ld32(taddr + c, r);
          } else {
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
            const bool wide_r = ((reinterpret_cast<uintptr_t>(rrow + c) & 31) == 0);     // full-sector 32-byte loads
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q * 16 < nc) {
                uint32_t u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (wide_r && q * 16 + 8 < nc) {
                  ptx::ld_global_256(rrow + c + q * 16, u);
                } else {
                  const uint4 lo = *reinterpret_cast<const uint4*>(rrow + c + q * 16);
                  u[0] = lo.x; u[1] = lo.y; u[2] = lo.z; u[3] = lo.w;
                  if (q * 16 + 8 < nc) {
                    const uint4 hi = *reinterpret_cast<const uint4*>(rrow + c + q * 16 + 8);
                    u[4] = hi.x; u[5] = hi.y; u[6] = hi.z; u[7] = hi.w;
                  }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[e]));
                  const int j = q * 16 + e * 2;
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
            // 32-byte stores (one full L2 sector each) when the channel slice is 32-byte aligned, else 16-byte pieces
            const bool wide = ((reinterpret_cast<uintptr_t>(orow + c) & 31) == 0);
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
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) ptx::mbar_arrive_leader(&tempty[st]); else ptx::mbar_arrive(&tempty[st]); }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();        // no CTA leaves while its peer may still signal its barriers / write its TMEM
  if (warp == 2) {
    if (PAIR) ptx::tmem_dealloc_pair<kTmemCols>(tmem_base); else ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

int g_num_sms2 = 0;

}  // namespace

// returns 0 ok, <0 "not eligible" (caller falls back to the per-tap kernel), >0 error
static int launch_v2(const TpzTcConvArgs* a, cudaStream_t stream, bool dry, bool allow_pair = true) {
  const int L = a->lattice;
  if (L < 1 || L > 8) return -1;
  const int rowb = a->KC * 2;
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.L = L;
  int cs = 32;
  while (cs < a->Co) cs <<= 1;
  p.CS = cs;
  // Co = 256: two M=128 tiles fill all 512 TMEM columns (single-buffered accumulators, B shared by both tiles);
  // TPZ_CO256_NTILE=1 selects one tile per CTA with double-buffered accumulators instead (measured slower on B200:
  // twice the weight traffic per pixel starves the B ring).
  static const int co256_ntile = getenv("TPZ_CO256_NTILE") ? atoi(getenv("TPZ_CO256_NTILE")) : 2;
  p.ntile = (4 * cs <= kTmemCols) ? 2 : (co256_ntile == 1 ? 1 : 2);
  p.acc_stages = (2 * p.ntile * cs <= kTmemCols) ? 2 : 1;
  const int th = 16 * p.ntile;
  int a_stage = 0;
  // lattice tap grid per source, derived from the k-block table: tap (dx,dy) = residue class (dx mod sl, dy mod sl)
  // + lattice index (dx div sl, dy div sl); offsets must be >= 0 (the source origin absorbs negative taps)
  int kwq[2] = {1, 1}, khq[2] = {1, 1};
  for (int s = 0; s < a->nsrc; ++s) {
    const int sl = a->src[s].lat > 0 ? a->src[s].lat : L;
    if (sl > 8) return -1;
    p.slat[s] = sl; p.sphase[s] = a->src[s].no_phase ? 0 : 1;
  }
  for (int i = 0; i < a->nkb; ++i) {
    const TcKBlock& k = a->kb[i];
    if (k.dx < 0 || k.dy < 0) return -1;
    const int sl = p.slat[k.src];
    if (k.dx / sl + 1 > kwq[k.src]) kwq[k.src] = k.dx / sl + 1;
    if (k.dy / sl + 1 > khq[k.src]) khq[k.src] = k.dy / sl + 1;
  }
  for (int s = 0; s < a->nsrc; ++s) {
    const int sl = p.slat[s];
    const int hx = T2W + kwq[s] - 1, hy = th + khq[s] - 1;
    if ((hx - 1) * sl + 1 > 256) return -1;
    const int ext = (hy - 1) * sl + 1;
    const int nsplit = (ext + 255) / 256;
    const int sr = (hy + nsplit - 1) / nsplit;
    if ((sr - 1) * sl + 1 > 256) return -1;
    p.hx[s] = hx; p.split_rows[s] = sr; p.rows_loaded[s] = sr * nsplit;
    const int bytes = (hx * sr * nsplit * rowb + 1023) / 1024 * 1024;
    if (bytes > a_stage) a_stage = bytes;
  }
  // group the k-blocks by (source, chunk, z-tap, residue class); all taps of a group read one halo tile
  int ng = 0, nt = 0;
  bool used[TPZ_TC_MAX_KB];
  memset(used, 0, sizeof(used));
  for (int i = 0; i < a->nkb; ++i) {
    if (used[i]) continue;
    if (ng >= kMaxGroups) return -1;
    Tc2Group& G = p.groups[ng];
    const int sl0 = p.slat[a->kb[i].src];
    G.src = (int16_t)a->kb[i].src; G.c0 = a->kb[i].c0; G.dz = a->kb[i].dz; G.tap_begin = (int16_t)nt; G.ntaps = 0;
    G.gox = (int16_t)(a->kb[i].dx % sl0); G.goy = (int16_t)(a->kb[i].dy % sl0); G.pad = 0;
    for (int j = i; j < a->nkb; ++j) {
      const TcKBlock& k = a->kb[j];
      if (used[j] || k.src != G.src || k.c0 != G.c0 || k.dz != G.dz || k.dx % sl0 != G.gox || k.dy % sl0 != G.goy) continue;
      p.taps[nt].kb = (uint16_t)j;
      p.taps[nt].row_off = (uint16_t)((k.dy / sl0) * p.hx[k.src] + (k.dx / sl0));
      used[j] = true; ++nt; ++G.ntaps;
    }
    ++ng;
  }
  p.ngroups = ng; p.nkb = a->nkb;
  // CTA-pair mode (tcgen05 cta_group::2): each CTA of a cluster of two keeps half of every weight block.
  // TPZ_TC_PAIR = 0 never; default: whenever eligible (measured on B200, 2048^2 layers: 64->64 +10%, 64->128 +11%,
  // 128->128 +18%, 128->256 +4%).
  static const int pair_env = getenv("TPZ_TC_PAIR") ? atoi(getenv("TPZ_TC_PAIR")) : -1;
  long long ntl_early;
  {
    const int lz = a->lattice_z > 1 ? a->lattice_z : 1;
    const long long dq = (a->Do - a->phase_z + lz - 1) / lz;
    ntl_early = (long long)(a->phase_sel > 0 ? 1 : L * L) * tpz_div_up(tpz_div_up(a->Wo, L), T2W) *
                tpz_div_up(tpz_div_up(a->Ho, L), th) * (dq > 0 ? dq : 0) * a->N;
  }
  const bool pair_ok = a->Co % 32 == 0 && a->Co >= 32 && ntl_early >= 2;
  const bool pair = allow_pair && pair_ok && pair_env != 0;
  p.b_block_bytes = (pair ? a->Co / 2 : a->Co) * rowb;
  const int tail = 4096;
  const int budget = 227 * 1024 - 1024 - tail;
  const int resident_bytes = a->nkb * p.b_block_bytes;
  p.a_stages = 2;
  if (p.a_stages * a_stage + 2 * p.b_block_bytes > budget) return -1;
  if (resident_bytes <= 112 * 1024 && p.a_stages * a_stage + resident_bytes <= budget && resident_bytes < (1 << 20)) {
    p.b_resident = 1; p.b_stages = 0;
    int as = (budget - resident_bytes) / a_stage;
    p.a_stages = as > kMaxAStages ? kMaxAStages : as;
  } else {
    p.b_resident = 0;
    int as = 2;
    if (3 * a_stage + 4 * p.b_block_bytes <= budget) as = 3;
    p.a_stages = as;
    int bs = (budget - as * a_stage) / p.b_block_bytes;
    p.b_stages = bs > kMaxBStages ? kMaxBStages : bs;
    if (p.b_stages < 2) return -1;
  }
  p.a_stage_bytes = a_stage;
  if (dry) return 0;

  for (int s = 0; s < a->nsrc; ++s) {
    const TpzTcSrc& src = a->src[s];
    TPZ_CHECK(src.C % a->KC == 0, "tpz_tc_conv: source %d channels %d not a multiple of KC=%d", s, src.C, a->KC);
    TPZ_CHECK(src.ld % 8 == 0, "tpz_tc_conv: source %d channel stride %d must be a multiple of 8", s, src.ld);
    uint64_t dims[5] = {(uint64_t)src.C, (uint64_t)src.W, (uint64_t)src.H, (uint64_t)src.D, (uint64_t)src.N};
    uint64_t strides[4] = {(uint64_t)src.ld * 2, (uint64_t)src.ld * 2 * src.W, (uint64_t)src.ld * 2 * src.W * src.H,
                           (uint64_t)src.ld * 2 * src.W * src.H * src.D};
    const uint32_t sl = (uint32_t)p.slat[s];
    uint32_t box[5] = {(uint32_t)a->KC, (uint32_t)((p.hx[s] - 1) * sl + 1), (uint32_t)((p.split_rows[s] - 1) * sl + 1), 1, 1};
    uint32_t es[5] = {1, sl, sl, 1, 1};
    int rc = tpz_encode_tmap(&p.tmA[s], src.ptr, 5, dims, strides, box, es, rowb);
    if (rc) return rc;
    p.org[s][0] = src.org[0]; p.org[s][1] = src.org[1]; p.org[s][2] = src.org[2];
  }
  {
    uint64_t dims[2] = {(uint64_t)a->KC, (uint64_t)a->nkb * a->Co};
    uint64_t strides[1] = {(uint64_t)rowb};
    uint32_t box[2] = {(uint32_t)a->KC, (uint32_t)(pair ? a->Co / 2 : a->Co)};
    uint32_t es[2] = {1, 1};
    int rc = tpz_encode_tmap(&p.tmB, a->weights, 2, dims, strides, box, es, rowb);
    if (rc) return rc;
  }
  p.nsrc = a->nsrc;
  p.N = a->N; p.Do = a->Do; p.Ho = a->Ho; p.Wo = a->Wo; p.Co = a->Co;
  const int qW = tpz_div_up(a->Wo, L), qH = tpz_div_up(a->Ho, L);
  p.tiles_x = tpz_div_up(qW, T2W);
  p.tiles_y = tpz_div_up(qH, th);
  p.phase_fixed = a->phase_sel > 0 ? a->phase_sel - 1 : -1;
  TPZ_CHECK(a->phase_sel >= 0 && a->phase_sel <= L * L, "tpz_tc_conv: phase_sel=%d out of range for lattice %d", a->phase_sel, L);
  p.Lz = a->lattice_z > 1 ? a->lattice_z : 1;
  p.phz = a->phase_z;
  TPZ_CHECK(p.phz >= 0 && p.phz < p.Lz, "tpz_tc_conv: phase_z=%d out of range for lattice_z=%d", p.phz, p.Lz);
  p.Dq = (a->Do - p.phz + p.Lz - 1) / p.Lz;
  for (int s = 0; s < a->nsrc; ++s) p.slatz[s] = a->src[s].lat_z > 0 ? a->src[s].lat_z : p.Lz;
  if (p.Dq <= 0) return 0;
  const long long ntl = (long long)(p.phase_fixed >= 0 ? 1 : L * L) * p.tiles_x * p.tiles_y * p.Dq * a->N;
  TPZ_CHECK(ntl > 0 && ntl < (1ll << 31), "tpz_tc_conv: bad tile count %lld", ntl);
  p.num_tiles = (int)ntl;
  p.bias = a->bias; p.neg_slope = a->neg_slope;
  p.res = reinterpret_cast<const __half*>(a->res); p.res_scale = a->res_scale; p.res_ld = a->res_ld;
  p.res_D = a->res_D; p.res_H = a->res_H; p.res_W = a->res_W;
  p.res_org[0] = a->res_org[0]; p.res_org[1] = a->res_org[1]; p.res_org[2] = a->res_org[2];
  p.out = reinterpret_cast<__half*>(a->out); p.out_ld = a->out_ld; p.out_coff = a->out_coff;
  p.dot_w = a->dot_w; p.dot_b = a->dot_b; p.dot_out = a->dot_out; p.dot_affine = a->dot_affine;
  p.oscale = a->oscale; p.range = a->range; p.out_lo = a->out_lo;
  TPZ_CHECK(a->out_lo % 8 == 0 && a->out_lo >= 0, "tpz_tc_conv: out_lo=%d must be a non-negative multiple of 8", a->out_lo);

  if (g_num_sms2 == 0) {
    int dev = 0;
    TPZ_CUDA(cudaGetDevice(&dev));
    TPZ_CUDA(cudaDeviceGetAttribute(&g_num_sms2, cudaDevAttrMultiProcessorCount, dev));
  }
  const int b_region = p.b_resident ? resident_bytes : p.b_stages * p.b_block_bytes;
  int smem = p.a_stages * a_stage + b_region + tail + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;  // 1 CTA / SM (every CTA owns all 512 TMEM columns)
  if (pair) {
    // clusters of two CTAs; an odd tile count leaves the second CTA of the last pair a repeat tile it does not write
    int grid = 2 * ((p.num_tiles + 1) / 2);
    const int cap = g_num_sms2 & ~1;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t le;
    if (a->KC == 64) {
      TPZ_CUDA(cudaFuncSetAttribute(tc_conv2_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      le = cudaLaunchKernelEx(&cfg, tc_conv2_kernel<64, true>, p);
    } else {
      TPZ_CUDA(cudaFuncSetAttribute(tc_conv2_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      le = cudaLaunchKernelEx(&cfg, tc_conv2_kernel<32, true>, p);
    }
    if (le != cudaSuccess) {          // clusters of two not schedulable here (e.g. a partitioned GPU): single-CTA plan
      (void)cudaGetLastError();
      return launch_v2(a, stream, dry, false);
    }
    TPZ_CUDA(cudaGetLastError());
    return 0;
  }
  const int grid = p.num_tiles < g_num_sms2 ? p.num_tiles : g_num_sms2;
  if (a->KC == 64) {
    TPZ_CUDA(cudaFuncSetAttribute(tc_conv2_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_conv2_kernel<64, false><<<grid, kThreads, smem, stream>>>(p);
  } else {
    TPZ_CUDA(cudaFuncSetAttribute(tc_conv2_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_conv2_kernel<32, false><<<grid, kThreads, smem, stream>>>(p);
  }
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

static int check_common(const TpzTcConvArgs* a) {
  TPZ_CHECK(a != nullptr, "tpz_tc_conv: null args");
  TPZ_CHECK(a->KC == 64 || a->KC == 32, "tpz_tc_conv: KC must be 32 or 64 (got %d)", a->KC);
  TPZ_CHECK(a->Co >= 16 && a->Co <= 256 && a->Co % 16 == 0, "tpz_tc_conv: Co=%d must be a multiple of 16 in [16,256]", a->Co);
  TPZ_CHECK(a->nsrc >= 1 && a->nsrc <= 2, "tpz_tc_conv: nsrc=%d", a->nsrc);
  TPZ_CHECK(a->nkb >= 1 && a->nkb <= TPZ_TC_MAX_KB, "tpz_tc_conv: nkb=%d exceeds %d", a->nkb, TPZ_TC_MAX_KB);
  TPZ_CHECK(a->out != nullptr || a->dot_out != nullptr, "tpz_tc_conv: no output");
  TPZ_CHECK(a->out == nullptr || (a->out_ld % 8 == 0 && a->out_coff % 8 == 0), "tpz_tc_conv: output channel stride/offset must be multiples of 8");
  TPZ_CHECK(a->res == nullptr || a->res_ld % 8 == 0, "tpz_tc_conv: residual channel stride must be a multiple of 8");
  for (int i = 0; i < a->nkb; ++i)
    TPZ_CHECK(a->kb[i].src >= 0 && a->kb[i].src < a->nsrc, "tpz_tc_conv: k-block %d bad source", i);
  return 0;
}

extern "C" int tpz_tc_conv_v2(const TpzTcConvArgs* a, void* stream) {
  int rc = check_common(a);
  if (rc) return rc;
  rc = launch_v2(a, reinterpret_cast<cudaStream_t>(stream), false);
  if (rc < 0) return tpz_fail(3, "tpz_tc_conv_v2: configuration not eligible for the halo-resident kernel");
  return rc;
}

extern "C" int tpz_tc_conv(const TpzTcConvArgs* a, void* stream) {
  int rc = check_common(a);
  if (rc) return rc;
  static const bool force_v1 = getenv("TPZ_TC_FORCE_V1") != nullptr;
  if (!force_v1 && launch_v2(a, reinterpret_cast<cudaStream_t>(stream), true) == 0)
    return launch_v2(a, reinterpret_cast<cudaStream_t>(stream), false);
  return tpz_tc_conv_v1(a, stream);
}

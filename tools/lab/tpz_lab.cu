// Hardware probe ("lab") for UMMA shared-memory descriptor behaviour on sm_100a.  Not on the product path:
// it answers one design question for the halo-reuse variant of the conv kernel — can an A operand start at
// an arbitrary ROW of a TMA-written 128B-swizzled tile (start address not 1024-B aligned), and with what
// base_offset / SBO?  One CTA: TMA-load A [rows][64] and B [N][64] fp16 (SWIZZLE_128B), issue 4 MMAs
// (K = 64) with A start = row `shift`, 8-row-group stride `sbo_rows`, write D [128][N] fp32.
#include "../../topaz_b200/csrc/tpz_common.cuh"
#include "tpz_lab.h"

namespace {
__device__ __forceinline__ bool lab_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128, 1) lab_kernel(const __grid_constant__ CUtensorMap tmA,
                                                     const __grid_constant__ CUtensorMap tmB, int rowsA, int N,
                                                     int shift, int sbo_rows, int base_off_mode, int rowb, float* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_bytes = rowsA * rowb, b_bytes = N * rowb;
  const uint32_t layout = rowb == 128 ? 2u : 4u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a_bytes + b_bytes);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar, (uint32_t)(a_bytes + b_bytes));
    for (int r = 0; r < rowsA; r += 256) ptx::tma_load_2d(smem + r * rowb, &tmA, bar, 0, r);
    ptx::tma_load_2d(smem + a_bytes, &tmB, bar, 0, 0);
    while (!lab_try(bar, 0)) {}
    ptx::tc_fence_after();
    const uint32_t a_addr = base + shift * rowb;
    const uint32_t b_addr = base + a_bytes;
    const uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7) : 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    for (int k = 0; k < rowb / 32; ++k) {
      const uint64_t da = ptx::umma_desc(a_addr + k * 32, sbo_rows * rowb, layout, bo);
      const uint64_t db = ptx::umma_desc(b_addr + k * 32, 8 * rowb, layout, 0);
      ptx::umma_f16(tmem, da, db, idesc, k != 0);
    }
    ptx::umma_commit(done);
  }
  __syncwarp();
  while (!lab_try(done, 0)) {}
  ptx::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<256>(tmem);
}
}  // namespace

// A: device fp16 [rowsA][kc], B: device fp16 [N][kc], D: device fp32 [128][N]; kc = 64 (SW128) or 32 (SW64)
extern "C" int tpz_lab_umma(const tpz_half* A, int rowsA, const tpz_half* B, int N, int shift, int sbo_rows,
                            int base_off_mode, int kc, float* D, void* stream) {
  TPZ_CHECK(kc == 64 || kc == 32, "tpz_lab_umma: kc must be 32 or 64");
  const int rowb = kc * 2;
  TPZ_CHECK(rowsA % 8 == 0 && rowsA <= 1024 && N % 16 == 0 && N <= 256, "tpz_lab_umma: bad sizes");
  CUtensorMap tmA, tmB;
  uint64_t dA[2] = {(uint64_t)kc, (uint64_t)rowsA}, sA[1] = {(uint64_t)rowb};
  uint32_t bA[2] = {(uint32_t)kc, (uint32_t)(rowsA < 256 ? rowsA : 256)}, es[2] = {1, 1};
  TPZ_CHECK(rowsA <= 256 || rowsA % 256 == 0, "tpz_lab_umma: rowsA > 256 must be a multiple of 256");
  int rc = tpz_encode_tmap(&tmA, A, 2, dA, sA, bA, es, rowb);
  if (rc) return rc;
  uint64_t dB[2] = {(uint64_t)kc, (uint64_t)N};
  uint32_t bB[2] = {(uint32_t)kc, (uint32_t)N};
  rc = tpz_encode_tmap(&tmB, B, 2, dB, sA, bB, es, rowb);
  if (rc) return rc;
  const int smem = rowsA * rowb + N * rowb + 1024 + 256;
  TPZ_CUDA(cudaFuncSetAttribute(lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lab_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, rowsA, N, shift, sbo_rows,
                                                                       base_off_mode, rowb, D);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// ---- probe 2: TMA tiled load with an element (traversal) stride: which rows land in shared memory? ----
namespace {
__global__ void lab_tma_stride_kernel(const __grid_constant__ CUtensorMap tm, int start, int nrows, __half* out) {
  extern __shared__ uint8_t smem_raw2[];
  const uint32_t raw = ptx::smem_u32(smem_raw2);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw2 + (base - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + nrows * 128);
  for (int i = threadIdx.x; i < nrows * 64; i += blockDim.x) reinterpret_cast<__half*>(smem)[i] = __float2half(-777.f);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar, (uint32_t)(nrows * 128));
    ptx::tma_load_2d(smem, &tm, bar, 0, start);
  }
  __syncthreads();
  int spins = 0;
  while (!lab_try(bar, 0) && ++spins < (1 << 22)) {}
  for (int i = threadIdx.x; i < nrows * 64; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}
}  // namespace

// A: device fp16 [rowsA][64].  Loads a box of `nrows` rows starting at row `start` with element stride `stride`
// along the row dimension (box extent (nrows-1)*stride+1) and copies the raw (swizzled) smem image to out.
extern "C" int tpz_lab_tma_stride(const tpz_half* A, int rowsA, int start, int stride, int nrows, tpz_half* out,
                                  void* stream) {
  CUtensorMap tm;
  uint64_t d[2] = {64, (uint64_t)rowsA}, s[1] = {128};
  uint32_t b[2] = {64, (uint32_t)((nrows - 1) * stride + 1)}, es[2] = {1, (uint32_t)stride};
  TPZ_CHECK(b[1] <= 256, "tpz_lab_tma_stride: box too large");
  int rc = tpz_encode_tmap(&tm, A, 2, d, s, b, es, 128);
  if (rc) return rc;
  const int smem = nrows * 128 + 2048;
  TPZ_CUDA(cudaFuncSetAttribute(lab_tma_stride_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lab_tma_stride_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, start, nrows,
                                                                                  reinterpret_cast<__half*>(out));
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

// ---- probe 3: MMA issue rate vs A-operand alignment.  One CTA issues `iters` x (K=64) MMA groups back to back
// on fixed smem operands (A start row = shift, 8-row-group stride = sbo_rows) and reports SM cycles per MMA. ----
namespace {
__global__ void __launch_bounds__(128, 1) lab_rate_kernel(int N, int shift, int sbo_rows, int iters, int two_acc,
                                                          long long* cycles) {
  extern __shared__ uint8_t smem_raw3[];
  const uint32_t raw = ptx::smem_u32(smem_raw3);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw3 + (base - raw);
  const int a_bytes = 512 * 128, b_bytes = N * 128;
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + a_bytes + b_bytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  for (int i = threadIdx.x; i < (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(done, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = base + shift * 128, b_addr = base + a_bytes;
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = ptx::umma_desc(b_addr + k * 32, 1024, 2, 0);
        ptx::umma_f16(tmem, ptx::umma_desc(a_addr + k * 32, sbo_rows * 128, 2, 0), db, idesc, 1);
        if (two_acc)
          ptx::umma_f16(tmem + 256, ptx::umma_desc(a_addr + 16 * sbo_rows * 128 + k * 32, sbo_rows * 128, 2, 0), db, idesc, 1);
      }
    }
    ptx::umma_commit(done);
    while (!lab_try(done, 0)) {}
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}
}  // namespace

extern "C" int tpz_lab_umma_rate(int N, int shift, int sbo_rows, int iters, int two_acc, long long* cycles, void* stream) {
  TPZ_CHECK(N % 16 == 0 && N >= 16 && N <= 256, "tpz_lab_umma_rate: bad N");
  TPZ_CHECK(shift + 31 * sbo_rows + 8 <= 512, "tpz_lab_umma_rate: tile out of range");
  const int smem = 512 * 128 + N * 128 + 2048;
  TPZ_CUDA(cudaFuncSetAttribute(lab_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lab_rate_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(N, shift, sbo_rows, iters, two_acc, cycles);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}


// ---- probe 4: CTA-pair MMA (tcgen05.mma.cta_group::2, M = 256 across two CTAs of a cluster).  Each CTA holds rows
// [128*rank, 128*rank+128) of A and rows [N/2*rank, N/2*rank+N/2) of B (K = 64, SWIZZLE_128B), filled either by generic
// stores (use_tma = 0) or by pair TMA loads that credit the leader's barrier (use_tma = 1).  D: fp32 [256][N].
// status[0] != 0 reports a timed-out wait (all waits are bounded so a wrong assumption cannot hang the GPU). ----
namespace {
__device__ __forceinline__ bool lab_wait(uint64_t* bar, uint32_t parity, int* status, int code) {
  for (int spins = 0; spins < (1 << 24); ++spins)
    if (lab_try(bar, parity)) return true;
  if (status) atomicExch(status, code);
  return false;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
lab_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __half* A,
                const __half* B, int N, int use_tma, float* D, int* status) {
  extern __shared__ uint8_t smem_raw4[];
  const uint32_t raw = ptx::smem_u32(smem_raw4);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw4 + (base - raw);
  uint8_t* sA = smem;                       // 128 rows x 128 B
  uint8_t* sB = smem + 16384;               // N/2 rows x 128 B
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 16384 + 16384);
  uint64_t* bar_done = bar_full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int nb = N / 2;
  if (tid == 0) {
    ptx::mbar_init(bar_full, 1);
    ptx::mbar_init(bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc_pair<256>(slot);
  if (!use_tma) {
    for (int i = tid; i < 128 * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) =
          *reinterpret_cast<const uint4*>(A + ((size_t)(rank * 128 + r) * 64 + c * 8));
    }
    for (int i = tid; i < nb * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) =
          *reinterpret_cast<const uint4*>(B + ((size_t)(rank * nb + r) * 64 + c * 8));
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (use_tma && tid == 0) {
    if (rank == 0) ptx::mbar_expect_tx(bar_full, (uint32_t)(2 * (16384 + nb * 128)));
    ptx::tma_load_2d_pair(sA, &tmA, bar_full, 0, (int)rank * 128);
    ptx::tma_load_2d_pair(sB, &tmB, bar_full, 0, (int)rank * nb);
  }
  if (rank == 0 && warp == 0) {
    bool ok = true;
    if (use_tma) ok = lab_wait(bar_full, 0, status, 1);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      const uint32_t hi = ptx::umma_desc_hi(1024, 2);
      const uint32_t a_lo = ((base) & 0x3FFFF) >> 4, b_lo = ((base + 16384) & 0x3FFFF) >> 4;
      const uint32_t idesc = ptx::umma_idesc_f16_pair(N);
      for (int k = 0; k < 4; ++k) ptx::umma_f16_pair(tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, k != 0);
      ptx::umma_commit_pair(bar_done);
    }
    __syncwarp();
    (void)ok;
  }
  lab_wait(bar_done, 0, status, 2 + (int)rank);
  ptx::tc_fence_after();
  const int row = (int)rank * 128 + tid;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 0) ptx::tmem_dealloc_pair<256>(tmem);
}
}  // namespace

// A: device fp16 [256][64], B: device fp16 [N][64] (N % 32 == 0, N <= 256), D: device fp32 [256][N], status: device int[1]
extern "C" int tpz_lab_umma_pair(const tpz_half* A, const tpz_half* B, int N, int use_tma, float* D, int* status, void* stream) {
  TPZ_CHECK(N % 32 == 0 && N >= 32 && N <= 256, "tpz_lab_umma_pair: bad N=%d", N);
  CUtensorMap tmA, tmB;
  uint64_t dA[2] = {64, 256}, sA[1] = {128};
  uint32_t bA[2] = {64, 128}, es[2] = {1, 1};
  int rc = tpz_encode_tmap(&tmA, A, 2, dA, sA, bA, es, 128);
  if (rc) return rc;
  uint64_t dB[2] = {64, (uint64_t)N};
  uint32_t bB[2] = {64, (uint32_t)(N / 2)};
  rc = tpz_encode_tmap(&tmB, B, 2, dB, sA, bB, es, 128);
  if (rc) return rc;
  const int smem = 16384 + 16384 + 1024 + 256;
  TPZ_CUDA(cudaFuncSetAttribute(lab_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  TPZ_CUDA(cudaMemsetAsync(status, 0, sizeof(int), reinterpret_cast<cudaStream_t>(stream)));
  lab_pair_kernel<<<2, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, reinterpret_cast<const __half*>(A),
                                                                            reinterpret_cast<const __half*>(B), N, use_tma, D,
                                                                            status);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// raw probe: one tcgen05.mma on caller-built shared-memory images and descriptors (see tpz_lab.h)
// ------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1) lab_raw_kernel(const uint4* __restrict__ imgA, int vecA, const uint4* __restrict__ imgB, int vecB,
                                                        unsigned long long descA, unsigned long long descB, uint32_t idesc, int N,
                                                        int kind, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < vecA; i += 128) reinterpret_cast<uint4*>(base)[i] = imgA[i];
  for (int i = tid; i < vecB; i += 128) reinterpret_cast<uint4*>(base + 65536)[i] = imgB[i];
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_base_s);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    if (ptx::elect_one()) {
      const uint64_t rel = (uint64_t)((ptx::smem_u32(base) & 0x3FFFF) >> 4);
      const uint64_t da = descA + rel, db = descB + rel;          // start address field: low 14 bits
      if (kind == 0) {
        ptx::umma_f16(tmem_base, da, db, idesc, 0u);
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(0u)
            : "memory");
      }
      ptx::umma_commit(&bar);
    }
    __syncwarp();
  }
  for (uint32_t spins = 0; !lab_try(&bar, 0); ++spins)
    if (spins > (1u << 26)) __trap();
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem_base);
}
}  // namespace

extern "C" int tpz_lab_umma_raw(const void* imgA, int bytesA, const void* imgB, int bytesB, unsigned long long descA,
                                unsigned long long descB, unsigned idesc, int N, int kind, float* D, void* stream) {
  TPZ_CHECK(bytesA > 0 && bytesA <= 65536 && bytesA % 16 == 0 && bytesB > 0 && bytesB <= 65536 && bytesB % 16 == 0,
            "tpz_lab_umma_raw: images must be 16..65536 bytes, multiples of 16");
  TPZ_CHECK(N % 16 == 0 && N >= 16 && N <= 256, "tpz_lab_umma_raw: bad N");
  const int smem = 2 * 65536 + 1024;
  static bool configured = false;
  if (!configured) {
    TPZ_CUDA(cudaFuncSetAttribute(lab_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  lab_raw_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(imgA), bytesA / 16,
                                                                           reinterpret_cast<const uint4*>(imgB), bytesB / 16, descA,
                                                                           descB, idesc, N, kind, D);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

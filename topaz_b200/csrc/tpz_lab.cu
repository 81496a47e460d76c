// Hardware probe ("lab") for UMMA shared-memory descriptor behaviour on sm_100a.  Not on the product path:
// it answers one design question for the halo-reuse variant of the conv kernel — can an A operand start at
// an arbitrary ROW of a TMA-written 128B-swizzled tile (start address not 1024-B aligned), and with what
// base_offset / SBO?  One CTA: TMA-load A [rows][64] and B [N][64] fp16 (SWIZZLE_128B), issue 4 MMAs
// (K = 64) with A start = row `shift`, 8-row-group stride `sbo_rows`, write D [128][N] fp32.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {
__device__ __forceinline__ bool lab_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128, 1) lab_kernel(const __grid_constant__ CUtensorMap tmA,
                                                     const __grid_constant__ CUtensorMap tmB, int rowsA, int N,
                                                     int shift, int sbo_rows, int base_off_mode, float* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_bytes = rowsA * 128, b_bytes = N * 128;
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
    for (int r = 0; r < rowsA; r += 256) ptx::tma_load_2d(smem + r * 128, &tmA, bar, 0, r);
    ptx::tma_load_2d(smem + a_bytes, &tmB, bar, 0, 0);
    while (!lab_try(bar, 0)) {}
    ptx::tc_fence_after();
    const uint32_t a_addr = base + shift * 128;
    const uint32_t b_addr = base + a_bytes;
    const uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7) : 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = ptx::umma_desc(a_addr + k * 32, sbo_rows * 128, 2, bo);
      const uint64_t db = ptx::umma_desc(b_addr + k * 32, 1024, 2, 0);
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

// A: device fp16 [rowsA][64], B: device fp16 [N][64], D: device fp32 [128][N]
extern "C" int tpz_lab_umma(const tpz_half* A, int rowsA, const tpz_half* B, int N, int shift, int sbo_rows,
                            int base_off_mode, float* D, void* stream) {
  TPZ_CHECK(rowsA % 8 == 0 && rowsA <= 1024 && N % 16 == 0 && N <= 256, "tpz_lab_umma: bad sizes");
  CUtensorMap tmA, tmB;
  uint64_t dA[2] = {64, (uint64_t)rowsA}, sA[1] = {128};
  uint32_t bA[2] = {64, (uint32_t)(rowsA < 256 ? rowsA : 256)}, es[2] = {1, 1};
  TPZ_CHECK(rowsA <= 256 || rowsA % 256 == 0, "tpz_lab_umma: rowsA > 256 must be a multiple of 256");
  int rc = tpz_encode_tmap(&tmA, A, 2, dA, sA, bA, es, 128);
  if (rc) return rc;
  uint64_t dB[2] = {64, (uint64_t)N};
  uint32_t bB[2] = {64, (uint32_t)N};
  rc = tpz_encode_tmap(&tmB, B, 2, dB, sA, bB, es, 128);
  if (rc) return rc;
  const int smem = rowsA * 128 + N * 128 + 1024 + 256;
  TPZ_CUDA(cudaFuncSetAttribute(lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lab_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, rowsA, N, shift, sbo_rows,
                                                                       base_off_mode, D);
  TPZ_CUDA(cudaGetLastError());
  return 0;
}

/* Hardware probes ("lab") for UMMA / TMA behaviour on sm_100a.  Bring-up tooling only: built into tools/lab/libtpz_lab.so by
 * __graft_entry__.build(), never linked into or loaded by the product library. */
#pragma once
#include "../../include/topaz_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int tpz_lab_umma(const tpz_half* A, int rowsA, const tpz_half* B, int N, int shift, int sbo_rows, int base_off_mode,
                 int kc, float* D, void* stream);
/* CTA-pair probe: tcgen05.mma.cta_group::2 (M = 256 over two CTAs of a cluster), operands by generic stores or pair TMA. */
int tpz_lab_umma_pair(const tpz_half* A, const tpz_half* B, int N, int use_tma, float* D, int* status, void* stream);
int tpz_lab_umma_rate(int N, int shift, int sbo_rows, int iters, int two_acc, long long* cycles, void* stream);
int tpz_lab_tma_stride(const tpz_half* A, int rowsA, int start, int stride, int nrows, tpz_half* out, void* stream);

/* One raw tcgen05.mma: the caller provides the byte images of the A and B shared-memory regions (each up to 64 KB, A at
 * offset 0 and B at offset 65536 of a 1024-byte aligned buffer), both 64-bit operand descriptors with start addresses RELATIVE to
 * that buffer, and the instruction descriptor; D [128][N] fp32 is read back from TMEM.  kind: 0 = f16, 1 = tf32. */
int tpz_lab_umma_raw(const void* imgA, int bytesA, const void* imgB, int bytesB, unsigned long long descA, unsigned long long descB,
                     unsigned idesc, int N, int kind, float* D, void* stream);
#ifdef __cplusplus
}
#endif

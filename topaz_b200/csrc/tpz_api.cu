// C-ABI plumbing: error reporting, device info, TMA descriptor encoding.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <stdarg.h>

thread_local char g_tpz_err[512] = {0};

int tpz_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_tpz_err, sizeof(g_tpz_err), fmt, ap);
  va_end(ap);
  return code ? code : 1;
}

extern "C" const char* tpz_last_error(void) { return g_tpz_err; }

extern "C" int tpz_device_info(int* num_sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  TPZ_CUDA(cudaGetDevice(&dev));
  TPZ_CUDA(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev));
  TPZ_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  TPZ_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int tpz_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TPZ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    TPZ_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else TPZ_CHECK(swizzle_bytes == 0, "bad swizzle %d", swizzle_bytes);
  TPZ_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    TPZ_CHECK(strides_bytes[i] % 16 == 0, "TMA stride %d (%llu B) must be a multiple of 16", i,
              (unsigned long long)strides_bytes[i]);
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TPZ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu,.. box %u,%u,..)",
            (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
            rank > 1 ? box[1] : 0);
  return 0;
}

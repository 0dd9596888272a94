// countr_b200 — C-ABI plumbing: error reporting, version, device probe, tensor-map encode.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

namespace countr {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

int make_tmap_4d_16b(CUtensorMap* out, const void* base, const uint64_t dims[4],
                     const uint64_t strides[4], const uint32_t box[4], TmapSwizzle swizzle) {
  return make_tmap_4d(out, base, 2, dims, strides, box, swizzle);
}

int make_tmap_4d(CUtensorMap* out, const void* base, int elt_bytes, const uint64_t dims[4], const uint64_t strides[4],
                 const uint32_t box[4], TmapSwizzle swizzle) {
  COUNTR_REQUIRE(elt_bytes == 2 || elt_bytes == 4, "tensor map element size %d unsupported", elt_bytes);
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(COUNTR_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  COUNTR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "tensor map base %p not 16-byte aligned", base);
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    COUNTR_REQUIRE(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "tensor map dim/box %d out of range (%llu, %u)", i,
                   (unsigned long long)dims[i], box[i]);
  }
  for (int i = 1; i < 4; ++i) {
    gstr[i - 1] = strides[i] * static_cast<unsigned long long>(elt_bytes);  // bytes
    COUNTR_REQUIRE((gstr[i - 1] & 15ull) == 0, "tensor map stride %d (%llu B) not a multiple of 16 B", i,
                   (unsigned long long)gstr[i - 1]);
  }
  CUtensorMapSwizzle sw = swizzle == TMAP_SW_128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == TMAP_SW_64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == TMAP_SW_32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(COUNTR_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
                     "box=(%u,%u,%u,%u)",
                     (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                     (unsigned long long)dims[3], (unsigned long long)strides[1], (unsigned long long)strides[2],
                     (unsigned long long)strides[3], box[0], box[1], box[2], box[3]);
  return COUNTR_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("COUNTR_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static int g_sm_budget = 0;   // 0 = every SM; set by countr_set_sm_budget (data-parallel runs leave a few SMs to NCCL)

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
    const char* e = getenv("COUNTR_SM_BUDGET");
    if (e != nullptr && g_sm_budget == 0) g_sm_budget = atoi(e);
  }
  // persistent kernels size their grids with this: an even number (CTA pairs), never more than the device has
  if (g_sm_budget > 0 && g_sm_budget < n) return g_sm_budget & ~1;
  return n;
}

}  // namespace countr

extern "C" {

const char* countr_last_error(void) { return countr::g_err; }

const char* countr_version(void) { return "countr_b200 0.1 (sm_100a)"; }

// Returns 0 when the current device is a Blackwell sm_100 part; a negative code otherwise.
int countr_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess)
    return countr::set_error(countr::COUNTR_ERR_CUDA, "no CUDA device: the countr_b200 kernels have no CPU fallback");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return countr::set_error(countr::COUNTR_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only",
                             major, minor);
  return 0;
}

int countr_num_sms(void) { return countr::num_sms(); }

int countr_set_sm_budget(int n) {
  countr::g_sm_budget = n > 0 ? n : 0;
  return countr::num_sms();
}

int countr_memset_zero(void* ptr, size_t bytes, countr_stream_t stream) {
  if (cudaMemsetAsync(ptr, 0, bytes, reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess)
    return countr::set_error(countr::COUNTR_ERR_CUDA, "cudaMemsetAsync failed");
  return 0;
}

}  // extern "C"

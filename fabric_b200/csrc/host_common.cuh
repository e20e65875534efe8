// Host-side helpers shared by the translation units of libfabric_b200.so: error reporting, device checks,
// TMA tensor-map construction.  (C++17 inline variables: one definition across TUs.)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/fabric_b200.h"

namespace fbh {

inline thread_local char g_err[512] = "";

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define FB_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(FB_ERR_LAUNCH, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceInfo {
  int ok = 0;  // 1 = sm_100, -1 = other
  int sms = 0;
  int smem_optin = 0;
};
inline DeviceInfo g_dev[64];
inline std::mutex g_mu;

inline int device_info(DeviceInfo* out) {
  int dev = 0;
  FB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(FB_ERR_ARG, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev[dev].ok == 0) {
    int major = 0, minor = 0;
    FB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    FB_CUDA(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    FB_CUDA(cudaDeviceGetAttribute(&g_dev[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_dev[dev].ok = (major == 10 && minor == 0) ? 1 : -1;
  }
  *out = g_dev[dev];
  if (out->ok != 1) return fail(FB_ERR_ARCH, "fabric_b200 kernels are built for sm_100a only; device %d is not", dev);
  return FB_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn g_encode = nullptr;

inline int get_encode(EncodeTiledFn* fn) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_encode) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p || q != cudaDriverEntryPointSuccess)
      return fail(FB_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    g_encode = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = g_encode;
  return FB_OK;
}

// bf16 NHWC5 tensor (C, W, H, B, G) with box (bc, bw, bh, bb, 1)
inline int make_tmap_act(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, int G, int bc, int bw, int bh, int bb,
                  CUtensorMapSwizzle sw) {
  EncodeTiledFn enc;
  int rc = get_encode(&enc);
  if (rc) return rc;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)G};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)B * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FB_ERR_LAUNCH, "cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d G=%d box %d,%d,%d,%d) -> %d", C, W, H, B, G,
                bc, bw, bh, bb, (int)r);
  return FB_OK;
}

// bf16 2-D row-major [rows][cols] with box (bcols, brows)
inline int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int brows, int bcols, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc;
  int rc = get_encode(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)bcols, (cuuint32_t)brows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FB_ERR_LAUNCH, "cuTensorMapEncodeTiled(2d %lld x %lld box %d x %d) -> %d", (long long)rows,
                (long long)cols, brows, bcols, (int)r);
  return FB_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


// grid for a grid-stride elementwise kernel
inline int ew_grid(size_t n, int block, int sms) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)sms * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace fbh

// Host-side TMA tensor-map construction (driver entry points are resolved at run time through the
// CUDA runtime so the library links only against cudart and still loads on a machine without libcuda).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace vtb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DriverApi {
  PFN_encodeTiled tiled = nullptr;
  PFN_encodeIm2col im2col = nullptr;
  bool ok = false;
};

inline const DriverApi& driver_api() {
  static DriverApi api = [] {
    DriverApi a;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.tiled = reinterpret_cast<PFN_encodeTiled>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.im2col = reinterpret_cast<PFN_encodeIm2col>(f);
    a.ok = a.tiled && a.im2col;
    return a;
  }();
  return api;
}

inline CUtensorMapSwizzle swizzle_enum(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

// 2-D row-major matrix [outer][inner] of `dt`, row pitch in bytes.
inline bool tmap_tiled_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                          uint32_t box_inner, uint32_t box_outer, int swizzle_bytes,
                          CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  const DriverApi& api = driver_api();
  if (!api.ok) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = api.tiled(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// NHWC bf16 activation view (N,H,W,C) with pixel pitch `ld` elements, walked in im2col mode.
// lower/upper = pixelBoxLower/UpperCorner in {W,H} order; stride = traversal stride.
inline bool tmap_im2col_nhwc(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int ld, int lower_w,
                             int lower_h, int upper_w, int upper_h, int chans_per_pixel, int pixels_per_col,
                             int stride, int swizzle_bytes) {
  const DriverApi& api = driver_api();
  if (!api.ok) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = api.im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper,
                          (cuuint32_t)chans_per_pixel, (cuuint32_t)pixels_per_col, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  // Driver quirk (<= 13.1): im2col maps over tensors smaller than 128 KiB carry a flag that makes the
  // load fault; clear it (same workaround NVIDIA's CUTLASS applies).
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010) {
    const uint64_t bytes = (uint64_t)N * H * W * ld * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return true;
}

}  // namespace vtb

// tmap.cuh -- host helpers: TMA tensor maps ([B,H,S,D] 16-bit tensors, 64-column boxes, SWIZZLE_128B).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"

namespace blade {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// [B,H,S,D] 16-bit tensor with element strides -> 4-D map, box = 64 x 128 x 1 x 1, SWIZZLE_128B
inline int make_tmap(CUtensorMap* map, const void* ptr, int dtype, int64_t B, int64_t H, int64_t S, int64_t D,
                     int64_t sb, int64_t sh, int64_t ss, int box_rows = 128) {
  PFN_encodeTiled enc = get_encode();
  BLADE_REQUIRE(enc != nullptr, BLADE_ERR_LAUNCH, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(ss * 2), (cuuint64_t)(sh * 2), (cuuint64_t)(sb * 2)};
  // size-1 dims may carry arbitrary strides; TMA wants multiples of 16 bytes
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0 || strides[i] == 0) strides[i] = (cuuint64_t)(D * 2);
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};  // box_rows < 128: sub-tiles of the multi-level pooled K/V
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dtype == BLADE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BLADE_REQUIRE(r == CUDA_SUCCESS, BLADE_ERR_LAUNCH, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return BLADE_OK;
}
inline int make_tmap(CUtensorMap* map, const BladeTensor* t, int box_rows = 128) {
  return make_tmap(map, t->ptr, t->dtype, t->shape[0], t->shape[1], t->shape[2], t->shape[3], t->stride[0],
                   t->stride[1], t->stride[2], box_rows);
}


}  // namespace blade

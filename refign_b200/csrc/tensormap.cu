// Host-side TMA descriptor (CUtensorMap) construction.  cuTensorMapEncodeTiled is a driver API; it is
// resolved at run time through cudaGetDriverEntryPoint so that the library does not link libcuda.
#include "rf_common.cuh"
#include "rf_sm100.cuh"

namespace rf {

rf_encode_tiled_fn get_encode_tiled() {
  static rf_encode_tiled_fn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<rf_encode_tiled_fn>(p);
  }
  return fn;
}

int make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t d0, uint64_t d1,
                 uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                 CUtensorMapSwizzle swizzle) {
  rf_encode_tiled_fn enc = get_encode_tiled();
  if (enc == nullptr) return RF_ECUDA;
  RF_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map: base address must be 16-byte aligned");
  RF_REQUIRE(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "tensor map: strides must be multiples of 16 bytes");
  RF_REQUIRE((uint64_t)box0 * elem_bytes <= 128, "tensor map: inner box exceeds the 128-byte swizzle span");
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for dims [%llu,%llu,%llu] box [%u,%u]", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1);
    return RF_ECUDA;
  }
  return RF_OK;
}

int make_tmap_nd(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  rf_encode_tiled_fn enc = get_encode_tiled();
  if (enc == nullptr) return RF_ECUDA;
  RF_REQUIRE(rank >= 2 && rank <= 5, "tensor map: rank must be 2..5");
  RF_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map: base address must be 16-byte aligned");
  RF_REQUIRE(swizzle == CU_TENSOR_MAP_SWIZZLE_NONE || (uint64_t)box[0] * elem_bytes <= 128,
             "tensor map: inner box exceeds the 128-byte swizzle span");
  RF_REQUIRE(((uint64_t)box[0] * elem_bytes) % 16 == 0, "tensor map: inner box must be a multiple of 16 bytes");
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    RF_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map: box dimension out of range");
    if (i + 1 < rank) {
      st[i] = strides[i];
      RF_REQUIRE(strides[i] % 16 == 0, "tensor map: strides must be multiples of 16 bytes");
    }
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a rank-%d map", (int)r, rank);
    return RF_ECUDA;
  }
  return RF_OK;
}

}  // namespace rf

// tmap.cu -- host-side TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library does not link libcuda and still loads on a box without a driver).
#include "common.cuh"
#include "handle.h"

namespace b200bo {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows][cols] f64 matrix with `ld` doubles per row; box = box_rows x KC doubles (128 B), 128B swizzle
static bool make2d(CUtensorMap* m, double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)KC, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t make_map2d(CUtensorMap* m, double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make2d(m, base, rows, cols, ld, box_rows) ? cudaSuccess : cudaErrorInvalidValue;
}

// [d2][d1][d0] bytes, d0 contiguous; box = box0 bytes x box1 rows x 1, 128B swizzle (the int8 slices of syrk_i8.cu)
cudaError_t make_map3d_u8(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorInvalidValue;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0, d0 * d1};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_tensor_maps(b200bo_handle_s* h) {
  const uint64_t cap = (uint64_t)h->cap, nb = cap / NB;
  bool ok = make2d(&h->tmL, h->dL, cap, cap, cap, NB);
  ok = ok && make2d(&h->tmL64, h->dL, cap, cap, cap, TILE_N);
  ok = ok && make2d(&h->tmV, h->dV, (uint64_t)h->nslots * TILE_N, cap, cap, TILE_N);
  ok = ok && make2d(&h->tmLinv, h->dLinv, nb * NB, NB, NB, NB);
  ok = ok && make2d(&h->tmLinvT, h->dLinvT, nb * NB, NB, NB, NB);
  return ok ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace b200bo

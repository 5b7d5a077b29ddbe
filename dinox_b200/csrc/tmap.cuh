// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver
// entry point, so the library does not link libcuda directly).
#pragma once
#include "common.cuh"
#include <cudaTypedefs.h>

namespace dinox {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  // cuTensorMapEncodeTiled is a DRIVER call: it fails with CUDA_ERROR_INVALID_CONTEXT on a thread that
  // has not bound the primary context yet (an autograd worker whose first CUDA work is one of our
  // launches, with every allocation served from PyTorch's cache).  A runtime call binds it.
  static thread_local bool bound = false;
  if (!bound) { cudaFree(nullptr); bound = true; }
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor, row-major: `rows` x `cols` with leading dimension ld (elements); the box is
// box_rows x 64 elements (128 B inner extent, SWIZZLE_128B).  OOB elements read as zero.
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                             int box_rows, const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  DINOX_REQUIRE(enc, DINOX_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  DINOX_REQUIRE(aligned16(base), DINOX_E_ALIGN, "%s: base pointer not 16-byte aligned", what);
  DINOX_REQUIRE((ld * 2) % 16 == 0, DINOX_E_ALIGN, "%s: leading dimension %lld not a multiple of 8 elements",
                what, (long long)ld);
  DINOX_REQUIRE(rows > 0 && cols > 0 && box_rows > 0 && box_rows <= 256, DINOX_E_BADARG, "%s: bad extent", what);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DINOX_REQUIRE(r == CUDA_SUCCESS, DINOX_E_CUDA, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
  return DINOX_OK;
}

// 3-D bf16 tensor (batch, rows, cols) row-major with batch stride `bs` elements; box 1 x box_rows x 64.
inline int make_tmap_bf16_3d(CUtensorMap* out, const void* base, int64_t batch, int64_t rows, int64_t cols,
                             int64_t ld, int64_t bs, int box_rows, const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  DINOX_REQUIRE(enc, DINOX_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  DINOX_REQUIRE(aligned16(base), DINOX_E_ALIGN, "%s: base pointer not 16-byte aligned", what);
  DINOX_REQUIRE((ld * 2) % 16 == 0 && (bs * 2) % 16 == 0, DINOX_E_ALIGN, "%s: strides not 16-byte multiples", what);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DINOX_REQUIRE(r == CUDA_SUCCESS, DINOX_E_CUDA, "%s: cuTensorMapEncodeTiled(3d) failed with CUresult %d", what, (int)r);
  return DINOX_OK;
}

// Output tensor map for TMA-store epilogues: (batch, rows, cols) row-major fp32 or bf16, leading
// dimension ld and batch stride bs in elements; box = 1 x 32 rows x row_bytes (128: SWIZZLE_128B, 64: SWIZZLE_64B) (one
// epilogue warp's staging buffer).  Stores clip at the tensor bounds, so ragged tiles need no masks.
inline int make_tmap_out_3d(CUtensorMap* out, const void* base, bool is_bf16, int64_t batch, int64_t rows,
                            int64_t cols, int64_t ld, int64_t bs, int row_bytes, const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  DINOX_REQUIRE(enc, DINOX_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = is_bf16 ? 2 : 4;
  DINOX_REQUIRE(aligned16(base), DINOX_E_ALIGN, "%s: base pointer not 16-byte aligned", what);
  DINOX_REQUIRE((ld * es) % 16 == 0 && (bs * es) % 16 == 0, DINOX_E_ALIGN, "%s: strides not 16-byte multiples", what);
  DINOX_REQUIRE(batch > 0 && rows > 0 && cols > 0 && ld >= cols, DINOX_E_BADARG, "%s: bad extent", what);
  if (batch == 1 && bs < rows * ld) bs = rows * ld;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)bs * es};
  DINOX_REQUIRE(row_bytes == 128 || row_bytes == 64, DINOX_E_BADARG, "%s: staging rows are 64 or 128 bytes", what);
  cuuint32_t box[3] = {(cuuint32_t)(row_bytes / es), 32u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DINOX_REQUIRE(r == CUDA_SUCCESS, DINOX_E_CUDA, "%s: cuTensorMapEncodeTiled(out) failed with CUresult %d", what, (int)r);
  return DINOX_OK;
}

}  // namespace dinox

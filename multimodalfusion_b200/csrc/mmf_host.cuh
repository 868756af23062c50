// mmf_host.cuh — host-side helpers shared by the C-ABI entry points: error codes,
// TMA descriptor (CUtensorMap) construction through the driver entry point (no link-time
// dependency on libcuda, so the library also loads on a box without a driver), launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mmf_b200.h"

namespace mmf {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2-D row-major bf16 array [rows][cols] with leading dimension `ld` (elements); box is
// [box_rows][64 cols] (128-byte rows) with the 128-byte swizzle. Out-of-bounds elements of a
// box read as zero and are dropped on store.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MMF_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld * 2) % 16 != 0) return MMF_E_ALIGN;
  if (rows == 0 || cols == 0) return MMF_E_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMF_OK : MMF_E_TMAP;
}

// 2-D row-major array of 2-byte elements, box [box_rows][32 cols] (64-byte rows, 64-byte swizzle): the target of the
// training forward's activation-stash stores (each epilogue warp stages a 32 x 32 fp16 block whose 16-byte chunk index is
// XORed with (row >> 1) & 3 — exactly CU_TENSOR_MAP_SWIZZLE_64B); rows past `rows` are dropped by the TMA.
inline int make_tmap_2b_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MMF_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld * 2) % 16 != 0) return MMF_E_ALIGN;
  if (rows == 0 || cols == 0) return MMF_E_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMF_OK : MMF_E_TMAP;
}

// 2-D row-major fp32 array, box [box_rows][32 cols] (128-byte rows, 128-byte swizzle): the target of the
// split-K epilogue's TMA reduce-add (cp.reduce.async.bulk.tensor ... .add); out-of-bounds elements are dropped.
inline int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MMF_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld * 4) % 16 != 0) return MMF_E_ALIGN;
  if (rows == 0 || cols == 0) return MMF_E_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMF_OK : MMF_E_TMAP;
}

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mmf: CUDA launch error: %s\n", cudaGetErrorString(e));
    return -(1000 + (int)e);
  }
  return MMF_OK;
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may be scheduled while the
// previous kernel in the stream drains; every kernel launched this way calls griddep_wait() before it
// touches global memory. MMF_NO_PDL=1 falls back to plain stream order (A/B timing, debugging).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMF_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
  if (e != cudaSuccess) {
    fprintf(stderr, "mmf: CUDA launch error: %s\n", cudaGetErrorString(e));
    return -(1000 + (int)e);
  }
  return launch_status();
}

#define MMF_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != MMF_OK) return _rc; \
  } while (0)

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? MMF_OK : -(1000 + (int)e); }

}  // namespace mmf

// Host-side plumbing shared by the C-ABI translation units: error convention, the per-device context,
// TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/dfdclip_b200.h"

struct dfd_ctx {
  int device;
  int num_sms;
  int smem_optin;
  void* encode_tiled;  // PFN_cuTensorMapEncodeTiled
};

namespace dfd {

// Thread-local last-error string (dfd_last_error). Returns `code` so callers can `return fail(...)`.
int fail(int code, const char* fmt, ...);
void clear_error();

#define DFD_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::dfd::fail(DFD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                 \
  } while (0)

#define DFD_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::dfd::fail(DFD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define DFD_TRY(expr)       \
  do {                      \
    int _rc = (expr);       \
    if (_rc != 0) return _rc; \
  } while (0)

// 2-D row-major tensor map: `rows` x `cols` elements, row pitch `ld` elements, box `box_rows` x `box_cols`,
// 128-byte swizzle (box_cols * elem_bytes must be 128).
int make_tmap_2d(const dfd_ctx* ctx, CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

}  // namespace dfd

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

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

// Optional per-kernel timing (dfd_timing_*): CUDA event pairs recorded on the launching stream around each
// kernel of the hot path, accumulated per tag. Off by default.
enum {
  DFD_TAG_PATCHIFY = 0,
  DFD_TAG_GEMM_PATCH,
  DFD_TAG_LAYERNORM,
  DFD_TAG_GEMM_QKV,
  DFD_TAG_MHA,
  DFD_TAG_GEMM_OUT,
  DFD_TAG_GEMM_FC,
  DFD_TAG_GEMM_PROJ,
  DFD_TAG_DEC_ATTN,
  DFD_TAG_DEC_LINEAR,
  DFD_TAG_DEC_OTHER,
  DFD_TAG_ADAPTER,
  DFD_TAG_COUNT
};

struct dfd_timing_slot {
  int tag;
  cudaEvent_t start, stop;
};

// Key of one encoded TMA descriptor: a descriptor depends only on these values, so repeated launches on the same
// buffers (the steady state of a predict / training loop: the caching allocator hands back the same blocks) reuse it
// instead of paying cuTensorMapEncodeTiled ~150 times per predict.
struct dfd_tmap_key {
  const void* base;
  uint64_t d0, d1, d2, ld, ld2;
  uint32_t box0, box1, misc;  // misc = dtype | elem_bytes << 8 | rank << 16 | swizzle << 24
  bool operator==(const dfd_tmap_key& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && ld2 == o.ld2 && box0 == o.box0 &&
           box1 == o.box1 && misc == o.misc;
  }
};
struct dfd_tmap_key_hash {
  size_t operator()(const dfd_tmap_key& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.d0, k.d1, k.d2, k.ld, k.ld2, static_cast<uint64_t>(k.box0) << 32 | k.box1,
                       static_cast<uint64_t>(k.misc)})
      h = (h ^ v) * 0x100000001B3ull + (h >> 29);
    return static_cast<size_t>(h);
  }
};

struct dfd_ctx {
  int device;
  int num_sms;
  int smem_optin;
  void* encode_tiled;  // PFN_cuTensorMapEncodeTiled
  mutable bool timing = false;
  mutable std::vector<dfd_timing_slot> slots;
  mutable size_t slots_used = 0;
  // dfd_predict_forward: the decoder blocks run on this stream beside the encoder layers that follow their tap
  // (created on first use; high priority so that the small decoder kernels are placed as soon as SMs free up)
  mutable cudaStream_t side_stream = nullptr;
  mutable cudaEvent_t fork_event = nullptr, join_event = nullptr;
  mutable std::vector<cudaEvent_t> tap_events;
  // encoded TMA descriptors (see dfd_tmap_key); guarded by tmap_mutex, cleared when it reaches tmap_cache_max entries
  mutable std::unordered_map<dfd_tmap_key, CUtensorMap, dfd_tmap_key_hash> tmap_cache;
  mutable std::mutex tmap_mutex;
  static constexpr size_t tmap_cache_max = 8192;
};

namespace dfd {

// One decoder pass split into steps (decoder.cu); dfd_decoder_forward runs them back to back on one stream,
// dfd_predict_forward issues block i on the context's side stream as soon as the encoder has produced tap i.
struct DecoderRun {
  const dfd_ctx* ctx;
  int D, H, n_blocks;
  const dfd_decoder_weights* w;
  const dfd_kv_taps* taps;
  const uint8_t* mask;
  int B, T, P;
  float* block_out;
  float* video_feature;
  void* workspace;
  size_t workspace_bytes;
  int begin(cudaStream_t stream);
  int block(int i, cudaStream_t stream);
  int end(cudaStream_t stream);
};

// The activation-saving decoder forward of the training step (decoder_train.cu) in the same steps; dfd_train_forward
// issues block i on the side stream behind its tap, dfd_decoder_train_forward runs them back to back.
struct DecoderTrainRun {
  const dfd_ctx* ctx;
  int D, H, n_blocks;
  const dfd_decoder_weights* w;
  const dfd_kv_taps* taps;
  const uint8_t* mask;
  int B, T, P;
  float* block_out;
  void* saved;
  size_t saved_bytes;
  int begin(cudaStream_t stream);
  int block(int i, cudaStream_t stream);
};

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

// 3-D tensor map over a [frames][rows][cols] view: per-frame row count `rows` (rows beyond it are out of bounds:
// zero-filled on load, clipped on store), row pitch `ld` elements, frame pitch `frame_ld` elements.
int make_tmap_3d(const dfd_ctx* ctx, CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 uint64_t frames, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t frame_ld, uint32_t box_rows,
                 uint32_t box_cols, bool swizzle128 = true);

// RAII event pair around one kernel launch (no-op unless timing is enabled on the context).
struct ScopedTimer {
  const dfd_ctx* ctx;
  cudaStream_t stream;
  dfd_timing_slot* slot = nullptr;
  ScopedTimer(const dfd_ctx* c, int tag, cudaStream_t s) : ctx(c), stream(s) {
    if (c && c->timing && c->slots_used < c->slots.size()) {
      slot = &c->slots[c->slots_used++];
      slot->tag = tag;
      cudaEventRecord(slot->start, stream);
    }
  }
  ~ScopedTimer() {
    if (slot) cudaEventRecord(slot->stop, stream);
  }
};

}  // namespace dfd

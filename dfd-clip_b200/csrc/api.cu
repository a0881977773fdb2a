// C-ABI plumbing: version, thread-local error string, per-device context, TMA tensor-map construction.
#include "host_common.h"

#include <cudaTypedefs.h>

namespace dfd {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void clear_error() { g_err[0] = 0; }

static bool tmap_lookup(const dfd_ctx* ctx, const dfd_tmap_key& key, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(ctx->tmap_mutex);
  auto it = ctx->tmap_cache.find(key);
  if (it == ctx->tmap_cache.end()) return false;
  *out = it->second;
  return true;
}

static void tmap_store(const dfd_ctx* ctx, const dfd_tmap_key& key, const CUtensorMap& map) {
  std::lock_guard<std::mutex> lock(ctx->tmap_mutex);
  if (ctx->tmap_cache.size() >= dfd_ctx::tmap_cache_max) ctx->tmap_cache.clear();
  ctx->tmap_cache.emplace(key, map);
}

int make_tmap_2d(const dfd_ctx* ctx, CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  if (!ctx || !ctx->encode_tiled) return fail(DFD_ERR_INVALID, "tensor map: context has no driver entry point");
  const dfd_tmap_key key{base, cols, rows, 0, ld, 0, box_cols, box_rows,
                         static_cast<uint32_t>(dtype) | static_cast<uint32_t>(elem_bytes) << 8 | 2u << 16 | 1u << 24};
  if (tmap_lookup(ctx, key, out)) return 0;
  if (box_cols * elem_bytes != 128) return fail(DFD_ERR_INVALID, "tensor map: box inner extent must be 128 bytes");
  if ((ld * elem_bytes) % 16 != 0) return fail(DFD_ERR_INVALID, "tensor map: row pitch must be a multiple of 16 bytes");
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail(DFD_ERR_INVALID, "tensor map: base must be 16-byte aligned");
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ctx->encode_tiled);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DFD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)",
                (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  tmap_store(ctx, key, *out);
  return 0;
}

int make_tmap_3d(const dfd_ctx* ctx, CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 uint64_t frames, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t frame_ld, uint32_t box_rows,
                 uint32_t box_cols, bool swizzle128) {
  if (!ctx || !ctx->encode_tiled) return fail(DFD_ERR_INVALID, "tensor map: context has no driver entry point");
  if (swizzle128 && box_cols * elem_bytes != 128)
    return fail(DFD_ERR_INVALID, "tensor map: box inner extent must be 128 bytes");
  if (!swizzle128 && (box_cols * elem_bytes) % 16 != 0)
    return fail(DFD_ERR_INVALID, "tensor map: box inner extent must be a multiple of 16 bytes");
  if ((ld * elem_bytes) % 16 != 0 || (frame_ld * elem_bytes) % 16 != 0)
    return fail(DFD_ERR_INVALID, "tensor map: pitches must be multiples of 16 bytes");
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail(DFD_ERR_INVALID, "tensor map: base must be 16-byte aligned");
  if (box_rows > 256) return fail(DFD_ERR_INVALID, "tensor map: box rows > 256");
  const dfd_tmap_key key{base, cols, rows, frames, ld, frame_ld, box_cols, box_rows,
                         static_cast<uint32_t>(dtype) | static_cast<uint32_t>(elem_bytes) << 8 | 3u << 16 |
                             (swizzle128 ? 1u : 0u) << 24};
  if (tmap_lookup(ctx, key, out)) return 0;
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ctx->encode_tiled);
  cuuint64_t dims[3] = {cols, rows, frames};
  cuuint64_t strides[2] = {ld * static_cast<uint64_t>(elem_bytes), frame_ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(out, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DFD_ERR_CUDA, "cuTensorMapEncodeTiled (3d) failed with CUresult %d", (int)r);
  tmap_store(ctx, key, *out);
  return 0;
}

}  // namespace dfd

extern "C" {

int dfd_version(void) { return DFD_ABI_VERSION; }

const char* dfd_last_error(void) { return dfd::g_err; }

int dfd_ctx_create(int device, dfd_ctx** out) {
  dfd::clear_error();
  if (!out) return dfd::fail(DFD_ERR_INVALID, "dfd_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  DFD_CUDA_OK(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return dfd::fail(DFD_ERR_INVALID, "dfd_ctx_create: no such device %d", device);
  cudaDeviceProp prop;
  DFD_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return dfd::fail(DFD_ERR_UNSUPPORTED,
                     "dfd_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)",
                     device, prop.major, prop.minor);
  DFD_CUDA_OK(cudaSetDevice(device));
  DFD_CUDA_OK(cudaFree(0));  // make sure the primary context exists
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DFD_CUDA_OK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn)
    return dfd::fail(DFD_ERR_CUDA, "dfd_ctx_create: cuTensorMapEncodeTiled not available from the driver");
  dfd_ctx* c = new dfd_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  c->encode_tiled = fn;
  // side stream + fork/join/tap events of dfd_predict_forward: created here, not on first use, so that a first
  // predict inside a CUDA-graph capture creates nothing (highest priority: the decoder's small kernels are placed as
  // soon as SMs free up)
  int lo = 0, hi = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork_event, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->join_event, cudaEventDisableTiming);
  for (int i = 0; i < 64 && e == cudaSuccess; ++i) {
    cudaEvent_t ev;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) c->tap_events.push_back(ev);
  }
  if (e != cudaSuccess) {
    dfd_ctx_destroy(c);
    return dfd::fail(DFD_ERR_CUDA, "dfd_ctx_create: stream/event creation failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return 0;
}

int dfd_ctx_destroy(dfd_ctx* ctx) {
  dfd::clear_error();
  if (ctx) {
    for (auto& s : ctx->slots) {
      cudaEventDestroy(s.start);
      cudaEventDestroy(s.stop);
    }
    for (auto& e : ctx->tap_events) cudaEventDestroy(e);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->join_event) cudaEventDestroy(ctx->join_event);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  }
  delete ctx;
  return 0;
}

int dfd_timing_enable(dfd_ctx* ctx, int on) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_timing_enable: ctx is NULL");
  if (on && ctx->slots.empty()) {
    ctx->slots.resize(8192);
    for (auto& s : ctx->slots) {
      DFD_CUDA_OK(cudaEventCreate(&s.start));
      DFD_CUDA_OK(cudaEventCreate(&s.stop));
    }
  }
  ctx->slots_used = 0;
  ctx->timing = on != 0;
  return 0;
}

int dfd_timing_read(dfd_ctx* ctx, int max_tags, float* total_ms, int* counts) {
  dfd::clear_error();
  if (!ctx || !total_ms || !counts) return dfd::fail(DFD_ERR_INVALID, "dfd_timing_read: null pointer");
  const int n = max_tags < DFD_TAG_COUNT ? max_tags : DFD_TAG_COUNT;
  for (int i = 0; i < n; ++i) {
    total_ms[i] = 0.f;
    counts[i] = 0;
  }
  for (size_t i = 0; i < ctx->slots_used; ++i) {
    const auto& s = ctx->slots[i];
    DFD_CUDA_OK(cudaEventSynchronize(s.stop));
    float ms = 0.f;
    DFD_CUDA_OK(cudaEventElapsedTime(&ms, s.start, s.stop));
    if (s.tag < n) {
      total_ms[s.tag] += ms;
      counts[s.tag] += 1;
    }
  }
  ctx->slots_used = 0;
  return DFD_OK;
}

int dfd_timing_num_tags(void) { return DFD_TAG_COUNT; }

const char* dfd_timing_tag_name(int tag) {
  static const char* names[DFD_TAG_COUNT] = {"patchify",  "gemm_patch_embed", "layernorm",  "gemm_qkv",
                                             "mha",       "gemm_out_proj",    "gemm_c_fc",  "gemm_c_proj",
                                             "dec_attn",  "dec_linear",       "dec_other",  "adapter"};
  return (tag >= 0 && tag < DFD_TAG_COUNT) ? names[tag] : "?";
}

}  // extern "C"

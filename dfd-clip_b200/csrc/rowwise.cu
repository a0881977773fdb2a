// HBM-bound row kernels of the encoder: LayerNorm (with the fused cls/positional-embedding add for ln_pre),
// patch extraction + fp32->bf16 cast, and the weight cast used when packing parameters.
// All are vectorised (128-bit loads/stores), coalesced along the row, warp-shuffle reductions, no smem staging
// needed (each element is touched once).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

// ------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row lives in registers (VEC float4 per lane, D = 128 * VEC).
// Reference: LayerNorm.forward src/clip/model.py:157-163 (fp32 math, eps 1e-5, biased variance);
// with `pos`: x = cat(cls, patches) + positional_embedding (model.py:280-291) folded into ln_pre (:292).
// fold_bf16 / fold_stats (OUT_F32 only; used for ln_pre when LayerNorm is folded into the GEMMs, encoder.cu): also
// write the bf16 copy of the output row and its (sum, sum of squares) into slot 0 of the row's `slots` partial pairs
// (the other slots are zeroed), the format the DFD_EPI_*_LNFOLD GEMM epilogues read.
template <int VEC, bool OUT_F32>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ pos, int pos_period, void* __restrict__ out, int64_t rows,
                 __nv_bfloat16* __restrict__ fold_bf16 = nullptr, float* __restrict__ fold_stats = nullptr,
                 int slots = 0) {
  constexpr int D = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = xr[i * 32 + lane];
  if (pos != nullptr) {
    const float4* pr = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(row % pos_period) * D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 p = __ldg(pr + i * 32 + lane);
      v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float o1 = 0.f, o2 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if constexpr (OUT_F32) {
      reinterpret_cast<float4*>(static_cast<float*>(out) + row * D)[i * 32 + lane] = y;
      if (fold_bf16) {
        reinterpret_cast<uint2*>(fold_bf16 + row * D)[i * 32 + lane] =
            make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        o1 += (y.x + y.y) + (y.z + y.w);
        o2 += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
      }
    } else {
      uint2 p = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + row * D)[i * 32 + lane] = p;
    }
  }
  if constexpr (OUT_F32) {
    if (fold_stats) {
      o1 = warp_sum(o1);
      o2 = warp_sum(o2);
      float2* st = reinterpret_cast<float2*>(fold_stats) + row * slots;
      for (int sl = lane; sl < slots; sl += 32) st[sl] = (sl == 0) ? make_float2(o1, o2) : make_float2(0.f, 0.f);
    }
  }
}

template <int VEC>
static int launch_ln(const float* x, const float* g, const float* b, const float* pos, int pos_period, void* out_bf16,
                     float* out_f32, int64_t rows, cudaStream_t stream, void* fold_bf16, float* fold_stats,
                     int slots) {
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  if (out_f32)
    layernorm_kernel<VEC, true><<<grid, warps * 32, 0, stream>>>(x, g, b, pos, pos_period, out_f32, rows,
                                                                static_cast<__nv_bfloat16*>(fold_bf16), fold_stats,
                                                                slots);
  else
    layernorm_kernel<VEC, false><<<grid, warps * 32, 0, stream>>>(x, g, b, pos, pos_period, out_bf16, rows);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int pos_period, void* out_bf16,
              float* out_f32, int64_t rows, int D, cudaStream_t stream, void* fold_bf16, float* fold_stats, int slots) {
  DFD_CHECK_ARG((fold_bf16 == nullptr) == (fold_stats == nullptr) && (fold_bf16 == nullptr || (out_f32 && slots > 0)),
                "layernorm: the folded outputs need out_f32, a bf16 buffer, a statistics buffer and slots > 0");
  DFD_CHECK_ARG(x && gamma && beta, "layernorm: null pointer");
  DFD_CHECK_ARG((out_bf16 != nullptr) != (out_f32 != nullptr), "layernorm: exactly one output must be given");
  DFD_CHECK_ARG(pos == nullptr || pos_period > 0, "layernorm: pos_period must be positive");
  if (rows == 0) return 0;
  DFD_CHECK_ARG(rows > 0, "layernorm: negative row count");
  switch (D) {
    case 128: return launch_ln<1>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 256: return launch_ln<2>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 384: return launch_ln<3>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 512: return launch_ln<4>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 640: return launch_ln<5>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 768: return launch_ln<6>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 896: return launch_ln<7>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 1024: return launch_ln<8>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 1280: return launch_ln<10>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 1536: return launch_ln<12>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    case 2048: return launch_ln<16>(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, stream, fold_bf16, fold_stats, slots);
    default: return fail(DFD_ERR_INVALID, "layernorm: unsupported width D=%d", D);
  }
}

// -------------------------------------------------------------------------------------------- patchify
// frames fp32 [F,3,R,R] -> bf16 [F*(P+1), Kp]: conv1 with stride = kernel = patch is a GEMM over these rows
// (src/clip/model.py:277-279). One thread moves 8 consecutive pixels of one image row (two float4 loads, one
// 16-byte store); consecutive threads walk an image row, so loads are fully coalesced and every 32-byte store
// sector is written whole. blockIdx.y = frame.
// U8: frames are uint8 pixels; (x / 255 - mean[c]) / std[c] — ConvertImageDtype(float32) + Normalize of the
// reference's CPU transform (src/models.py:762-768), same fp32 operation order — is applied on the fly.
template <bool U8>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ frames_raw, __nv_bfloat16* __restrict__ out, int R, int patch, int G, int Kp,
                float3 mean, float3 stdv) {
  const int f = blockIdx.y;
  const int per_row = R / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 3 * R * (R/8)
  if (idx >= 3 * R * per_row) return;
  const int xs = (idx % per_row) * 8;
  const int y = (idx / per_row) % R;
  const int c = idx / (per_row * R);
  const int64_t off = ((static_cast<int64_t>(f) * 3 + c) * R + y) * R + xs;
  float vals[8];
  if constexpr (U8) {
    const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(frames_raw) + off);
    const float mu = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
    const float sd = c == 0 ? stdv.x : (c == 1 ? stdv.y : stdv.z);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t word = e < 4 ? raw.x : raw.y;
      const float px = static_cast<float>((word >> (8 * (e & 3))) & 0xffu);
      vals[e] = __fdiv_rn(__fsub_rn(__fdiv_rn(px, 255.0f), mu), sd);
    }
  } else {
    const float* src = static_cast<const float*>(frames_raw) + off;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    vals[0] = a.x; vals[1] = a.y; vals[2] = a.z; vals[3] = a.w;
    vals[4] = b.x; vals[5] = b.y; vals[6] = b.z; vals[7] = b.w;
  }
  const int py = y / patch, i = y % patch;
  const int P = G * G;
  // 8 consecutive pixels may straddle two patches when patch % 8 != 0 (e.g. patch 14): split per element then.
  if ((patch & 7) == 0) {
    const int px = xs / patch, j = xs % patch;
    const int64_t row = static_cast<int64_t>(f) * (P + 1) + 1 + py * G + px;
    const int col = (c * patch + i) * patch + j;
    uint4 pk = make_uint4(pack_bf16(vals[0], vals[1]), pack_bf16(vals[2], vals[3]), pack_bf16(vals[4], vals[5]),
                          pack_bf16(vals[6], vals[7]));
    *reinterpret_cast<uint4*>(out + row * Kp + col) = pk;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int x = xs + e;
      const int px = x / patch, j = x % patch;
      if (px >= G) continue;
      const int64_t row = static_cast<int64_t>(f) * (P + 1) + 1 + py * G + px;
      out[row * Kp + (c * patch + i) * patch + j] = __float2bfloat16_rn(vals[e]);
    }
  }
}

// zero the cls rows and the K padding columns
__global__ void patchify_pad_kernel(__nv_bfloat16* __restrict__ out, int n_frames, int P, int K, int Kp) {
  const int64_t rows = static_cast<int64_t>(n_frames) * (P + 1);
  const int pad = Kp - K;
  // part 1: cls rows (Kp elements each)
  const int64_t n1 = static_cast<int64_t>(n_frames) * Kp;
  const int64_t n2 = rows * pad;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n1 + n2;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (t < n1) {
      out[(t / Kp) * (P + 1) * Kp + (t % Kp)] = __float2bfloat16_rn(0.f);
    } else {
      const int64_t u = t - n1;
      out[(u / pad) * Kp + K + (u % pad)] = __float2bfloat16_rn(0.f);
    }
  }
}

int patchify(const void* frames, void* out, int n_frames, int R, int patch, int Kp, cudaStream_t stream,
             const float* mean_std /* NULL: fp32 frames; else uint8 frames with {mean[3], std[3]} */) {
  DFD_CHECK_ARG(frames && out, "patchify: null pointer");
  DFD_CHECK_ARG(n_frames >= 0 && R > 0 && patch > 0, "patchify: bad shape");
  DFD_CHECK_ARG(R % 8 == 0, "patchify: image size %d must be a multiple of 8", R);
  const int G = R / patch;
  DFD_CHECK_ARG(G > 0, "patchify: patch larger than image");
  const int K = 3 * patch * patch;
  DFD_CHECK_ARG(Kp >= K && Kp % 8 == 0, "patchify: Kp=%d must be >= %d and a multiple of 8", Kp, K);
  if (n_frames == 0) return 0;
  // pixels right of / below the last full patch are dropped exactly like a strided conv does
  const int per_frame = 3 * R * (R / 8);
  dim3 grid((per_frame + 255) / 256, n_frames);
  if (G * patch != R) {
    // rows beyond G*patch would index py >= G: mask by zero-filling first and guarding below is not needed for
    // the CLIP shapes (224 = 14*16 = 16*14); reject anything else instead of silently diverging.
    return fail(DFD_ERR_INVALID, "patchify: image size %d is not a multiple of patch %d", R, patch);
  }
  patchify_pad_kernel<<<296, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(out), n_frames, G * G, K, Kp);
  DFD_CUDA_OK(cudaGetLastError());
  if (mean_std) {
    DFD_CHECK_ARG(mean_std[3] != 0.f && mean_std[4] != 0.f && mean_std[5] != 0.f, "patchify: zero std");
    DFD_CHECK_ARG(reinterpret_cast<uintptr_t>(frames) % 8 == 0, "patchify: uint8 frames must be 8-byte aligned");
    patchify_kernel<true><<<grid, 256, 0, stream>>>(frames, static_cast<__nv_bfloat16*>(out), R, patch, G, Kp,
                                                    make_float3(mean_std[0], mean_std[1], mean_std[2]),
                                                    make_float3(mean_std[3], mean_std[4], mean_std[5]));
  } else {
    patchify_kernel<false><<<grid, 256, 0, stream>>>(frames, static_cast<__nv_bfloat16*>(out), R, patch, G, Kp,
                                                     make_float3(0, 0, 0), make_float3(1, 1, 1));
  }
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------- fp32 -> bf16 cast
// dst[r, 0:cols] = bf16(src[r, 0:cols]); dst[r, cols:dst_ld] = 0. Used once, when packing weights.
__global__ void cast_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int cols,
                                int dst_ld) {
  const int64_t total = rows * dst_ld;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = t / dst_ld;
    const int c = static_cast<int>(t % dst_ld);
    dst[t] = __float2bfloat16_rn(c < cols ? src[r * cols + c] : 0.f);
  }
}

// LayerNorm folded into a Linear (encoder.cu): one warp per output row n of W fp32 [N, K]:
//   wf[n, k]  = bf16(gamma[k] * W[n, k])           the GEMM operand
//   colsum[n] = sum_k float(wf[n, k])              (of the ROUNDED operand: the mean correction must cancel exactly
//                                                   what the tensor core multiplies)
//   bias_f[n] = bias[n] + sum_k beta[k] * W[n, k]
__global__ void __launch_bounds__(256)
fold_ln_linear_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ wf, float* __restrict__ colsum,
                      float* __restrict__ bias_f, int N, int K) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[static_cast<int64_t>(n) * K + k];
    const __nv_bfloat16 r = __float2bfloat16_rn(gamma[k] * w);
    wf[static_cast<int64_t>(n) * K + k] = r;
    cs += __bfloat162float(r);
    bs = fmaf(beta[k], w, bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = bias[n] + bs;
  }
}

int fold_ln_linear(const float* W, const float* bias, const float* gamma, const float* beta, void* wf, float* colsum,
                   float* bias_f, int N, int K, cudaStream_t stream) {
  fold_ln_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(W, bias, gamma, beta, static_cast<__nv_bfloat16*>(wf), colsum,
                                                         bias_f, N, K);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int cast_pad_bf16(const float* src, void* dst, int64_t rows, int cols, int dst_ld, cudaStream_t stream) {
  if (rows == 0) return 0;
  cast_pad_kernel<<<592, 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols, dst_ld);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

extern "C" int dfd_layernorm(dfd_ctx* ctx, const float* x, const float* gamma, const float* beta, const float* pos,
                             int pos_period, void* out_bf16, float* out_f32, int64_t rows, int D, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_layernorm: ctx is NULL");
  return dfd::layernorm(x, gamma, beta, pos, pos_period, out_bf16, out_f32, rows, D, static_cast<cudaStream_t>(stream),
                        nullptr, nullptr, 0);
}

extern "C" int dfd_patchify(dfd_ctx* ctx, const float* frames, void* out_bf16, int n_frames, int R, int patch, int Kp,
                            void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_patchify: ctx is NULL");
  return dfd::patchify(frames, out_bf16, n_frames, R, patch, Kp, static_cast<cudaStream_t>(stream), nullptr);
}

extern "C" int dfd_patchify_u8(dfd_ctx* ctx, const uint8_t* frames, const float* mean_std, void* out_bf16,
                               int n_frames, int R, int patch, int Kp, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_patchify_u8: ctx is NULL");
  if (!mean_std) return dfd::fail(DFD_ERR_INVALID, "dfd_patchify_u8: mean_std is NULL");
  return dfd::patchify(frames, out_bf16, n_frames, R, patch, Kp, static_cast<cudaStream_t>(stream), mean_std);
}

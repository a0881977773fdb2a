// CompInvAdapter on the tapped K/V (src/models.py:783-940, called at :546-547): per tapped layer and per {k, v} a
// bottleneck  Linear(D -> x, no bias) -> [LayerNorm / GELU in the order the struct type fixes] -> Linear(x -> D, no
// bias)  whose output is added to the tap (residual) — or, for the "linear" struct, replaces it.
//
// The taps are bf16 column slices of the encoder's packed QKV buffers ([rows, 3D], K at column D, V at 2D), so one
// adapter application is
//   H  = tap . Wdown^T                  tcgen05 GEMM straight out of the strided slice   (A pitch = 3D)
//   H  = act(H)                         row kernel below, in place (fp32 math on the bf16 H)
//   tap += H . Wup^T                    tcgen05 GEMM whose epilogue is a bf16 TMA reduce-add into the slice
// Nothing is copied out of or back into the QKV buffer; the decoder then streams the adapted K/V in place.
#include "common.cuh"
#include "host_common.h"

namespace dfd {

int gemm_bf16(const dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
              void* out, int64_t ldo, int M, int N, int K, int epilogue, cudaStream_t stream);

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---------------------------------------------------------------------------------- per-token LayerNorm + GELU
// One warp per row of H bf16 [rows, X], X = 256 * NV, in place. GELU_FIRST: LayerNorm(GELU(h)) ("768-x-768",
// "legacy-768-x-768", src/models.py:797-818); otherwise GELU(LayerNorm(h)) ("768-x-768-ln", "-z0", :835-861).
// nn.LayerNorm(x): fp32, eps 1e-5, biased variance; nn.GELU(): erf form.
template <int NV, bool GELU_FIRST>
__global__ void __launch_bounds__(256)
adapter_rownorm_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                       int64_t rows) {
  constexpr int X = 256 * NV;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint4* hr = reinterpret_cast<uint4*>(h + row * X);
  float v[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    unpack8(hr[i * 32 + lane], v[i]);
    if constexpr (GELU_FIRST) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = gelu_erf(v[i][j]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  const float mean = warp_sum(s) * (1.0f / X);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      ss += d * d;
    }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / X) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      if constexpr (!GELU_FIRST) y[j] = gelu_erf(y[j]);
    }
    hr[i * 32 + lane] = pack8(y);
  }
}

// ------------------------------------------------------------------- per-frame LayerNorm((P, x)) + GELU ("nln")
// "768-x-768-nln" (src/models.py:819-834): nn.LayerNorm((patches, x)) normalises each frame's P*x values jointly,
// with an elementwise affine of shape [P, x]; GELU follows. One CTA per frame (group of `group_rows` rows whose
// first `group_skip` rows — the CLS token — take no part and are zeroed). Three passes over the frame's 100 KB
// (mean, variance about the mean, apply): it stays in L1/L2, DRAM sees one read and one write.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

__global__ void __launch_bounds__(1024)
adapter_groupnorm_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                         int X, int group_rows, int group_skip) {
  __shared__ float red[32];
  const int P = group_rows - group_skip;
  const int chunks_per_row = X / 8;
  const int n_chunks = P * chunks_per_row;
  uint4* base = reinterpret_cast<uint4*>(h + (static_cast<int64_t>(blockIdx.x) * group_rows + group_skip) * X);
  float s = 0.f;
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    float f[8];
    unpack8(base[c], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  const float inv_n = 1.0f / (static_cast<float>(P) * static_cast<float>(X));
  const float mean = block_sum(s, red) * inv_n;
  float ss = 0.f;
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    float f[8];
    unpack8(base[c], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mean;
      ss += d * d;
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) * inv_n + 1e-5f);
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    float f[8], y[8];
    unpack8(base[c], f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = gelu_erf((f[j] - mean) * rstd * g[j] + b[j]);
    base[c] = pack8(y);
  }
  // rows excluded from the statistics (CLS): defined contents for the up-projection that runs over every row
  uint4* skip = reinterpret_cast<uint4*>(h + static_cast<int64_t>(blockIdx.x) * group_rows * X);
  for (int c = threadIdx.x; c < group_skip * chunks_per_row; c += blockDim.x) skip[c] = make_uint4(0, 0, 0, 0);
}

template <bool GELU_FIRST>
static int launch_rownorm(void* h, const float* g, const float* b, int64_t rows, int X, cudaStream_t stream) {
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(h);
  switch (X / 256) {
    case 1: adapter_rownorm_kernel<1, GELU_FIRST><<<grid, warps * 32, 0, stream>>>(hp, g, b, rows); break;
    case 2: adapter_rownorm_kernel<2, GELU_FIRST><<<grid, warps * 32, 0, stream>>>(hp, g, b, rows); break;
    case 3: adapter_rownorm_kernel<3, GELU_FIRST><<<grid, warps * 32, 0, stream>>>(hp, g, b, rows); break;
    case 4: adapter_rownorm_kernel<4, GELU_FIRST><<<grid, warps * 32, 0, stream>>>(hp, g, b, rows); break;
    default: return fail(DFD_ERR_INVALID, "adapter: inner width %d not in {256, 512, 768, 1024}", X);
  }
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// "768-bn": kv[row, :] += scale[f] * h[row, :] + shift[f], f = row / group_rows. One thread per 8 channels.
__global__ void __launch_bounds__(256)
adapter_frame_affine_kernel(__nv_bfloat16* __restrict__ kv, int64_t ld, const __nv_bfloat16* __restrict__ h,
                            const float* __restrict__ scale, const float* __restrict__ shift, int64_t rows, int D,
                            int group_rows) {
  const int vec_per_row = D / 8;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * vec_per_row) return;
  const int64_t row = i / vec_per_row;
  const int c = static_cast<int>(i % vec_per_row) * 8;
  const int64_t f = row / group_rows;
  const float sc = scale[f], sh = shift[f];
  uint4 a = *reinterpret_cast<const uint4*>(kv + row * ld + c);
  const uint4 b = *reinterpret_cast<const uint4*>(h + row * D + c);
  uint32_t* aw = reinterpret_cast<uint32_t*>(&a);
  const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a0 = __uint_as_float(aw[j] << 16), a1 = __uint_as_float(aw[j] & 0xffff0000u);
    const float b0 = __uint_as_float(bw[j] << 16), b1 = __uint_as_float(bw[j] & 0xffff0000u);
    aw[j] = pack_bf16(a0 + fmaf(sc, b0, sh), a1 + fmaf(sc, b1, sh));
  }
  *reinterpret_cast<uint4*>(kv + row * ld + c) = a;
}

static size_t adapter_ws(int type, int D, int inner, int64_t rows) {
  const size_t r = static_cast<size_t>(rows);
  if (type == DFD_ADAPTER_LINEAR || type == DFD_ADAPTER_BN) return r * D * 2;
  return r * inner * 2 * (type == DFD_ADAPTER_XXX ? 2 : 1);
}

int adapter_apply(const dfd_ctx* ctx, int type, int D, int inner, const dfd_adapter_weights* w, void* kv, int64_t ld,
                  int64_t rows, int group_rows, int group_skip, void* workspace, size_t ws_bytes,
                  cudaStream_t stream) {
  DFD_CHECK_ARG(w && kv, "adapter: null pointer");
  DFD_CHECK_ARG(type >= DFD_ADAPTER_GELU_LN && type <= DFD_ADAPTER_BN, "adapter: unknown struct type %d", type);
  DFD_CHECK_ARG(D > 0 && D % 256 == 0, "adapter: width %d must be a multiple of 256", D);
  DFD_CHECK_ARG(rows >= 0 && rows < (1ll << 31) - 256, "adapter: row count out of range");
  if (rows == 0) return 0;
  if (type == DFD_ADAPTER_LINEAR || type == DFD_ADAPTER_BN) inner = D;
  DFD_CHECK_ARG(inner > 0 && inner % 256 == 0 && inner <= 1024, "adapter: inner width %d must be 256, 512, 768 or 1024",
                inner);
  const size_t need = adapter_ws(type, D, inner, rows);
  if (!workspace || ws_bytes < need)
    return fail(DFD_ERR_WORKSPACE, "adapter: workspace %zu < %zu bytes", ws_bytes, need);
  const int M = static_cast<int>(rows);
  DFD_CHECK_ARG(w->w_down != nullptr, "adapter: missing down-projection weight");
  ScopedTimer timer(ctx, DFD_TAG_ADAPTER, stream);

  if (type == DFD_ADAPTER_LINEAR) {
    // kvs[i][k] = Linear(D, D)(tap), no residual (src/models.py:900-917, 933): the GEMM cannot overwrite the slice
    // it is still reading, so it goes through the workspace.
    DFD_TRY(gemm_bf16(ctx, kv, ld, w->w_down, D, nullptr, workspace, D, M, D, D, DFD_EPI_STORE_BF16, stream));
    DFD_CUDA_OK(cudaMemcpy2DAsync(kv, static_cast<size_t>(ld) * 2, workspace, static_cast<size_t>(D) * 2,
                                  static_cast<size_t>(D) * 2, static_cast<size_t>(rows), cudaMemcpyDeviceToDevice,
                                  stream));
    return 0;
  }
  if (type == DFD_ADAPTER_BN) {
    // tap += BatchNorm2d_eval(Linear(D, D)(tap)) (src/models.py:877-887, 930-931): the normalisation is one scalar
    // affine per frame, applied while the product is added to the tap
    DFD_CHECK_ARG(w->ln_weight && w->ln_bias, "adapter: missing per-frame scale / shift");
    DFD_CHECK_ARG(group_rows > 0 && rows % group_rows == 0, "adapter: %lld rows are not whole frames of %d rows",
                  (long long)rows, group_rows);
    DFD_CHECK_ARG(ld % 8 == 0 && reinterpret_cast<uintptr_t>(kv) % 16 == 0, "adapter: tap must be 16-byte aligned");
    DFD_TRY(gemm_bf16(ctx, kv, ld, w->w_down, D, nullptr, workspace, D, M, D, D, DFD_EPI_STORE_BF16, stream));
    const int64_t vecs = rows * (D / 8);
    adapter_frame_affine_kernel<<<static_cast<unsigned>((vecs + 255) / 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(kv), ld, static_cast<const __nv_bfloat16*>(workspace), w->ln_weight, w->ln_bias,
        rows, D, group_rows);
    DFD_CUDA_OK(cudaGetLastError());
    return 0;
  }
  DFD_CHECK_ARG(w->w_up != nullptr, "adapter: missing up-projection weight");
  void* h = workspace;
  if (type == DFD_ADAPTER_XXX) {
    // Linear -> GELU -> Linear -> GELU -> Linear (src/models.py:881-899)
    DFD_CHECK_ARG(w->w_mid != nullptr, "adapter: missing middle weight");
    void* h2 = static_cast<uint8_t*>(workspace) + static_cast<size_t>(rows) * inner * 2;
    DFD_TRY(gemm_bf16(ctx, kv, ld, w->w_down, D, nullptr, h, inner, M, inner, D, DFD_EPI_STORE_BF16_GELU, stream));
    DFD_TRY(gemm_bf16(ctx, h, inner, w->w_mid, inner, nullptr, h2, inner, M, inner, inner, DFD_EPI_STORE_BF16_GELU,
                      stream));
    h = h2;
  } else {
    DFD_CHECK_ARG(w->ln_weight && w->ln_bias, "adapter: missing LayerNorm parameters");
    DFD_TRY(gemm_bf16(ctx, kv, ld, w->w_down, D, nullptr, h, inner, M, inner, D, DFD_EPI_STORE_BF16, stream));
    if (type == DFD_ADAPTER_GELU_LN) {
      DFD_TRY(launch_rownorm<true>(h, w->ln_weight, w->ln_bias, rows, inner, stream));
    } else if (type == DFD_ADAPTER_LN_GELU) {
      DFD_TRY(launch_rownorm<false>(h, w->ln_weight, w->ln_bias, rows, inner, stream));
    } else {
      DFD_CHECK_ARG(group_rows > 0 && group_skip >= 0 && group_skip < group_rows && rows % group_rows == 0,
                    "adapter: %lld rows are not whole frames of %d rows", (long long)rows, group_rows);
      adapter_groupnorm_kernel<<<static_cast<unsigned>(rows / group_rows), 1024, 0, stream>>>(
          static_cast<__nv_bfloat16*>(h), w->ln_weight, w->ln_bias, inner, group_rows, group_skip);
      DFD_CUDA_OK(cudaGetLastError());
    }
  }
  // tap += h . Wup^T (residual, src/models.py:930-931)
  DFD_TRY(gemm_bf16(ctx, h, inner, w->w_up, inner, nullptr, kv, ld, M, D, inner, DFD_EPI_ADD_BF16, stream));
  return 0;
}

}  // namespace dfd

extern "C" size_t dfd_adapter_workspace_bytes(int type, int D, int inner, int64_t rows) {
  if (rows <= 0 || D <= 0) return 0;
  return dfd::adapter_ws(type, D, (type == DFD_ADAPTER_LINEAR || type == DFD_ADAPTER_BN) ? D : inner, rows);
}

extern "C" int dfd_adapter_apply(dfd_ctx* ctx, int type, int D, int inner, const dfd_adapter_weights* w, void* kv,
                                 int64_t ld, int64_t rows, int group_rows, int group_skip, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_adapter_apply: ctx is NULL");
  return dfd::adapter_apply(ctx, type, D, inner, w, kv, ld, rows, group_rows, group_skip, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

// Decoder cross-attention for op_mode.attn_mode = "frame" / "temporal" / "temporal+frame" (src/models.py:107-115):
// the softmax term is normalised inside every frame (over its P patches) and/or across the T frames of every patch
// position instead of over all S = T*P keys, so a clip needs all its scores before any weight is known. Three passes,
// each touching K or V exactly once (same HBM bytes as the fused default-mode kernel):
//   1. dec_scores_kernel    s0 = q0.(K+pe)/8 (masked: -inf), a1 = tanh(q1.(K+pe)/8) * 2 sigmoid(-|q1-K-pe|_1/8)
//                           (masked: 0) for every key and head                                  — reads K
//   2. dec_weights_kernel   w = 1/2 (softmax_frame(s0) [+ softmax_temporal(s0)]) + 1/2 a1       — [B,S,H] floats only
//   3. dec_mix_kernel       mix = sum_s w (V + pe), per-frame partial sums, fixed-order reduce   — reads V
// A softmax over an all-masked group is NaN here exactly as in the reference (exp(-inf - -inf)).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

// one warp per key token; 8 lanes per head (8 channels = one 16-byte load each), 4 heads per pass
__global__ void __launch_bounds__(256)
dec_scores_kernel(const float* __restrict__ qs, const __nv_bfloat16* __restrict__ kbase, int64_t stride_b,
                  int64_t stride_t, int64_t stride_p, const float* __restrict__ pos_emb,
                  const uint8_t* __restrict__ mask, int B, int T, int P, int H, float* __restrict__ s0,
                  float* __restrict__ a1) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t S = static_cast<int64_t>(T) * P;
  if (tok >= B * S) return;
  const int b = static_cast<int>(tok / S), s = static_cast<int>(tok % S), t = s / P, p = s % P;
  const bool present = mask[b * T + t] != 0;
  const __nv_bfloat16* krow = kbase + b * stride_b + t * stride_t + p * stride_p;
  const int sub = lane & 7;  // channel octet of the head
  for (int h0 = 0; h0 < H; h0 += 4) {
    const int head = h0 + (lane >> 3);
    const uint4 raw = *reinterpret_cast<const uint4*>(krow + head * 64 + sub * 8);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    const float* q = qs + (static_cast<int64_t>(b) * H + head) * 128 + sub * 8;
    const float* pe = pos_emb ? pos_emb + (static_cast<int64_t>(t) * H + head) * 64 + sub * 8 : nullptr;
    float d0 = 0.f, d1 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float kv = __uint_as_float((i & 1) ? (w[i >> 1] & 0xffff0000u) : (w[i >> 1] << 16));
      if (pe) kv += pe[i];
      const float q0 = q[i], q1 = q[64 + i];
      d0 = fmaf(q0, kv, d0);
      d1 = fmaf(q1, kv, d1);
      l1 += fabsf(q1 - kv);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    if (sub == 0) {
      const int64_t o = tok * H + head;
      s0[o] = present ? d0 * 0.125f : -INFINITY;
      a1[o] = present ? tanhf(d1 * 0.125f) * (2.f / (1.f + __expf(l1 * 0.125f))) : 0.f;
    }
  }
}

// grid (B, H): all S scores of one (clip, head) in shared memory; w overwrites s0.
__global__ void __launch_bounds__(256)
dec_weights_kernel(float* __restrict__ s0, const float* __restrict__ a1, int T, int P, int H, int attn_mode) {
  extern __shared__ float sc[];  // [S] scores, then [S] accumulated softmax terms
  const int b = blockIdx.x, head = blockIdx.y, S = T * P;
  float* acc = sc + S;
  const int64_t base = static_cast<int64_t>(b) * S * H + head;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    sc[s] = s0[base + static_cast<int64_t>(s) * H];
    acc[s] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (attn_mode & DFD_ATTN_FRAME) {  // softmax over the P patches of each frame (dim=-2 of [n,q,T,P,h])
    for (int t = warp; t < T; t += nw) {
      const float* row = sc + t * P;
      float mx = -INFINITY;
      for (int p = lane; p < P; p += 32) mx = fmaxf(mx, row[p]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int p = lane; p < P; p += 32) sum += __expf(row[p] - mx);  // all -inf: exp(NaN) = NaN, as torch
      sum = warp_sum(sum);
      for (int p = lane; p < P; p += 32) acc[t * P + p] += __expf(row[p] - mx) / sum;
    }
    __syncthreads();
  }
  if (attn_mode & DFD_ATTN_TEMPORAL) {  // softmax over the T frames of each patch position (dim=-3)
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      float mx = -INFINITY;
      for (int t = 0; t < T; ++t) mx = fmaxf(mx, sc[t * P + p]);
      float sum = 0.f;
      for (int t = 0; t < T; ++t) sum += __expf(sc[t * P + p] - mx);
      for (int t = 0; t < T; ++t) acc[t * P + p] += __expf(sc[t * P + p] - mx) / sum;
    }
    __syncthreads();
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int64_t o = base + static_cast<int64_t>(s) * H;
    s0[o] = 0.5f * acc[s] + 0.5f * a1[o];  // aff = sum_i act_i / n_act, n_act = 2 (models.py:140-142)
  }
}

// grid (T, B), H*32 threads: thread = (head, channel pair); partial[b,t,h,c] = sum_p w[b,t,p,h] (V[b,t,p,h,c] + pe)
__global__ void __launch_bounds__(512)
dec_mix_kernel(const float* __restrict__ w, const __nv_bfloat16* __restrict__ vbase, int64_t stride_b,
               int64_t stride_t, int64_t stride_p, const float* __restrict__ pos_emb, int T, int P, int H,
               float* __restrict__ partial) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int head = threadIdx.x >> 5, c2 = threadIdx.x & 31;
  const __nv_bfloat16* vrow = vbase + b * stride_b + t * stride_t + head * 64 + 2 * c2;
  const float* wrow = w + (static_cast<int64_t>(b) * T + t) * P * H + head;
  float pe0 = 0.f, pe1 = 0.f;
  if (pos_emb) {
    pe0 = pos_emb[(static_cast<int64_t>(t) * H + head) * 64 + 2 * c2];
    pe1 = pos_emb[(static_cast<int64_t>(t) * H + head) * 64 + 2 * c2 + 1];
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
  for (int p = 0; p < P; ++p) {
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(vrow + p * stride_p);
    const float wt = wrow[static_cast<int64_t>(p) * H];
    a0 = fmaf(wt, __uint_as_float(raw << 16) + pe0, a0);
    a1 = fmaf(wt, __uint_as_float(raw & 0xffff0000u) + pe1, a1);
  }
  float* out = partial + ((static_cast<int64_t>(b) * T + t) * H + head) * 64 + 2 * c2;
  out[0] = a0;
  out[1] = a1;
}

// mix[b, h*64 + c] = sum_t partial[b,t,h,c] in frame order
__global__ void dec_mix_reduce_kernel(const float* __restrict__ partial, int T, int HD, float* __restrict__ mix,
                                      int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / HD, c = i % HD;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += partial[(b * T + t) * HD + c];
  mix[i] = s;
}

size_t dec_attn_modes_workspace_bytes(int B, int T, int P, int H) {
  const size_t keys = static_cast<size_t>(B) * T * P * H;
  return (2 * keys + static_cast<size_t>(B) * T * H * 64) * sizeof(float);
}

int decoder_attention_modes(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                            int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T,
                            int P, int H, int attn_mode, float* mix, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream) {
  DFD_CHECK_ARG(B >= 0 && T > 0 && P > 0, "decoder_attention_modes: bad shape B=%d T=%d P=%d", B, T, P);
  DFD_CHECK_ARG(attn_mode > 0 && attn_mode <= (DFD_ATTN_FRAME | DFD_ATTN_TEMPORAL),
                "decoder_attention_modes: attn_mode=%d", attn_mode);
  if (B == 0) return 0;
  DFD_CHECK_ARG(qs && k && v && mask && mix, "decoder_attention_modes: null pointer");
  DFD_CHECK_ARG(H % 4 == 0 && H >= 4 && H <= 16, "decoder_attention_modes: heads=%d unsupported (need 4,8,12,16)", H);
  DFD_CHECK_ARG(stride_p % 8 == 0 && stride_t % 8 == 0 && stride_b % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
                "decoder_attention_modes: K/V must be 16-byte aligned with strides that are multiples of 8 elements");
  const size_t need = dec_attn_modes_workspace_bytes(B, T, P, H);
  if (!workspace || workspace_bytes < need)
    return fail(DFD_ERR_WORKSPACE, "decoder_attention_modes: workspace %zu < %zu bytes", workspace_bytes, need);
  const int S = T * P;
  const size_t smem = static_cast<size_t>(2) * S * sizeof(float);
  DFD_CHECK_ARG(smem <= 200 * 1024, "decoder_attention_modes: %d keys per clip exceed the weight kernel's shared memory",
                S);
  const int64_t keys = static_cast<int64_t>(B) * S;
  float* s0 = static_cast<float*>(workspace);
  float* a1 = s0 + keys * H;
  float* partial = a1 + keys * H;
  dec_scores_kernel<<<static_cast<unsigned>((keys + 7) / 8), 256, 0, stream>>>(
      qs, static_cast<const __nv_bfloat16*>(k), stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, H, s0, a1);
  DFD_CUDA_OK(cudaGetLastError());
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(dec_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[ctx->device & 63] = true;
  }
  dec_weights_kernel<<<dim3(B, H), 256, smem, stream>>>(s0, a1, T, P, H, attn_mode);
  DFD_CUDA_OK(cudaGetLastError());
  dec_mix_kernel<<<dim3(T, B), H * 32, 0, stream>>>(s0, static_cast<const __nv_bfloat16*>(v), stride_b, stride_t,
                                                     stride_p, pos_emb, T, P, H, partial);
  DFD_CUDA_OK(cudaGetLastError());
  const int64_t total = static_cast<int64_t>(B) * H * 64;
  dec_mix_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(partial, T, H * 64, mix, total);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

extern "C" size_t dfd_decoder_attention_modes_workspace_bytes(int B, int T, int P, int H) {
  if (B <= 0 || T <= 0 || P <= 0 || H <= 0) return 0;
  return dfd::dec_attn_modes_workspace_bytes(B, T, P, H);
}

extern "C" int dfd_decoder_attention_modes(dfd_ctx* ctx, const float* qs, const void* k, const void* v,
                                           int64_t stride_b, int64_t stride_t, int64_t stride_p, const float* pos_emb,
                                           const uint8_t* mask, int B, int T, int P, int H, int attn_mode, float* mix,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_attention_modes: ctx is NULL");
  return dfd::decoder_attention_modes(ctx, qs, k, v, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, H,
                                      attn_mode, mix, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

// Backward of the decoder cross-attention (src/models.py:99-146 under autograd) for the training step (config C5):
// the encoder is frozen, so the gradients that always exist are those of the per-clip queries (-> in_proj) and of the
// temporal position embedding, which is added to K and to V (:326-329). With a trainable CompInvAdapter between the
// taps and the decoder (:546-547) the per-key gradients dK, dV are needed as well: they are the two summands of the
// position-embedding gradient and are written out on request.
//
// With g = dmix / 2, K~ = K + pe_t, V~ = V + pe_t, per key s and head h:
//   gv  = g . V~                      p^ = exp(q0.K~/8 - M) / L              (M, L, o0 = sum p^ V~ saved by the forward)
//   ds0 = p^ (gv - g.o0)              u = q1.K~/8,  th = tanh u,  y = |q1 - K~|_1 / 8,  G = 2 / (1 + e^y),  a1 = th G
//   du  = gv G (1 - th^2)             dy = -gv th G (1 - G/2)
//   dq0 += ds0 K~ / 8                 dq1 += (du K~ + dy sign(q1 - K~)) / 8
//   dpe_t += (ds0 q0 + du q1 - dy sign(q1 - K~)) / 8 + (p^ + a1) g          (d/dK~ + d/dV~)
// One more streaming pass over K and V (HBM-bound like the forward): one CTA per (clip, frame), lane = (head, 8
// channels), 8-lane shuffle reductions for the four per-key scalars, key subsets merged through shared memory;
// per-(clip, frame) partials are summed over frames (dq) and over clips (dpe) in a fixed order (deterministic).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

constexpr int DBW_REC = 192;  // per (clip, frame, head): dq0[64], dq1[64], dpe[64]

__device__ __forceinline__ void bw_unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i + 0] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int H>
__global__ void __launch_bounds__(384)
dec_attn_bwd_kernel(const float* __restrict__ qs, const __nv_bfloat16* __restrict__ kbase,
                    const __nv_bfloat16* __restrict__ vbase, int64_t stride_b, int64_t stride_t, int64_t stride_p,
                    const float* __restrict__ pos_emb, const uint8_t* __restrict__ mask,
                    const float* __restrict__ stats, const float* __restrict__ dmix, int T, int P,
                    float* __restrict__ part, float* __restrict__ dk_out, float* __restrict__ dv_out) {
  constexpr int HG = H / 4;
  constexpr int KS = (H == 4) ? 8 : (H == 8 ? 4 : (H == 12 ? 4 : 3));
  extern __shared__ float bsm[];  // [KS][H][DBW_REC]
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out = part + static_cast<int64_t>(blockIdx.x) * H * DBW_REC;
  // optional per-key gradients (trainable adapter): contiguous fp32 [B, T, P, H, 64]
  float* dkf = dk_out ? dk_out + static_cast<int64_t>(blockIdx.x) * P * H * 64 : nullptr;
  float* dvf = dv_out ? dv_out + static_cast<int64_t>(blockIdx.x) * P * H * 64 : nullptr;
  if (mask[b * T + t] == 0) {
    for (int i = threadIdx.x; i < H * DBW_REC; i += blockDim.x) out[i] = 0.f;
    if (dkf)  // keys of an absent frame get no gradient
      for (int i = threadIdx.x; i < P * H * 64; i += blockDim.x) dkf[i] = dvf[i] = 0.f;
    return;
  }
  const int hg = warp % HG, ks = warp / HG;
  const int head = hg * 4 + (lane >> 3), d0 = (lane & 7) * 8;

  float q0[8], q1[8], pe[8], g[8], dq0[8], dq1[8], dpe[8];
  float go = 0.f;
  const float* st = stats + (static_cast<int64_t>(b) * H + head) * 66;
  const float M = st[0], invL = 1.f / st[1];
  {
    const float* qh = qs + (static_cast<int64_t>(b) * H + head) * 128;
    const float* dm = dmix + (static_cast<int64_t>(b) * H + head) * 64;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      q0[e] = qh[d0 + e];
      q1[e] = qh[64 + d0 + e];
      pe[e] = pos_emb ? pos_emb[(static_cast<int64_t>(t) * H + head) * 64 + d0 + e] : 0.f;
      g[e] = 0.5f * dm[d0 + e];
      go = fmaf(g[e], st[2 + d0 + e], go);
      dq0[e] = dq1[e] = dpe[e] = 0.f;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) go += __shfl_xor_sync(0xffffffffu, go, o);
  }
  const int64_t off = b * stride_b + t * stride_t + hg * 256 + lane * 8;
  const __nv_bfloat16* kf = kbase + off;
  const __nv_bfloat16* vf = vbase + off;

  // Four keys of this warp's subset per trip, all eight 16-byte loads issued before the first one is used: with one
  // CTA per (clip, frame) — 96 CTAs at the training batch of 12 clips — a one-key-per-trip loop waits out a full HBM
  // round trip per key (54 us per launch for 58 MB). The keys are still consumed in increasing order, so the sums
  // are bit-identical to the one-key loop.
  constexpr int KU = 4;
  auto process = [&](const uint4& kraw, const uint4& vraw, int p) {
    float kt[8], vt[8];
    bw_unpack8(kraw, kt);
    bw_unpack8(vraw, vt);
    float d0s = 0.f, d1s = 0.f, l1s = 0.f, gv = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      kt[e] += pe[e];
      vt[e] += pe[e];
      d0s = fmaf(q0[e], kt[e], d0s);
      d1s = fmaf(q1[e], kt[e], d1s);
      l1s += fabsf(q1[e] - kt[e]);
      gv = fmaf(g[e], vt[e], gv);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      d0s += __shfl_xor_sync(0xffffffffu, d0s, o);
      d1s += __shfl_xor_sync(0xffffffffu, d1s, o);
      l1s += __shfl_xor_sync(0xffffffffu, l1s, o);
      gv += __shfl_xor_sync(0xffffffffu, gv, o);
    }
    const float ph = __expf(d0s * 0.125f - M) * invL;
    const float ds0 = ph * (gv - go);
    const float th = tanhf(d1s * 0.125f);
    const float G = 2.f / (1.f + __expf(l1s * 0.125f));
    const float du = gv * G * (1.f - th * th);
    const float dy = -gv * th * G * (1.f - 0.5f * G);
    const float wv = ph + th * G;
    float dkk[8], dvv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float diff = q1[e] - kt[e];
      const float sg = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
      dq0[e] = fmaf(ds0 * 0.125f, kt[e], dq0[e]);
      dq1[e] += 0.125f * (du * kt[e] + dy * sg);
      dkk[e] = 0.125f * (ds0 * q0[e] + du * q1[e] - dy * sg);  // d/dK~ of this key
      dvv[e] = wv * g[e];                                       // d/dV~ of this key
      dpe[e] += dkk[e] + dvv[e];
    }
    if (dkf) {
      float4* dk4 = reinterpret_cast<float4*>(dkf + (static_cast<int64_t>(p) * H + head) * 64 + d0);
      float4* dv4 = reinterpret_cast<float4*>(dvf + (static_cast<int64_t>(p) * H + head) * 64 + d0);
      dk4[0] = make_float4(dkk[0], dkk[1], dkk[2], dkk[3]);
      dk4[1] = make_float4(dkk[4], dkk[5], dkk[6], dkk[7]);
      dv4[0] = make_float4(dvv[0], dvv[1], dvv[2], dvv[3]);
      dv4[1] = make_float4(dvv[4], dvv[5], dvv[6], dvv[7]);
    }
  };
  for (int p0 = ks; p0 < P; p0 += KS * KU) {
    uint4 kr[KU], vr[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int p = p0 + u * KS;
      if (p < P) {
        kr[u] = *reinterpret_cast<const uint4*>(kf + p * stride_p);
        vr[u] = *reinterpret_cast<const uint4*>(vf + p * stride_p);
      }
    }
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int p = p0 + u * KS;
      if (p < P) process(kr[u], vr[u], p);
    }
  }
  // ---- merge the KS key subsets
  {
    float* rec = bsm + (static_cast<int64_t>(ks) * H + head) * DBW_REC;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      rec[d0 + e] = dq0[e];
      rec[64 + d0 + e] = dq1[e];
      rec[128 + d0 + e] = dpe[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * DBW_REC; i += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < KS; ++w) sacc += bsm[w * H * DBW_REC + i];
    out[i] = sacc;
  }
}

// dqs[b, h, 0:64] = sum_t dq0, dqs[b, h, 64:128] = sum_t dq1.   grid = B*H, 128 threads
__global__ void dec_attn_bwd_dq_kernel(const float* __restrict__ part, int T, int H, float* __restrict__ dqs) {
  const int b = blockIdx.x / H, head = blockIdx.x % H, d = threadIdx.x;
  float sacc = 0.f;
  for (int t = 0; t < T; ++t) sacc += part[((static_cast<int64_t>(b) * T + t) * H + head) * DBW_REC + d];
  dqs[(static_cast<int64_t>(b) * H + head) * 128 + d] = sacc;
}

// dpe[t, h, :] = sum_b dpe partial.   grid = T*H, 64 threads
__global__ void dec_attn_bwd_dpe_kernel(const float* __restrict__ part, int B, int T, int H, float* __restrict__ dpe) {
  const int t = blockIdx.x / H, head = blockIdx.x % H, d = threadIdx.x;
  float sacc = 0.f;
  for (int b = 0; b < B; ++b) sacc += part[((static_cast<int64_t>(b) * T + t) * H + head) * DBW_REC + 128 + d];
  dpe[(static_cast<int64_t>(t) * H + head) * 64 + d] = sacc;
}

size_t dec_attn_bwd_workspace_bytes(int B, int T, int H) {
  return static_cast<size_t>(B) * T * H * DBW_REC * sizeof(float);
}

int decoder_attention_backward(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                               int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask,
                               const float* stats, const float* dmix, int B, int T, int P, int H, float* dqs,
                               float* dpos_emb, float* dk, float* dv, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
  DFD_CHECK_ARG(B >= 0 && T > 0 && P > 0, "decoder_attention_backward: bad shape B=%d T=%d P=%d", B, T, P);
  if (B == 0) return 0;
  DFD_CHECK_ARG(qs && k && v && mask && stats && dmix && dqs, "decoder_attention_backward: null pointer");
  DFD_CHECK_ARG((pos_emb == nullptr) == (dpos_emb == nullptr),
                "decoder_attention_backward: pos_emb and dpos_emb must both be given or both be NULL");
  DFD_CHECK_ARG((dk == nullptr) == (dv == nullptr), "decoder_attention_backward: dk and dv go together");
  DFD_CHECK_ARG(dk == nullptr || (reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) % 16 == 0,
                "decoder_attention_backward: dk/dv must be 16-byte aligned");
  DFD_CHECK_ARG(H % 4 == 0 && H >= 4 && H <= 16, "decoder_attention_backward: heads=%d unsupported", H);
  DFD_CHECK_ARG(stride_p % 8 == 0 && stride_t % 8 == 0 && stride_b % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
                "decoder_attention_backward: K/V must be 16-byte aligned with strides that are multiples of 8");
  const size_t need = dec_attn_bwd_workspace_bytes(B, T, H);
  if (!workspace || workspace_bytes < need)
    return fail(DFD_ERR_WORKSPACE, "decoder_attention_backward: workspace %zu < %zu bytes", workspace_bytes, need);
  float* part = static_cast<float*>(workspace);
  const unsigned grid = static_cast<unsigned>(B) * T;
  const __nv_bfloat16* kb = static_cast<const __nv_bfloat16*>(k);
  const __nv_bfloat16* vb = static_cast<const __nv_bfloat16*>(v);
#define DFD_LAUNCH_BWD(HH, KSV)                                                                                  \
  do {                                                                                                           \
    const size_t smem = static_cast<size_t>(KSV) * HH * DBW_REC * sizeof(float);                                 \
    static std::atomic<bool> configured{false};                                                                        \
    if (!configured) {                                                                                           \
      DFD_CUDA_OK(cudaFuncSetAttribute(dec_attn_bwd_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       (int)smem));                                                              \
      configured = true;                                                                                         \
    }                                                                                                            \
    dec_attn_bwd_kernel<HH><<<grid, (HH / 4) * KSV * 32, smem, stream>>>(qs, kb, vb, stride_b, stride_t,         \
                                                                        stride_p, pos_emb, mask, stats, dmix, T, \
                                                                        P, part, dk, dv);                        \
  } while (0)
  switch (H) {
    case 4: DFD_LAUNCH_BWD(4, 8); break;
    case 8: DFD_LAUNCH_BWD(8, 4); break;
    case 12: DFD_LAUNCH_BWD(12, 4); break;
    default: DFD_LAUNCH_BWD(16, 3); break;
  }
#undef DFD_LAUNCH_BWD
  DFD_CUDA_OK(cudaGetLastError());
  dec_attn_bwd_dq_kernel<<<B * H, 128, 0, stream>>>(part, T, H, dqs);
  DFD_CUDA_OK(cudaGetLastError());
  if (dpos_emb) {
    dec_attn_bwd_dpe_kernel<<<T * H, 64, 0, stream>>>(part, B, T, H, dpos_emb);
    DFD_CUDA_OK(cudaGetLastError());
  }
  (void)ctx;
  return 0;
}

int decoder_attention(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                      int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T, int P,
                      int H, float* mix, void* workspace, size_t workspace_bytes, cudaStream_t stream, float* stats);

}  // namespace dfd

extern "C" {

size_t dfd_decoder_attention_workspace_bytes(int B, int T, int H) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  const size_t fwd = static_cast<size_t>(B) * T * 2 * H * 130 * sizeof(float);
  const size_t bwd = dfd::dec_attn_bwd_workspace_bytes(B, T, H);
  return fwd > bwd ? fwd : bwd;
}

int dfd_decoder_attention_train(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                                int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                                int T, int P, int H, float* mix, float* stats, void* workspace, size_t workspace_bytes,
                                void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_attention_train: ctx is NULL");
  if (!stats && B > 0) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_attention_train: stats is NULL");
  return dfd::decoder_attention(ctx, qs, k, v, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, H, mix, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream), stats);
}

int dfd_decoder_attention_backward(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                                   int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask,
                                   const float* stats, const float* dmix, int B, int T, int P, int H, float* dqs,
                                   float* dpos_emb, float* dk, float* dv, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_attention_backward: ctx is NULL");
  return dfd::decoder_attention_backward(ctx, qs, k, v, stride_b, stride_t, stride_p, pos_emb, mask, stats, dmix, B, T,
                                         P, H, dqs, dpos_emb, dk, dv, workspace, workspace_bytes,
                                         static_cast<cudaStream_t>(stream));
}

}  // extern "C"

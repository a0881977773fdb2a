// Temporal decoder / classification head (src/models.py:81-146, 149-176, 259-269, 323-361) in fp32.
//
// The only heavy part is the cross-attention of ONE learnable query per clip against the tapped encoder K/V of
// all T*P patch tokens (bf16, read in place from the encoder's QKV buffers through strides, never copied):
//   K~ = K + pe[t], V~ = V + pe[t]                                            (models.py:326-329)
//   a0 = softmax_s( q0.K~/8  masked -inf )                                    (smax, :99-106)
//   a1 = tanh(q1.K~/8) * 2*sigmoid(-|q1 - K~|_1 / 8)   masked 0               (coda, :117-125)
//   mix = sum_s 0.5*(a0 + a1) * V~                                            (:142-144)
// That is one streaming pass over K and V (HBM-bound: 2*S*D*2 bytes per clip and block), done by
// dec_attn_partial_kernel (one CTA per (clip, frame), online softmax, 8-lane shuffle reductions) and a tiny
// cross-frame combine. The 1-token-per-clip linear layers are small fp32 GEMMs (weights-bandwidth bound).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

int layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int pos_period, void* out_bf16,
              float* out_f32, int64_t rows, int D, cudaStream_t stream);

// ------------------------------------------------------------------------------ decoder attention (partial)
constexpr int DEC_WARPS = 8;
constexpr int DEC_REC = 130;  // per (clip, frame, head): m, l, acc0[64], acc1[64]

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i + 0] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// NCH = H/4: 16-byte chunks of one K (or V) row handled by each lane; lane l, slot j covers head 4j + l/8,
// channels (l%8)*8 .. +8.
template <int NCH>
__global__ void __launch_bounds__(DEC_WARPS * 32, 1)
dec_attn_partial_kernel(const float* __restrict__ qs, const __nv_bfloat16* __restrict__ kbase,
                        const __nv_bfloat16* __restrict__ vbase, int64_t stride_b, int64_t stride_t, int64_t stride_p,
                        const float* __restrict__ pos_emb, const uint8_t* __restrict__ mask, int T, int P,
                        float* __restrict__ part) {
  constexpr int H = NCH * 4;
  extern __shared__ float dsm[];  // [DEC_WARPS][H][DEC_REC]
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out = part + (static_cast<int64_t>(blockIdx.x) * H) * DEC_REC;

  if (mask[b * T + t] == 0) {
    // frame absent: neutral element of the combine (m = -inf, l = 0, acc = 0)
    for (int i = threadIdx.x; i < H * DEC_REC; i += blockDim.x) out[i] = (i % DEC_REC == 0) ? -INFINITY : 0.f;
    return;
  }

  float q0[NCH][8], q1[NCH][8], pe[NCH][8], acc0[NCH][8], acc1[NCH][8], m[NCH], l[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int head = j * 4 + (lane >> 3), d0 = (lane & 7) * 8;
    const float* qh = qs + (static_cast<int64_t>(b) * H + head) * 128;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      q0[j][e] = qh[d0 + e];
      q1[j][e] = qh[64 + d0 + e];
      pe[j][e] = pos_emb ? pos_emb[(static_cast<int64_t>(t) * H + head) * 64 + d0 + e] : 0.f;
      acc0[j][e] = 0.f;
      acc1[j][e] = 0.f;
    }
    m[j] = -INFINITY;
    l[j] = 0.f;
  }

  const __nv_bfloat16* kf = kbase + b * stride_b + t * stride_t + lane * 8;
  const __nv_bfloat16* vf = vbase + b * stride_b + t * stride_t + lane * 8;

  constexpr int KB = 2;  // keys in flight per warp iteration
  for (int p0 = warp * KB; p0 < P; p0 += DEC_WARPS * KB) {
    uint4 kraw[KB][NCH], vraw[KB][NCH];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int p = min(p0 + u, P - 1);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        kraw[u][j] = ldg_stream16(kf + p * stride_p + j * 256);
        vraw[u][j] = ldg_stream16(vf + p * stride_p + j * 256);
      }
    }
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      if (p0 + u < P) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          float kk[8], vv[8];
          bf16x8_to_float(kraw[u][j], kk);
          bf16x8_to_float(vraw[u][j], vv);
          float d0 = 0.f, d1 = 0.f, l1 = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float kt = kk[e] + pe[j][e];
            d0 = fmaf(q0[j][e], kt, d0);
            d1 = fmaf(q1[j][e], kt, d1);
            l1 += fabsf(q1[j][e] - kt);
            vv[e] += pe[j][e];
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o);
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
          }
          const float s0 = d0 * 0.125f;
          const float mn = fmaxf(m[j], s0);
          const float resc = __expf(m[j] - mn);  // first key: exp(-inf) = 0
          const float pr = __expf(s0 - mn);
          m[j] = mn;
          l[j] = l[j] * resc + pr;
          // tanh(x) = 1 - 2/(1+exp(2x));  2*sigmoid(-y) = 2/(1+exp(y))
          const float th = 1.f - 2.f / (1.f + __expf(0.25f * d1));
          const float a1 = th * (2.f / (1.f + __expf(0.125f * l1)));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc0[j][e] = fmaf(acc0[j][e], resc, pr * vv[e]);
            acc1[j][e] = fmaf(a1, vv[e], acc1[j][e]);
          }
        }
      }
    }
  }

  // ---- combine the 8 warps of this CTA
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int head = j * 4 + (lane >> 3), d0 = (lane & 7) * 8;
    float* rec = dsm + (static_cast<int64_t>(warp) * H + head) * DEC_REC;
    if ((lane & 7) == 0) {
      rec[0] = m[j];
      rec[1] = l[j];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      rec[2 + d0 + e] = acc0[j][e];
      rec[66 + d0 + e] = acc1[j][e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * 64; i += blockDim.x) {
    const int head = i >> 6, d = i & 63;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) M = fmaxf(M, dsm[(w * H + head) * DEC_REC]);
    float Ls = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float* rec = dsm + (w * H + head) * DEC_REC;
      const float sc = (rec[0] == -INFINITY) ? 0.f : __expf(rec[0] - M);
      Ls += rec[1] * sc;
      a0 += rec[2 + d] * sc;
      a1 += rec[66 + d];
    }
    float* o = out + head * DEC_REC;
    if (d == 0) {
      o[0] = M;
      o[1] = Ls;
    }
    o[2 + d] = a0;
    o[66 + d] = a1;
  }
}

// mix[b, h*64 + d] = 0.5 * acc0/l + 0.5 * acc1 over the T frame partials.  grid = B*H, 64 threads.
__global__ void dec_attn_combine_kernel(const float* __restrict__ part, int T, int H, float* __restrict__ mix) {
  const int b = blockIdx.x / H, head = blockIdx.x % H, d = threadIdx.x;
  float M = -INFINITY;
  for (int t = 0; t < T; ++t) M = fmaxf(M, part[((static_cast<int64_t>(b) * T + t) * H + head) * DEC_REC]);
  float Ls = 0.f, a0 = 0.f, a1 = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* rec = part + ((static_cast<int64_t>(b) * T + t) * H + head) * DEC_REC;
    // all frames masked: M = -inf and exp(-inf - -inf) = NaN, as softmax over an all -inf row gives in the
    // reference (SURVEY 8c caveat 4)
    const float sc = __expf(rec[0] - M);
    Ls += rec[1] * sc;
    a0 += rec[2 + d] * sc;
    a1 += rec[66 + d];
  }
  mix[(static_cast<int64_t>(b) * H + head) * 64 + d] = 0.5f * (a0 / Ls) + 0.5f * a1;
}

size_t dec_attn_workspace_bytes(int B, int T, int H) {
  return static_cast<size_t>(B) * T * H * DEC_REC * sizeof(float);
}

int decoder_attention(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                      int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T, int P,
                      int H, float* mix, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DFD_CHECK_ARG(B >= 0 && T > 0 && P > 0, "decoder_attention: bad shape B=%d T=%d P=%d", B, T, P);
  if (B == 0) return 0;
  DFD_CHECK_ARG(qs && k && v && mask && mix, "decoder_attention: null pointer");
  DFD_CHECK_ARG(H % 4 == 0 && H >= 4 && H <= 16, "decoder_attention: heads=%d unsupported (need 4,8,12,16)", H);
  DFD_CHECK_ARG(stride_p % 8 == 0 && stride_t % 8 == 0 && stride_b % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
                "decoder_attention: K/V must be 16-byte aligned with strides that are multiples of 8 elements");
  if (B == 0) return 0;
  const size_t need = dec_attn_workspace_bytes(B, T, H);
  if (!workspace || workspace_bytes < need)
    return fail(DFD_ERR_WORKSPACE, "decoder_attention: workspace %zu < %zu bytes", workspace_bytes, need);
  float* part = static_cast<float*>(workspace);
  const size_t smem = static_cast<size_t>(DEC_WARPS) * H * DEC_REC * sizeof(float);
  const unsigned grid = static_cast<unsigned>(B) * T;
  const __nv_bfloat16* kb = static_cast<const __nv_bfloat16*>(k);
  const __nv_bfloat16* vb = static_cast<const __nv_bfloat16*>(v);
#define DFD_LAUNCH_DEC(NCH)                                                                                        \
  do {                                                                                                             \
    DFD_CUDA_OK(cudaFuncSetAttribute(dec_attn_partial_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)smem));                                                                  \
    dec_attn_partial_kernel<NCH><<<grid, DEC_WARPS * 32, smem, stream>>>(qs, kb, vb, stride_b, stride_t, stride_p, \
                                                                         pos_emb, mask, T, P, part);               \
  } while (0)
  switch (H / 4) {
    case 1: DFD_LAUNCH_DEC(1); break;
    case 2: DFD_LAUNCH_DEC(2); break;
    case 3: DFD_LAUNCH_DEC(3); break;
    default: DFD_LAUNCH_DEC(4); break;
  }
#undef DFD_LAUNCH_DEC
  DFD_CUDA_OK(cudaGetLastError());
  dec_attn_combine_kernel<<<B * H, 64, 0, stream>>>(part, T, H, mix);
  DFD_CUDA_OK(cudaGetLastError());
  (void)ctx;
  return 0;
}

// ------------------------------------------------------------------------------------- small fp32 linear
// out[b, n] = act( sum_k x[b,k] * W[n,k] + bias[n] ) (+ res[b,n]).  M = B is one token per clip, so these are
// weights-bandwidth bound: each CTA streams a [16 x K] slab of W once for up to 64 rows of x.
constexpr int LIN_BM = 64, LIN_BN = 16, LIN_BK = 64, LIN_PAD = 4;

template <bool QGELU>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                  const float* res, float* out, int B, int N, int K) {
  __shared__ __align__(16) float sx[LIN_BM][LIN_BK + LIN_PAD];
  __shared__ __align__(16) float sw[LIN_BN][LIN_BK + LIN_PAD];
  const int n0 = blockIdx.x * LIN_BN, b0 = blockIdx.y * LIN_BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx: column, ty: 4 rows
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += LIN_BK) {
    // x tile: 64 rows x 16 float4
    for (int i = threadIdx.x; i < LIN_BM * (LIN_BK / 4); i += 256) {
      const int r = i / (LIN_BK / 4), c = (i % (LIN_BK / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < B && k0 + c < K) v = *reinterpret_cast<const float4*>(x + static_cast<int64_t>(b0 + r) * K + k0 + c);
      *reinterpret_cast<float4*>(&sx[r][c]) = v;
    }
    {
      const int r = threadIdx.x / (LIN_BK / 4), c = (threadIdx.x % (LIN_BK / 4)) * 4;  // 16 rows x 16 float4
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < N && k0 + c < K) v = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(n0 + r) * K + k0 + c));
      *reinterpret_cast<float4*>(&sw[r][c]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LIN_BK; k += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&sw[tx][k]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(&sx[ty * 4 + r][k]);
        acc[r] = fmaf(xv.x, w.x, acc[r]);
        acc[r] = fmaf(xv.y, w.y, acc[r]);
        acc[r] = fmaf(xv.z, w.z, acc[r]);
        acc[r] = fmaf(xv.w, w.w, acc[r]);
      }
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
    const float bn = bias ? bias[n] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + ty * 4 + r;
      if (b < B) {
        float v = acc[r] + bn;
        if (QGELU) v = v / (1.f + __expf(-1.702f * v));
        if (res) v += res[static_cast<int64_t>(b) * N + n];
        out[static_cast<int64_t>(b) * N + n] = v;
      }
    }
  }
}

int linear_f32(const float* x, const float* W, const float* bias, const float* res, float* out, int B, int N, int K,
               bool qgelu, cudaStream_t stream) {
  DFD_CHECK_ARG(K % 4 == 0, "linear_f32: K=%d must be a multiple of 4", K);
  dim3 grid((N + LIN_BN - 1) / LIN_BN, (B + LIN_BM - 1) / LIN_BM);
  if (qgelu)
    linear_f32_kernel<true><<<grid, 256, 0, stream>>>(x, W, bias, res, out, B, N, K);
  else
    linear_f32_kernel<false><<<grid, 256, 0, stream>>>(x, W, bias, res, out, B, N, K);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// rows b of dst [B, ld] <- src[D]
__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) dst[i] = src[i % D];
}

// dst[b, slot, :] = src[b, :]
__global__ void scatter_block_out_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int D, int slot,
                                         int n_slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) dst[(static_cast<int64_t>(i / D) * n_slots + slot) * D + (i % D)] = src[i];
}

// ------------------------------------------------------------------------------- projection + logit scaling
// One CTA per clip: l = feature[b] @ proj[D,O]; logits = scale * l / (||l||_2 + 1e-10)  (models.py:359, 551-553)
__global__ void __launch_bounds__(256)
project_logits_kernel(const float* __restrict__ feature, const float* __restrict__ proj, int D, int O, float scale,
                      float* __restrict__ logits) {
  extern __shared__ float sl[];  // [O] raw logits, then [8] reduction scratch
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* f = feature + static_cast<int64_t>(b) * D;
  for (int o = warp; o < O; o += 8) {
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s = fmaf(f[k], proj[static_cast<int64_t>(k) * O + o], s);
    s = warp_sum(s);
    if (lane == 0) sl[o] = s;
  }
  __syncthreads();
  float ss = 0.f;
  for (int o = threadIdx.x; o < O; o += blockDim.x) ss += sl[o] * sl[o];
  ss = warp_sum(ss);
  float* red = sl + O;
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float mul = scale > 0.f ? scale / (sqrtf(tot) + 1e-10f) : 1.f;
  for (int o = threadIdx.x; o < O; o += blockDim.x) logits[static_cast<int64_t>(b) * O + o] = sl[o] * mul;
}

int project_logits(const float* feature, const float* proj, int B, int D, int O, float scale, float* logits,
                   cudaStream_t stream) {
  DFD_CHECK_ARG(B >= 0 && D > 0 && O > 0 && O <= 8192, "project_logits: bad shape B=%d D=%d O=%d", B, D, O);
  if (B == 0) return 0;
  DFD_CHECK_ARG(feature && proj && logits, "project_logits: null pointer");
  project_logits_kernel<<<B, 256, (O + 8) * sizeof(float), stream>>>(feature, proj, D, O, scale, logits);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------- whole decoder
struct DecWs {
  float *x, *y, *qs, *mix, *hid, *part;
  size_t total;
};

static DecWs carve_decoder_ws(void* base, int B, int T, int D, int H) {
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  DecWs w;
  auto take = [&](size_t bytes) {
    float* r = reinterpret_cast<float*>(p + off);
    off += up(bytes);
    return r;
  };
  w.x = take(sizeof(float) * B * D);
  w.y = take(sizeof(float) * B * D);
  w.qs = take(sizeof(float) * B * 2 * D);
  w.mix = take(sizeof(float) * B * D);
  w.hid = take(sizeof(float) * B * 4 * D);
  w.part = take(dec_attn_workspace_bytes(B, T, H));
  w.total = off;
  return w;
}

int decoder_forward(const dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                    const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                    float* video_feature, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DFD_CHECK_ARG(D == 64 * H, "decoder_forward: width %d != 64 * heads %d", D, H);
  DFD_CHECK_ARG(n_blocks > 0 && B >= 0 && T > 0 && P > 0, "decoder_forward: bad shape");
  if (B == 0) return 0;  // empty batch: nothing to do (buffers of empty tensors are NULL)
  DFD_CHECK_ARG(w && taps && mask && block_out && video_feature, "decoder_forward: null pointer");
  DecWs ws = carve_decoder_ws(workspace, B, T, D, H);
  if (!workspace || workspace_bytes < ws.total)
    return fail(DFD_ERR_WORKSPACE, "decoder_forward: workspace %zu < %zu bytes", workspace_bytes, ws.total);
  const int nthr = 256, nblk = (B * D + nthr - 1) / nthr;
#define DFD_TIMED(tag, call)          \
  do {                                \
    ScopedTimer _t(ctx, tag, stream); \
    DFD_TRY(call);                    \
  } while (0)

  // x = ln_pre(class_embedding) repeated for every clip (models.py:336-337; dropout p = 0)
  DFD_TRY(layernorm(w->class_embedding, w->ln_pre_weight, w->ln_pre_bias, nullptr, 0, nullptr, ws.y, 1, D, stream));
  broadcast_rows_kernel<<<nblk, nthr, 0, stream>>>(ws.y, ws.x, B, D);
  DFD_CUDA_OK(cudaGetLastError());

  for (int i = 0; i < n_blocks; ++i) {
    // x = x + out_proj(attn(in_proj(ln_1(x)), K_i, V_i, m))        (models.py:173-174, 136-146)
    DFD_TIMED(DFD_TAG_DEC_OTHER,
              layernorm(ws.x, w->ln_1_weight[i], w->ln_1_bias[i], nullptr, 0, nullptr, ws.y, B, D, stream));
    DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ws.y, w->in_proj_weight[i], w->in_proj_bias[i], nullptr, ws.qs, B, 2 * D,
                                             D, false, stream));
    DFD_TIMED(DFD_TAG_DEC_ATTN,
              decoder_attention(ctx, ws.qs, taps->k[i], taps->v[i], taps->stride_b, taps->stride_t, taps->stride_p,
                                w->positional_embedding, mask, B, T, P, H, ws.mix, ws.part,
                                dec_attn_workspace_bytes(B, T, H), stream));
    DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ws.mix, w->out_proj_weight[i], w->out_proj_bias[i], ws.x, ws.x, B, D, D,
                                             false, stream));
    // x = x + c_proj(quickgelu(c_fc(ln_2(x))))                     (models.py:175)
    DFD_TIMED(DFD_TAG_DEC_OTHER,
              layernorm(ws.x, w->ln_2_weight[i], w->ln_2_bias[i], nullptr, 0, nullptr, ws.y, B, D, stream));
    DFD_TIMED(DFD_TAG_DEC_LINEAR,
              linear_f32(ws.y, w->c_fc_weight[i], w->c_fc_bias[i], nullptr, ws.hid, B, 4 * D, D, true, stream));
    DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ws.hid, w->c_proj_weight[i], w->c_proj_bias[i], ws.x, ws.x, B, D, 4 * D,
                                             false, stream));
    scatter_block_out_kernel<<<nblk, nthr, 0, stream>>>(ws.x, block_out, B, D, i, n_blocks);
    DFD_CUDA_OK(cudaGetLastError());
  }
  // video_feature = ln_post(x_last)                                  (models.py:340-343)
  DFD_TRY(layernorm(ws.x, w->ln_post_weight, w->ln_post_bias, nullptr, 0, nullptr, video_feature, B, D, stream));
  return 0;
}

}  // namespace dfd

extern "C" {

size_t dfd_decoder_workspace_bytes(int B, int T, int D, int n_blocks) {
  (void)n_blocks;
  if (B <= 0 || T <= 0 || D <= 0) return 0;
  return dfd::carve_decoder_ws(nullptr, B, T, D, D / 64).total;
}

int dfd_decoder_forward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                        const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                        float* video_feature, void* workspace, size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_forward: ctx is NULL");
  return dfd::decoder_forward(ctx, D, H, n_blocks, w, taps, mask, B, T, P, block_out, video_feature, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dfd_project_logits(dfd_ctx* ctx, const float* feature, const float* proj, int B, int D, int O, float scale,
                       float* logits, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_project_logits: ctx is NULL");
  return dfd::project_logits(feature, proj, B, D, O, scale, logits, static_cast<cudaStream_t>(stream));
}

int dfd_decoder_attention(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                          int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T,
                          int P, int H, float* mix, void* workspace, size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_attention: ctx is NULL");
  return dfd::decoder_attention(ctx, qs, k, v, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, H, mix, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Temporal decoder / classification head (src/models.py:81-146, 149-176, 259-269, 323-361) in fp32.
//
// The only heavy part is the cross-attention of ONE learnable query per clip against the tapped encoder K/V of
// all T*P patch tokens (bf16, read in place from the encoder's QKV buffers through strides, never copied):
//   K~ = K + pe[t], V~ = V + pe[t]                                            (models.py:326-329)
//   a0 = softmax_s( q0.K~/8  masked -inf )                                    (smax, :99-106)
//   a1 = tanh(q1.K~/8) * 2*sigmoid(-|q1 - K~|_1 / 8)   masked 0               (coda, :117-125)
//   mix = sum_s 0.5*(a0 + a1) * V~                                            (:142-144)
// That is one streaming pass over K and V (HBM-bound: 2*S*D*2 bytes per clip and block), done by
// dec_attn_stream_kernel (decoder_attn_sm100.cu: persistent producer/consumer pipeline over a bulk-copy ring) and a
// tiny cross-unit combine. The 1-token-per-clip linear layers are small fp32 GEMMs (weights-bandwidth bound).
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace dfd {

int layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int pos_period, void* out_bf16,
              float* out_f32, int64_t rows, int D, cudaStream_t stream, void* fold_bf16 = nullptr,
              float* fold_stats = nullptr, int slots = 0);

// ------------------------------------------------------------------------------ decoder attention (partial)
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

// mix[b, h*64 + d] = 0.5 * acc0/l + 0.5 * acc1 over the T frame partials.  grid = B*H, 64 threads.
// stats (optional, for the backward pass): [B, H, 66] = (M, L, o0[64]) with o0 = softmax-weighted mean of V~.
__global__ void dec_attn_combine_kernel(const float* __restrict__ part, int T, int H, float* __restrict__ mix,
                                        float* __restrict__ stats = nullptr) {
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
  if (stats) {
    float* st = stats + (static_cast<int64_t>(b) * H + head) * 66;
    if (d == 0) {
      st[0] = M;
      st[1] = Ls;
    }
    st[2 + d] = a0 / Ls;
  }
}

size_t dec_attn_stream_workspace_bytes(int B, int T, int H);
int decoder_attention_stream(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                             int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                             int T, int P, int H, float* part, int* recs_per_clip, cudaStream_t stream);

size_t dec_attn_workspace_bytes(int B, int T, int H) { return dec_attn_stream_workspace_bytes(B, T, H); }

size_t dec_attn_modes_workspace_bytes(int B, int T, int P, int H);
int decoder_attention_modes(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                            int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T,
                            int P, int H, int attn_mode, float* mix, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream);

int decoder_attention(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                      int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T, int P,
                      int H, float* mix, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                      float* stats = nullptr) {
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
  int recs = 0;
  DFD_TRY(decoder_attention_stream(ctx, qs, k, v, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, H, part, &recs,
                                   stream));
  dec_attn_combine_kernel<<<B * H, 64, 0, stream>>>(part, recs, H, mix, stats);
  DFD_CUDA_OK(cudaGetLastError());
  (void)ctx;
  return 0;
}

// ------------------------------------------------------------------------------------- small fp32 linear
// out[b, n] = act( sum_k x[b,k] * W[n,k] + bias[n] ) (+ res[b,n]).  M = B is one token per clip, so these layers
// are weights-bandwidth bound (6.5 M fp32 parameters per block, each read once). To put every SM on the weight
// stream the K range is split across CTAs: CTA (n-tile of 16 columns, k-split, 64-row b-tile) accumulates a partial
// [64 x 16] tile with a cp.async double-buffered smem pipeline; partials are reduced in a fixed order by
// linear_reduce_kernel (deterministic, no atomics), which also applies bias / QuickGELU / residual.
// Measured and rejected (round 1, profiles/r1d_rejected.md): ONE kernel per linear with 64 x 64 tiles, 8 x 4 register
// tiles and the split-K reduction done by the last-arriving CTA of a tile (fence + counter) — with 144 CTAs of four
// warps there is nothing to hide the chain load -> FMA -> store -> fence -> reduce behind: 15-29 us per linear in
// three variants (row-major smem, k-major smem, whole k-range prefetched) against 13 + 5 us for this pair, whose
// 384 small CTAs (5 per SM) hide the latencies by occupancy.
constexpr int LIN_BM = 64, LIN_BN = 16, LIN_BK = 64, LIN_LD = LIN_BK + 4, LIN_THREADS = 128;

__device__ __forceinline__ void cp_async_f4(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}

__global__ void __launch_bounds__(LIN_THREADS)
linear_partial_kernel(const float* __restrict__ x, const float* __restrict__ W, float* __restrict__ part, int B, int N,
                      int K, int k_per_split) {
  __shared__ __align__(16) float sx[2][LIN_BM][LIN_LD];
  __shared__ __align__(16) float sw[2][LIN_BN][LIN_LD];
  const int n0 = blockIdx.x * LIN_BN, b0 = blockIdx.z * LIN_BM;
  const int kbeg = blockIdx.y * k_per_split, kend = min(K, kbeg + k_per_split);
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;  // tx: 2 columns (tx, tx+8); ty: 4 rows (ty, ty+16, ty+32, ty+48)

  auto load_stage = [&](int stage, int k0) {
    // x tile: 64 rows x 16 float4 = 1024 float4 -> 8 per thread; W tile: 16 x 16 = 256 -> 2 per thread
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * LIN_THREADS;
      const int r = idx >> 4, c = (idx & 15) * 4;
      const bool ok = (b0 + r < B) && (k0 + c < kend);
      cp_async_f4(&sx[stage][r][c], x + static_cast<int64_t>(ok ? b0 + r : 0) * K + (ok ? k0 + c : 0), ok);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * LIN_THREADS;
      const int r = idx >> 4, c = (idx & 15) * 4;
      const bool ok = (n0 + r < N) && (k0 + c < kend);
      cp_async_f4(&sw[stage][r][c], W + static_cast<int64_t>(ok ? n0 + r : 0) * K + (ok ? k0 + c : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
  const int nsteps = (kend - kbeg + LIN_BK - 1) / LIN_BK;
  if (nsteps > 0) load_stage(0, kbeg);
  for (int s = 0; s < nsteps; ++s) {
    if (s + 1 < nsteps) {
      load_stage((s + 1) & 1, kbeg + (s + 1) * LIN_BK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int st = s & 1;
#pragma unroll
    for (int k = 0; k < LIN_BK; k += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(&sw[st][tx][k]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sw[st][tx + 8][k]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(&sx[st][ty + 16 * r][k]);
        acc[r][0] = fmaf(xv.x, w0.x, acc[r][0]);
        acc[r][0] = fmaf(xv.y, w0.y, acc[r][0]);
        acc[r][0] = fmaf(xv.z, w0.z, acc[r][0]);
        acc[r][0] = fmaf(xv.w, w0.w, acc[r][0]);
        acc[r][1] = fmaf(xv.x, w1.x, acc[r][1]);
        acc[r][1] = fmaf(xv.y, w1.y, acc[r][1]);
        acc[r][1] = fmaf(xv.z, w1.z, acc[r][1]);
        acc[r][1] = fmaf(xv.w, w1.w, acc[r][1]);
      }
    }
    __syncthreads();
  }
  float* dst = part + static_cast<int64_t>(blockIdx.y) * B * N;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = b0 + ty + 16 * r;
    if (b < B) {
      if (n0 + tx < N) dst[static_cast<int64_t>(b) * N + n0 + tx] = acc[r][0];
      if (n0 + tx + 8 < N) dst[static_cast<int64_t>(b) * N + n0 + tx + 8] = acc[r][1];
    }
  }
}

template <bool QGELU>
__global__ void linear_reduce_kernel(const float* __restrict__ part, int splits, const float* __restrict__ bias,
                                     const float* res, float* out, int B, int N, float* pre_out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * N;
  if (i >= total) return;
  float v = bias ? bias[i % N] : 0.f;
  for (int s = 0; s < splits; ++s) v += part[s * total + i];
  if (pre_out) pre_out[i] = v;  // the pre-activation, saved for the training step's backward
  if (QGELU) v = v / (1.f + __expf(-1.702f * v));
  if (res) v += res[i];
  out[i] = v;
}

constexpr int LIN_MAX_SPLITS = 8;

size_t linear_workspace_bytes(int B, int max_n) {
  return static_cast<size_t>(LIN_MAX_SPLITS) * B * max_n * sizeof(float);
}

int linear_f32(const dfd_ctx* ctx, const float* x, const float* W, const float* bias, const float* res, float* out,
               int B, int N, int K, bool qgelu, float* part, cudaStream_t stream, float* pre_out = nullptr) {
  DFD_CHECK_ARG(K % 4 == 0, "linear_f32: K=%d must be a multiple of 4", K);
  const int n_tiles = (N + LIN_BN - 1) / LIN_BN, b_tiles = (B + LIN_BM - 1) / LIN_BM;
  // enough k-splits to cover the SMs about twice, each at least one BK step, at most LIN_MAX_SPLITS
  int splits = (2 * ctx->num_sms + n_tiles * b_tiles - 1) / (n_tiles * b_tiles);
  const int max_by_k = (K + LIN_BK - 1) / LIN_BK;
  if (splits > max_by_k) splits = max_by_k;
  if (splits > LIN_MAX_SPLITS) splits = LIN_MAX_SPLITS;
  if (splits < 1) splits = 1;
  int k_per_split = (K + splits - 1) / splits;
  k_per_split = (k_per_split + LIN_BK - 1) / LIN_BK * LIN_BK;
  splits = (K + k_per_split - 1) / k_per_split;
  dim3 grid(n_tiles, splits, b_tiles);
  linear_partial_kernel<<<grid, LIN_THREADS, 0, stream>>>(x, W, part, B, N, K, k_per_split);
  DFD_CUDA_OK(cudaGetLastError());
  const int64_t total = static_cast<int64_t>(B) * N;
  const unsigned rgrid = static_cast<unsigned>((total + 255) / 256);
  if (qgelu)
    linear_reduce_kernel<true><<<rgrid, 256, 0, stream>>>(part, splits, bias, res, out, B, N, pre_out);
  else
    linear_reduce_kernel<false><<<rgrid, 256, 0, stream>>>(part, splits, bias, res, out, B, N, pre_out);
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

// x[b, :] += vec[:]   (op_mode.aug_query, models.py:265-267)
__global__ void add_row_vector_kernel(float* __restrict__ x, const float* __restrict__ vec, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) x[i] += vec[i % D];
}

// op_mode.ema_frame (models.py:572-578): the reference's recurrence, same operation order (no FMA contraction)
__global__ void __launch_bounds__(256)
ema_frames_kernel(const float4* __restrict__ x, float4* __restrict__ out, int T, int64_t frame_vec4, float ratio,
                  int64_t total_vec4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_vec4) return;
  const int64_t b = i / frame_vec4, e = i % frame_vec4;
  const float4* src = x + b * T * frame_vec4 + e;
  const float keep = 1.0f - ratio;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float4 v = __ldg(src + t * frame_vec4);
    acc.x = __fadd_rn(__fmul_rn(acc.x, ratio), __fmul_rn(v.x, keep));
    acc.y = __fadd_rn(__fmul_rn(acc.y, ratio), __fmul_rn(v.y, keep));
    acc.z = __fadd_rn(__fmul_rn(acc.z, ratio), __fmul_rn(v.z, keep));
    acc.w = __fadd_rn(__fmul_rn(acc.w, ratio), __fmul_rn(v.w, keep));
  }
  out[i] = acc;
}

int ema_frames(const float* x, float* out, int B, int T, int64_t frame_elems, float ratio, cudaStream_t stream) {
  DFD_CHECK_ARG(B >= 0 && T > 0 && frame_elems > 0 && frame_elems % 4 == 0, "ema_frames: bad shape");
  if (B == 0) return 0;
  DFD_CHECK_ARG(x && out, "ema_frames: null pointer");
  DFD_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "ema_frames: buffers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * (frame_elems / 4);
  ema_frames_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), T, frame_elems / 4, ratio, total);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
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
  float *x, *y, *qs, *mix, *hid, *part, *lin;
  size_t total;
};

static DecWs carve_decoder_ws(void* base, int B, int T, int P, int D, int H, int attn_mode) {
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
  w.part = take(attn_mode ? dec_attn_modes_workspace_bytes(B, T, P, H) : dec_attn_workspace_bytes(B, T, H));
  w.lin = take(linear_workspace_bytes(B, 4 * D));
  w.total = off;
  return w;
}

#define DFD_TIMED(tag, call)          \
  do {                                \
    ScopedTimer _t(ctx, tag, stream); \
    DFD_TRY(call);                    \
  } while (0)

// The decoder as three steps, so that dfd_predict_forward can issue block i as soon as tapped layer i exists:
// begin (query = ln_pre(class_embedding)), one call per block, end (ln_post).
int DecoderRun::begin(cudaStream_t stream) {
  DFD_CHECK_ARG(D == 64 * H, "decoder_forward: width %d != 64 * heads %d", D, H);
  DFD_CHECK_ARG(n_blocks > 0 && B >= 0 && T > 0 && P > 0, "decoder_forward: bad shape");
  if (B == 0) return 0;  // empty batch: nothing to do (buffers of empty tensors are NULL)
  DFD_CHECK_ARG(w && taps && mask && block_out && video_feature, "decoder_forward: null pointer");
  const DecWs ws = carve_decoder_ws(workspace, B, T, P, D, H, w->attn_mode);
  if (!workspace || workspace_bytes < ws.total)
    return fail(DFD_ERR_WORKSPACE, "decoder_forward: workspace %zu < %zu bytes", workspace_bytes, ws.total);
  const int nthr = 256, nblk = (B * D + nthr - 1) / nthr;
  // x = ln_pre(class_embedding) repeated for every clip (models.py:336-337; dropout p = 0)
  DFD_TRY(layernorm(w->class_embedding, w->ln_pre_weight, w->ln_pre_bias, nullptr, 0, nullptr, ws.y, 1, D, stream));
  broadcast_rows_kernel<<<nblk, nthr, 0, stream>>>(ws.y, ws.x, B, D);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int DecoderRun::block(int i, cudaStream_t stream) {
  if (B == 0) return 0;
  const DecWs ws = carve_decoder_ws(workspace, B, T, P, D, H, w->attn_mode);
  const int nthr = 256, nblk = (B * D + nthr - 1) / nthr;
  // x = x + out_proj(attn(in_proj(ln_1(x)), K_i, V_i, m))        (models.py:173-174, 136-146)
  DFD_TIMED(DFD_TAG_DEC_OTHER,
            layernorm(ws.x, w->ln_1_weight[i], w->ln_1_bias[i], nullptr, 0, nullptr, ws.y, B, D, stream));
  DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ctx, ws.y, w->in_proj_weight[i], w->in_proj_bias[i], nullptr, ws.qs, B,
                                           2 * D, D, false, ws.lin, stream));
  if (w->attn_mode) {
    DFD_TIMED(DFD_TAG_DEC_ATTN,
              decoder_attention_modes(ctx, ws.qs, taps->k[i], taps->v[i], taps->stride_b, taps->stride_t,
                                      taps->stride_p, w->positional_embedding, mask, B, T, P, H, w->attn_mode, ws.mix,
                                      ws.part, dec_attn_modes_workspace_bytes(B, T, P, H), stream));
  } else {
    DFD_TIMED(DFD_TAG_DEC_ATTN,
              decoder_attention(ctx, ws.qs, taps->k[i], taps->v[i], taps->stride_b, taps->stride_t, taps->stride_p,
                                w->positional_embedding, mask, B, T, P, H, ws.mix, ws.part,
                                dec_attn_workspace_bytes(B, T, H), stream));
  }
  DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ctx, ws.mix, w->out_proj_weight[i], w->out_proj_bias[i], ws.x, ws.x, B, D,
                                           D, false, ws.lin, stream));
  // x = x + c_proj(quickgelu(c_fc(ln_2(x))))                     (models.py:175)
  DFD_TIMED(DFD_TAG_DEC_OTHER,
            layernorm(ws.x, w->ln_2_weight[i], w->ln_2_bias[i], nullptr, 0, nullptr, ws.y, B, D, stream));
  DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ctx, ws.y, w->c_fc_weight[i], w->c_fc_bias[i], nullptr, ws.hid, B, 4 * D, D,
                                           true, ws.lin, stream));
  DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ctx, ws.hid, w->c_proj_weight[i], w->c_proj_bias[i], ws.x, ws.x, B, D,
                                           4 * D, false, ws.lin, stream));
  scatter_block_out_kernel<<<nblk, nthr, 0, stream>>>(ws.x, block_out, B, D, i, n_blocks);
  DFD_CUDA_OK(cudaGetLastError());
  if (w->augment_query && i + 1 < n_blocks) {  // models.py:265-267 (the recorded block output excludes it)
    DFD_CHECK_ARG(w->augment_query[i] != nullptr, "decoder_forward: augment_query[%d] is NULL", i);
    add_row_vector_kernel<<<nblk, nthr, 0, stream>>>(ws.x, w->augment_query[i], B, D);
    DFD_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int DecoderRun::end(cudaStream_t stream) {
  if (B == 0) return 0;
  const DecWs ws = carve_decoder_ws(workspace, B, T, P, D, H, w->attn_mode);
  // video_feature = ln_post(x_last)                                  (models.py:340-343)
  DFD_TRY(layernorm(ws.x, w->ln_post_weight, w->ln_post_bias, nullptr, 0, nullptr, video_feature, B, D, stream));
  return 0;
}

int decoder_forward(const dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                    const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                    float* video_feature, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DecoderRun run{ctx, D, H, n_blocks, w, taps, mask, B, T, P, block_out, video_feature, workspace, workspace_bytes};
  DFD_TRY(run.begin(stream));
  for (int i = 0; i < n_blocks; ++i) DFD_TRY(run.block(i, stream));
  return run.end(stream);
}

}  // namespace dfd

extern "C" {

size_t dfd_decoder_workspace_bytes(int B, int T, int P, int D, int n_blocks, int attn_mode) {
  (void)n_blocks;
  if (B <= 0 || T <= 0 || P <= 0 || D <= 0) return 0;
  return dfd::carve_decoder_ws(nullptr, B, T, P, D, D / 64, attn_mode).total;
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

size_t dfd_linear_f32_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return dfd::linear_workspace_bytes(B, N);
}

int dfd_linear_f32(dfd_ctx* ctx, const float* x, const float* W, const float* bias, const float* residual, float* out,
                   int B, int N, int K, int quick_gelu, void* workspace, size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32: ctx is NULL");
  if (B < 0 || N <= 0 || K <= 0) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32: bad shape B=%d N=%d K=%d", B, N, K);
  if (B == 0) return 0;
  if (!x || !W || !out) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32: null pointer");
  if (!workspace || workspace_bytes < dfd::linear_workspace_bytes(B, N))
    return dfd::fail(DFD_ERR_WORKSPACE, "dfd_linear_f32: workspace %zu < %zu bytes", workspace_bytes,
                     dfd::linear_workspace_bytes(B, N));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dfd::linear_f32(ctx, x, W, bias, residual, out, B, N, K, quick_gelu != 0, static_cast<float*>(workspace), st);
}

int dfd_ema_frames(dfd_ctx* ctx, const float* x, float* out, int B, int T, int64_t frame_elems, float ratio,
                   void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_ema_frames: ctx is NULL");
  return dfd::ema_frames(x, out, B, T, frame_elems, ratio, static_cast<cudaStream_t>(stream));
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

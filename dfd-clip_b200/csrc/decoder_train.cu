// Training step of the temporal decoder (BASELINE config C5): forward with saved activations and the hand-written
// backward of the one-token-per-clip chain
//   x0 = ln_pre(class_embedding);  per block i:  x += out_proj(attn(in_proj(ln_1(x)), K_i, V_i, m));
//                                                x += c_proj(quickgelu(c_fc(ln_2(x))))          (src/models.py:173-176)
// driven by the reference's trainer (src/trainer.py:147-178: forward(train=True) -> backward -> optimizer.step()).
// The K/V-streaming attention uses dec_attn_stream_kernel / dec_attn_bwd_kernel; everything else here is fp32 SIMT on
// B rows (B = clips per GPU, 12 in the shipped configs):
//   * forward linears: linear_partial/reduce (decoder.cu) — weights streamed once, split-K over all SMs;
//   * dX = dY W: the same weight stream read along its other axis (lin_dx_partial_kernel, split over the output
//     features, fixed-order reduce with the QuickGELU derivative / residual gradient fused in);
//   * dW = dY^T X (+ db): rank-B outer products written once at HBM store speed (lin_dw_kernel);
//   * LayerNorm forward with saved (mean, rstd), backward for dx and for (dgamma, dbeta).
// Everything is deterministic (no atomics). The tail (ln_post, task projection, logit normalisation, loss) stays with
// the caller: it is a few [B, D] x [D, out] operations.
#include "common.cuh"
#include "host_common.h"

namespace dfd {

int linear_f32(const dfd_ctx* ctx, const float* x, const float* W, const float* bias, const float* res, float* out,
               int B, int N, int K, bool qgelu, float* part, cudaStream_t stream, float* pre_out = nullptr);
size_t linear_workspace_bytes(int B, int max_n);
size_t dec_attn_workspace_bytes(int B, int T, int H);
size_t dec_attn_bwd_workspace_bytes(int B, int T, int H);
int decoder_attention(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                      int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B, int T, int P,
                      int H, float* mix, void* workspace, size_t workspace_bytes, cudaStream_t stream, float* stats);
int decoder_attention_backward(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                               int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask,
                               const float* stats, const float* dmix, int B, int T, int P, int H, float* dqs,
                               float* dpos_emb, float* dk, float* dv, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream);

// ------------------------------------------------------------------------------------------- LayerNorm rows
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// y[b,:] = (x[b,:] - mean) * rstd * gamma + beta, stats[b] = (mean, rstd); eps 1e-5, biased variance (models.py:58-68).
// x_row_stride = 0 broadcasts one input row to all B output rows (ln_pre of the class embedding).
__global__ void __launch_bounds__(256)
ln_rows_fwd_kernel(const float* __restrict__ x, int64_t x_row_stride, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ stats, int D) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* xr = x + b * x_row_stride;
  float s = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) s += xr[k];
  const float mean = block_sum_256(s, red) / D;
  float v = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float d = xr[k] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(block_sum_256(v, red) / D + 1e-5f);
  for (int k = threadIdx.x; k < D; k += 256)
    y[static_cast<int64_t>(b) * D + k] = (xr[k] - mean) * rstd * gamma[k] + beta[k];
  if (stats && threadIdx.x == 0) {
    stats[2 * b] = mean;
    stats[2 * b + 1] = rstd;
  }
}

// dx[b,:] = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres[b,:]),  g = dy * gamma,  xhat = (x - mean) * rstd
__global__ void __launch_bounds__(256)
ln_rows_bwd_dx_kernel(const float* __restrict__ x, int64_t x_row_stride, const float* __restrict__ gamma,
                      const float* __restrict__ stats, const float* __restrict__ dy, const float* dres, float* dx,
                      int D) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* xr = x + b * x_row_stride;
  const float* dyr = dy + static_cast<int64_t>(b) * D;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float g = dyr[k] * gamma[k];
    s1 += g;
    s2 = fmaf(g, (xr[k] - mean) * rstd, s2);
  }
  const float m1 = block_sum_256(s1, red) / D;
  const float m2 = block_sum_256(s2, red) / D;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float g = dyr[k] * gamma[k];
    float v = rstd * (g - m1 - (xr[k] - mean) * rstd * m2);
    if (dres) v += dres[static_cast<int64_t>(b) * D + k];
    dx[static_cast<int64_t>(b) * D + k] = v;
  }
}

// dgamma[k] = sum_b dy[b,k] * xhat[b,k], dbeta[k] = sum_b dy[b,k]   (rows summed in order: deterministic)
__global__ void ln_rows_bwd_param_kernel(const float* __restrict__ x, int64_t x_row_stride,
                                         const float* __restrict__ stats, const float* __restrict__ dy,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D) return;
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dy[static_cast<int64_t>(b) * D + k];
    dg = fmaf(d, (x[b * x_row_stride + k] - stats[2 * b]) * stats[2 * b + 1], dg);
    db += d;
  }
  dgamma[k] = dg;
  dbeta[k] = db;
}

// out[k] = sum_b in[b,k]  (optionally out += )
__global__ void column_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int D, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D) return;
  float s = accumulate ? out[k] : 0.f;
  for (int b = 0; b < B; ++b) s += in[static_cast<int64_t>(b) * D + k];
  out[k] = s;
}

// x[b,:] += vec[:]
__global__ void add_rows_kernel(float* __restrict__ x, const float* __restrict__ vec, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) x[i] += vec[i % D];
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------------- linear backward
// dX[b,k] = sum_n dY[b,n] W[n,k]: CTA = (128 columns k, one chunk of DX_NC output features n, 16 rows b); thread = one
// column, 16 row accumulators; W is read once, coalesced along k, DX_NC independent loads per thread issued up front
// (the kernel is a latency chain otherwise: a few hundred small CTAs, one dependent global load per feature); the dY
// chunk sits transposed in shared memory so that one feature's 16 row values are four broadcast 16-byte loads.
// Partials [chunk][B][K] are reduced in a fixed order by lin_dx_reduce_kernel.
constexpr int DX_KT = 128, DX_NC = 16, DX_BT = 16;

__global__ void __launch_bounds__(DX_KT)
lin_dx_partial_kernel(const float* __restrict__ dy, const float* __restrict__ W, float* __restrict__ part, int B, int N,
                      int K) {
  __shared__ __align__(16) float sdy[DX_NC][DX_BT];
  const int k = blockIdx.x * DX_KT + threadIdx.x;
  const int n0 = blockIdx.y * DX_NC, b0 = blockIdx.z * DX_BT;
  const int nn = min(DX_NC, N - n0);
  for (int i = threadIdx.x; i < DX_BT * DX_NC; i += DX_KT) {
    const int c = i / DX_BT, r = i % DX_BT;
    sdy[c][r] = (b0 + r < B && c < nn) ? dy[static_cast<int64_t>(b0 + r) * N + n0 + c] : 0.f;
  }
  float w[DX_NC];
  if (k < K) {
    const float* wp = W + static_cast<int64_t>(n0) * K + k;
#pragma unroll
    for (int c = 0; c < DX_NC; ++c) w[c] = (c < nn) ? __ldg(wp + static_cast<int64_t>(c) * K) : 0.f;
  }
  __syncthreads();
  if (k >= K) return;
  float acc[DX_BT];
#pragma unroll
  for (int r = 0; r < DX_BT; ++r) acc[r] = 0.f;
#pragma unroll
  for (int c = 0; c < DX_NC; ++c) {
#pragma unroll
    for (int r4 = 0; r4 < DX_BT; r4 += 4) {
      const float4 d = *reinterpret_cast<const float4*>(&sdy[c][r4]);
      acc[r4 + 0] = fmaf(d.x, w[c], acc[r4 + 0]);
      acc[r4 + 1] = fmaf(d.y, w[c], acc[r4 + 1]);
      acc[r4 + 2] = fmaf(d.z, w[c], acc[r4 + 2]);
      acc[r4 + 3] = fmaf(d.w, w[c], acc[r4 + 3]);
    }
  }
  float* dst = part + static_cast<int64_t>(blockIdx.y) * B * K;
#pragma unroll
  for (int r = 0; r < DX_BT; ++r)
    if (b0 + r < B) dst[static_cast<int64_t>(b0 + r) * K + k] = acc[r];
}

// dx[i] = (sum_s part[s][i]) * quickgelu'(pre[i]) + add[i];   quickgelu(x) = x sigmoid(1.702 x)  (model.py:166-168)
// CTA = 32 consecutive elements x 8 groups of splits: thread (g, e) sums splits g, g + 8, ... of element e (coalesced
// 128-byte rows, independent loads), the 8 group sums are added in a fixed order through shared memory.
__global__ void __launch_bounds__(256)
lin_dx_reduce_kernel(const float* __restrict__ part, int splits, const float* __restrict__ pre, const float* add,
                     float* dx, int64_t total) {
  __shared__ float red[8][32];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + e;
  float v = 0.f;
  if (i < total)
    for (int s = g; s < splits; s += 8) v += part[s * total + i];
  red[g][e] = v;
  __syncthreads();
  if (g != 0 || i >= total) return;
  v = red[0][e];
#pragma unroll
  for (int j = 1; j < 8; ++j) v += red[j][e];
  if (pre) {
    const float z = pre[i];
    const float sg = 1.f / (1.f + __expf(-1.702f * z));
    v *= sg * (1.f + 1.702f * z * (1.f - sg));
  }
  if (add) v += add[i];
  dx[i] = v;
}

// dW[n,k] = sum_b dY[b,n] X[b,k]; db[n] = sum_b dY[b,n].  CTA tile 32 n x 128 k, 256 threads, thread = 4 n x 4 k.
constexpr int DW_NT = 32, DW_KT = 128, DW_BCH = 16;

__global__ void __launch_bounds__(256)
lin_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dW, float* __restrict__ db,
              int B, int N, int K) {
  __shared__ float sdy[DW_BCH][DW_NT];
  __shared__ __align__(16) float sx[DW_BCH][DW_KT];
  const int n0 = blockIdx.y * DW_NT, k0 = blockIdx.x * DW_KT;
  const int tk = threadIdx.x & 31, tn = threadIdx.x >> 5;  // columns k0 + 4 tk .. +3, rows n0 + 4 tn .. +3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;  // threads 0..31: db of row n0 + threadIdx.x
  for (int bb = 0; bb < B; bb += DW_BCH) {
    const int nb = min(DW_BCH, B - bb);
    __syncthreads();
    for (int i = threadIdx.x; i < DW_BCH * DW_NT; i += 256) {
      const int r = i / DW_NT, c = i % DW_NT;
      sdy[r][c] = (r < nb && n0 + c < N) ? dy[static_cast<int64_t>(bb + r) * N + n0 + c] : 0.f;
    }
    for (int i = threadIdx.x; i < DW_BCH * DW_KT; i += 256) {
      const int r = i / DW_KT, c = i % DW_KT;
      sx[r][c] = (r < nb && k0 + c < K) ? x[static_cast<int64_t>(bb + r) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < DW_BCH; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&sx[r][4 * tk]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = sdy[r][4 * tn + i];
        acc[i][0] = fmaf(d, xv.x, acc[i][0]);
        acc[i][1] = fmaf(d, xv.y, acc[i][1]);
        acc[i][2] = fmaf(d, xv.z, acc[i][2]);
        acc[i][3] = fmaf(d, xv.w, acc[i][3]);
      }
    }
    if (db && blockIdx.x == 0 && threadIdx.x < DW_NT)
      for (int r = 0; r < DW_BCH; ++r) bsum += sdy[r][threadIdx.x];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * tn + i;
    const int k = k0 + 4 * tk;
    if (n < N && k + 3 < K)
      *reinterpret_cast<float4*>(dW + static_cast<int64_t>(n) * K + k) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  if (db && blockIdx.x == 0 && threadIdx.x < DW_NT && n0 + threadIdx.x < N) db[n0 + threadIdx.x] = bsum;
}

size_t linear_bwd_workspace_bytes(int B, int max_n, int max_k) {
  const size_t splits = (static_cast<size_t>(max_n) + DX_NC - 1) / DX_NC;
  return splits * B * max_k * sizeof(float);
}

// dx = (dy W) [* quickgelu'(gelu_pre)] [+ dx_add]; dW = dy^T x; db = column sums of dy. Any output may be NULL.
// dw_stream: the stream of the weight-gradient kernel (independent of the dx chain); NULL = `stream`.
int linear_f32_backward(const dfd_ctx* ctx, const float* x, const float* W, const float* dy, const float* gelu_pre,
                        const float* dx_add, float* dx, float* dW, float* db, int B, int N, int K, float* part,
                        cudaStream_t stream, cudaStream_t dw_stream = nullptr) {
  DFD_CHECK_ARG(K % 4 == 0, "linear_f32_backward: K=%d must be a multiple of 4", K);
  if (dx) {
    const int splits = (N + DX_NC - 1) / DX_NC;
    dim3 grid((K + DX_KT - 1) / DX_KT, splits, (B + DX_BT - 1) / DX_BT);
    lin_dx_partial_kernel<<<grid, DX_KT, 0, stream>>>(dy, W, part, B, N, K);
    DFD_CUDA_OK(cudaGetLastError());
    const int64_t total = static_cast<int64_t>(B) * K;
    lin_dx_reduce_kernel<<<static_cast<unsigned>((total + 31) / 32), 256, 0, stream>>>(part, splits, gelu_pre, dx_add, dx,
                                                                                       total);
    DFD_CUDA_OK(cudaGetLastError());
  }
  if (dW) {
    dim3 grid((K + DW_KT - 1) / DW_KT, (N + DW_NT - 1) / DW_NT);
    lin_dw_kernel<<<grid, 256, 0, dw_stream ? dw_stream : stream>>>(dy, x, dW, db, B, N, K);
    DFD_CUDA_OK(cudaGetLastError());
  }
  (void)ctx;
  return 0;
}

// ------------------------------------------------------------------------------------------- whole decoder
// Saved activations of one forward (fp32), per block: x_in, y1, qs, mix, x1, y2, hpre, h, LayerNorm stats, attention
// stats; plus x0's LayerNorm stats. Scratch (gradients in flight, linear partials, attention partials) behind them.
struct TrainBuf {
  // per block (index with blk(i))
  size_t x_in, y1, qs, mix, x1, y2, hpre, h, st1, st2, ast, blk_stride, blk0;
  size_t st_pre;
  // scratch
  size_t dx, dx1, dy, dqs, dmix, dh, dpe_tmp, lin, attn, total;
};

static TrainBuf train_layout(int B, int T, int D, int H, int n_blocks) {
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  TrainBuf t{};
  const size_t f = sizeof(float), BD = static_cast<size_t>(B) * D;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += up(bytes); return r; };
  t.x_in = take(BD * f); t.y1 = take(BD * f); t.qs = take(2 * BD * f); t.mix = take(BD * f); t.x1 = take(BD * f);
  t.y2 = take(BD * f); t.hpre = take(4 * BD * f); t.h = take(4 * BD * f); t.st1 = take(2 * B * f);
  t.st2 = take(2 * B * f); t.ast = take(static_cast<size_t>(B) * H * 66 * f);
  t.blk_stride = o;
  t.blk0 = 0;
  o = t.blk_stride * n_blocks;
  t.st_pre = take(2 * f * B);
  t.dx = take(BD * f); t.dx1 = take(BD * f); t.dy = take(BD * f); t.dqs = take(2 * BD * f); t.dmix = take(BD * f);
  t.dh = take(4 * BD * f);
  t.dpe_tmp = take(static_cast<size_t>(T) * D * f);
  const size_t lin_f = linear_workspace_bytes(B, 4 * D), lin_b = linear_bwd_workspace_bytes(B, 4 * D, 4 * D);
  t.lin = take(lin_f > lin_b ? lin_f : lin_b);
  const size_t a_f = dec_attn_workspace_bytes(B, T, H), a_b = dec_attn_bwd_workspace_bytes(B, T, H);
  t.attn = take(a_f > a_b ? a_f : a_b);
  t.total = o;
  return t;
}

#define DFD_TIMED(tag, call)          \
  do {                                \
    ScopedTimer _t(ctx, tag, stream); \
    DFD_TRY(call);                    \
  } while (0)

static int check_train_args(int D, int H, int n_blocks, const dfd_decoder_weights* w, const dfd_kv_taps* taps,
                            const uint8_t* mask, int B, int T, int P, const void* saved, size_t saved_bytes,
                            const TrainBuf& tb) {
  DFD_CHECK_ARG(D == 64 * H && D % 4 == 0, "decoder_train: width %d != 64 * heads %d", D, H);
  DFD_CHECK_ARG(n_blocks > 0 && B > 0 && T > 0 && P > 0, "decoder_train: bad shape");
  DFD_CHECK_ARG(w && taps && mask, "decoder_train: null pointer");
  DFD_CHECK_ARG(w->attn_mode == 0, "decoder_train: op_mode.attn_mode is not supported by the native training step");
  if (!saved || saved_bytes < tb.total)
    return fail(DFD_ERR_WORKSPACE, "decoder_train: buffer %zu < %zu bytes", saved_bytes, tb.total);
  return 0;
}

int DecoderTrainRun::begin(cudaStream_t stream) {
  const TrainBuf tb = train_layout(B, T, D, H, n_blocks);
  DFD_TRY(check_train_args(D, H, n_blocks, w, taps, mask, B, T, P, saved, saved_bytes, tb));
  DFD_CHECK_ARG(block_out != nullptr, "decoder_train_forward: block_out is NULL");
  uint8_t* base = static_cast<uint8_t*>(saved);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  // x0 = ln_pre(class_embedding) for every clip (models.py:336-337), written as block 0's input
  ln_rows_fwd_kernel<<<B, 256, 0, stream>>>(w->class_embedding, 0, w->ln_pre_weight, w->ln_pre_bias, F(tb.x_in),
                                            F(tb.st_pre), D);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

int DecoderTrainRun::block(int i, cudaStream_t stream) {
  const TrainBuf tb = train_layout(B, T, D, H, n_blocks);
  uint8_t* base = static_cast<uint8_t*>(saved);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const size_t BD = static_cast<size_t>(B) * D;
  float* lin = F(tb.lin);
  const size_t bo = tb.blk_stride * i;
  float *x_in = F(bo + tb.x_in), *y1 = F(bo + tb.y1), *qs = F(bo + tb.qs), *mix = F(bo + tb.mix), *x1 = F(bo + tb.x1),
        *y2 = F(bo + tb.y2), *hpre = F(bo + tb.hpre), *h = F(bo + tb.h);
  if (i > 0 && w->augment_query) {  // models.py:265-267: added after the block output has been recorded
    DFD_CHECK_ARG(w->augment_query[i - 1] != nullptr, "decoder_train: augment_query[%d] is NULL", i - 1);
    add_rows_kernel<<<(static_cast<int>(BD) + 255) / 256, 256, 0, stream>>>(x_in, w->augment_query[i - 1], B, D);
    DFD_CUDA_OK(cudaGetLastError());
  }
  {
    ScopedTimer _t(ctx, DFD_TAG_DEC_OTHER, stream);
    ln_rows_fwd_kernel<<<B, 256, 0, stream>>>(x_in, D, w->ln_1_weight[i], w->ln_1_bias[i], y1, F(bo + tb.st1), D);
    DFD_CUDA_OK(cudaGetLastError());
  }
  DFD_TIMED(DFD_TAG_DEC_LINEAR,
            linear_f32(ctx, y1, w->in_proj_weight[i], w->in_proj_bias[i], nullptr, qs, B, 2 * D, D, false, lin, stream));
  DFD_TIMED(DFD_TAG_DEC_ATTN,
            decoder_attention(ctx, qs, taps->k[i], taps->v[i], taps->stride_b, taps->stride_t, taps->stride_p,
                              w->positional_embedding, mask, B, T, P, H, mix, F(tb.attn),
                              dec_attn_workspace_bytes(B, T, H), stream, F(bo + tb.ast)));
  DFD_TIMED(DFD_TAG_DEC_LINEAR,
            linear_f32(ctx, mix, w->out_proj_weight[i], w->out_proj_bias[i], x_in, x1, B, D, D, false, lin, stream));
  {
    ScopedTimer _t(ctx, DFD_TAG_DEC_OTHER, stream);
    ln_rows_fwd_kernel<<<B, 256, 0, stream>>>(x1, D, w->ln_2_weight[i], w->ln_2_bias[i], y2, F(bo + tb.st2), D);
    DFD_CUDA_OK(cudaGetLastError());
  }
  DFD_TIMED(DFD_TAG_DEC_LINEAR, linear_f32(ctx, y2, w->c_fc_weight[i], w->c_fc_bias[i], nullptr, h, B, 4 * D, D, true,
                                           lin, stream, hpre));
  // the block output is the next block's input (dense, saved for its backward) and row i of block_out (strided)
  float* x2 = (i + 1 < n_blocks) ? F(bo + tb.blk_stride + tb.x_in) : F(tb.dx);
  DFD_TIMED(DFD_TAG_DEC_LINEAR,
            linear_f32(ctx, h, w->c_proj_weight[i], w->c_proj_bias[i], x1, x2, B, D, 4 * D, false, lin, stream));
  DFD_CUDA_OK(cudaMemcpy2DAsync(block_out + static_cast<size_t>(i) * D, static_cast<size_t>(n_blocks) * D * sizeof(float),
                                x2, D * sizeof(float), D * sizeof(float), B, cudaMemcpyDeviceToDevice, stream));
  return 0;
}

int decoder_train_forward(const dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                          const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                          void* saved, size_t saved_bytes, cudaStream_t stream) {
  DecoderTrainRun run{ctx, D, H, n_blocks, w, taps, mask, B, T, P, block_out, saved, saved_bytes};
  DFD_TRY(run.begin(stream));
  for (int i = 0; i < n_blocks; ++i) DFD_TRY(run.block(i, stream));
  return 0;
}

// grads: same layout as dfd_decoder_weights, every pointer an OUTPUT (fp32, shapes of the parameters); ln_post_* are
// not touched (the tail belongs to the caller). d_block_out: fp32 [B, n_blocks, D] gradient of every block output.
// Blocks block_hi .. block_lo (descending) are processed by this call; a caller that wants to do something between
// blocks (all-reduce a finished block's gradients while the next block's backward runs) walks from n_blocks - 1 down
// to 0 in several calls: the gradient in flight lives in `saved`. positional_embedding's gradient is complete, and
// class_embedding / ln_pre gradients are written, by the call that includes block 0.
// The weight gradients (dW = dy^T x, rank-B outer products: 26 MB written per block) do not feed the chain
// dx2 -> dh -> dy2 -> dx1 -> dmix -> dqs -> dy1 -> dx, which is a string of small latency-bound kernels at B = 12. With
// `side` != NULL each lin_dw_kernel runs on that stream behind an event recorded when its dy exists, beside the chain;
// the chain waits for them once per block, before the kernel that overwrites the block's incoming gradient (the last
// buffer a dW kernel still reads). Event fork/join only: CUDA-graph capturable, results bit-identical.
static int decoder_train_backward_on(const dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                                     const dfd_decoder_weights* grads, const dfd_kv_taps* taps, const uint8_t* mask,
                                     int B, int T, int P, const float* d_block_out, float* const* dk, float* const* dv,
                                     void* saved, size_t saved_bytes, int block_hi, int block_lo, cudaStream_t stream,
                                     cudaStream_t side, bool* side_forked) {
  // events 32.. of the context's pool (0..n_blocks-1 belong to the forward's taps)
  auto dy_ready = [&](int which) -> int {  // the side stream may start the dW kernel of the gradient just produced
    if (!side) return 0;
    cudaEvent_t e = ctx->tap_events[32 + which];
    DFD_CUDA_OK(cudaEventRecord(e, stream));
    DFD_CUDA_OK(cudaStreamWaitEvent(side, e, 0));
    *side_forked = true;
    return 0;
  };
  auto join_side = [&]() -> int {
    if (!side || !*side_forked) return 0;
    DFD_CUDA_OK(cudaEventRecord(ctx->tap_events[36], side));
    DFD_CUDA_OK(cudaStreamWaitEvent(stream, ctx->tap_events[36], 0));
    *side_forked = false;
    return 0;
  };
  const TrainBuf tb = train_layout(B, T, D, H, n_blocks);
  DFD_TRY(check_train_args(D, H, n_blocks, w, taps, mask, B, T, P, saved, saved_bytes, tb));
  DFD_CHECK_ARG(grads && d_block_out, "decoder_train_backward: null pointer");
  DFD_CHECK_ARG(0 <= block_lo && block_lo <= block_hi && block_hi < n_blocks,
                "decoder_train_backward: block range [%d, %d] outside [0, %d)", block_lo, block_hi, n_blocks);
  uint8_t* base = static_cast<uint8_t*>(saved);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  auto G = [](const float* p) { return const_cast<float*>(p); };
  const size_t BD = static_cast<size_t>(B) * D;
  float *dx = F(tb.dx), *dx1 = F(tb.dx1), *dy = F(tb.dy), *dqs = F(tb.dqs), *dmix = F(tb.dmix), *dh = F(tb.dh);
  float* lin = F(tb.lin);
  const int cthr = 256, cblk = (D + cthr - 1) / cthr;
  bool pe_written = block_hi != n_blocks - 1;  // an earlier call already wrote the last block's share
  // dx = gradient wrt the output of the last block
  for (int i = block_hi; i >= block_lo; --i) {
    const size_t bo = tb.blk_stride * i;
    float *x_in = F(bo + tb.x_in), *y1 = F(bo + tb.y1), *qs = F(bo + tb.qs), *mix = F(bo + tb.mix),
          *x1 = F(bo + tb.x1), *y2 = F(bo + tb.y2), *hpre = F(bo + tb.hpre), *h = F(bo + tb.h);
    // dx2 = d_block_out[:, i] (+ the gradient arriving from block i + 1, already in dx)
    if (i == n_blocks - 1) {
      DFD_CUDA_OK(cudaMemcpy2DAsync(dx, D * sizeof(float), d_block_out + static_cast<size_t>(i) * D,
                                    static_cast<size_t>(n_blocks) * D * sizeof(float), D * sizeof(float), B,
                                    cudaMemcpyDeviceToDevice, stream));
    } else {
      DFD_CUDA_OK(cudaMemcpy2DAsync(dy, D * sizeof(float), d_block_out + static_cast<size_t>(i) * D,
                                    static_cast<size_t>(n_blocks) * D * sizeof(float), D * sizeof(float), B,
                                    cudaMemcpyDeviceToDevice, stream));
      add_inplace_kernel<<<static_cast<unsigned>((BD + 255) / 256), 256, 0, stream>>>(dx, dy, static_cast<int64_t>(BD));
      DFD_CUDA_OK(cudaGetLastError());
    }
    // x2 = x1 + c_proj(h): dW_proj, db_proj, dhpre = (dx2 W_proj) * quickgelu'(hpre)
    DFD_TRY(dy_ready(0));
    DFD_TIMED(DFD_TAG_DEC_LINEAR,
              linear_f32_backward(ctx, h, w->c_proj_weight[i], dx, hpre, nullptr, dh, G(grads->c_proj_weight[i]),
                                  G(grads->c_proj_bias[i]), B, D, 4 * D, lin, stream, side));
    // hpre = c_fc(y2): dW_fc, db_fc, dy2
    DFD_TRY(dy_ready(1));
    DFD_TIMED(DFD_TAG_DEC_LINEAR,
              linear_f32_backward(ctx, y2, w->c_fc_weight[i], dh, nullptr, nullptr, dy, G(grads->c_fc_weight[i]),
                                  G(grads->c_fc_bias[i]), B, 4 * D, D, lin, stream, side));
    // y2 = ln_2(x1): dx1 = dx2 + ln_2'(dy2)
    {
      ScopedTimer _t(ctx, DFD_TAG_DEC_OTHER, stream);
      ln_rows_bwd_param_kernel<<<cblk, cthr, 0, stream>>>(x1, D, F(bo + tb.st2), dy, G(grads->ln_2_weight[i]),
                                                          G(grads->ln_2_bias[i]), B, D);
      ln_rows_bwd_dx_kernel<<<B, 256, 0, stream>>>(x1, D, w->ln_2_weight[i], F(bo + tb.st2), dy, dx, dx1, D);
      DFD_CUDA_OK(cudaGetLastError());
    }
    // x1 = x_in + out_proj(mix): dW_out, db_out, dmix
    DFD_TRY(dy_ready(2));
    DFD_TIMED(DFD_TAG_DEC_LINEAR,
              linear_f32_backward(ctx, mix, w->out_proj_weight[i], dx1, nullptr, nullptr, dmix,
                                  G(grads->out_proj_weight[i]), G(grads->out_proj_bias[i]), B, D, D, lin, stream, side));
    // mix = attention(qs, K_i, V_i): dqs, dpos_emb (summed over the blocks), optionally dK_i / dV_i
    float* dpe = nullptr;
    if (w->positional_embedding) dpe = pe_written ? F(tb.dpe_tmp) : G(grads->positional_embedding);
    DFD_TIMED(DFD_TAG_DEC_ATTN,
              decoder_attention_backward(ctx, qs, taps->k[i], taps->v[i], taps->stride_b, taps->stride_t,
                                         taps->stride_p, w->positional_embedding, mask, F(bo + tb.ast), dmix, B, T, P,
                                         H, dqs, dpe, dk ? dk[i] : nullptr, dv ? dv[i] : nullptr, F(tb.attn),
                                         dec_attn_bwd_workspace_bytes(B, T, H), stream));
    if (dpe && pe_written) {
      const int64_t n = static_cast<int64_t>(T) * D;
      add_inplace_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(G(grads->positional_embedding),
                                                                                     dpe, n);
      DFD_CUDA_OK(cudaGetLastError());
    }
    if (dpe) pe_written = true;
    // qs = in_proj(y1): dW_in, db_in, dy1
    DFD_TRY(dy_ready(3));
    DFD_TIMED(DFD_TAG_DEC_LINEAR,
              linear_f32_backward(ctx, y1, w->in_proj_weight[i], dqs, nullptr, nullptr, dy, G(grads->in_proj_weight[i]),
                                  G(grads->in_proj_bias[i]), B, 2 * D, D, lin, stream, side));
    // y1 = ln_1(x_in): dx_in = dx1 + ln_1'(dy1)
    {
      ScopedTimer _t(ctx, DFD_TAG_DEC_OTHER, stream);
      ln_rows_bwd_param_kernel<<<cblk, cthr, 0, stream>>>(x_in, D, F(bo + tb.st1), dy, G(grads->ln_1_weight[i]),
                                                          G(grads->ln_1_bias[i]), B, D);
      // dx is about to be overwritten: the block's weight-gradient kernels (c_proj's reads dx) must be done
      DFD_TRY(join_side());
      ln_rows_bwd_dx_kernel<<<B, 256, 0, stream>>>(x_in, D, w->ln_1_weight[i], F(bo + tb.st1), dy, dx1, dx, D);
      DFD_CUDA_OK(cudaGetLastError());
    }
    if (i > 0 && w->augment_query) {  // x_in = block_out[i-1] + aug[i-1]: daug = column sums of dx
      column_sum_kernel<<<cblk, cthr, 0, stream>>>(dx, G(grads->augment_query[i - 1]), B, D, 0);
      DFD_CUDA_OK(cudaGetLastError());
    }
  }
  if (block_lo > 0) return 0;
  // x0[b] = ln_pre(class_embedding) for every b: the B gradient rows collapse onto the one input row
  column_sum_kernel<<<cblk, cthr, 0, stream>>>(dx, dy, B, D, 0);                        // dy[0,:] = sum_b dx[b,:]
  ln_rows_bwd_param_kernel<<<cblk, cthr, 0, stream>>>(w->class_embedding, 0, F(tb.st_pre), dy,
                                                      G(grads->ln_pre_weight), G(grads->ln_pre_bias), 1, D);
  ln_rows_bwd_dx_kernel<<<1, 256, 0, stream>>>(w->class_embedding, 0, w->ln_pre_weight, F(tb.st_pre), dy, nullptr,
                                               G(grads->class_embedding), D);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

static bool bwd_side_stream_enabled() {  // read per call (a host-side getenv): tests and A/B runs toggle it
  const char* e = getenv("DFD_BWD_STREAMS");
  return !(e && e[0] == '0');
}

int decoder_train_backward(const dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                           const dfd_decoder_weights* grads, const dfd_kv_taps* taps, const uint8_t* mask, int B, int T,
                           int P, const float* d_block_out, float* const* dk, float* const* dv, void* saved,
                           size_t saved_bytes, int block_hi, int block_lo, cudaStream_t stream) {
  // per-kernel timing wants every kernel alone on the device; DFD_BWD_STREAMS=0 keeps everything on `stream`
  cudaStream_t side = nullptr;
  if (!ctx->timing && bwd_side_stream_enabled() && ctx->side_stream && ctx->tap_events.size() > 36)
    side = ctx->side_stream;
  bool forked = false;
  const int rc = decoder_train_backward_on(ctx, D, H, n_blocks, w, grads, taps, mask, B, T, P, d_block_out, dk, dv,
                                           saved, saved_bytes, block_hi, block_lo, stream, side, &forked);
  if (forked) {  // an error return in the middle of a block: never leave the side stream forked (graph capture)
    cudaEventRecord(ctx->tap_events[36], side);
    cudaStreamWaitEvent(stream, ctx->tap_events[36], 0);
  }
  return rc;
}

}  // namespace dfd

extern "C" {

size_t dfd_decoder_train_bytes(int B, int T, int D, int n_blocks) {
  if (B <= 0 || T <= 0 || D <= 0 || n_blocks <= 0) return 0;
  return dfd::train_layout(B, T, D, D / 64, n_blocks).total;
}

int dfd_decoder_train_forward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                              const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                              void* saved, size_t saved_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_train_forward: ctx is NULL");
  return dfd::decoder_train_forward(ctx, D, H, n_blocks, w, taps, mask, B, T, P, block_out, saved, saved_bytes,
                                    static_cast<cudaStream_t>(stream));
}

int dfd_decoder_train_backward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                               const dfd_decoder_weights* grads, const dfd_kv_taps* taps, const uint8_t* mask, int B,
                               int T, int P, const float* d_block_out, float* const* dk, float* const* dv, void* saved,
                               size_t saved_bytes, int block_hi, int block_lo, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_decoder_train_backward: ctx is NULL");
  return dfd::decoder_train_backward(ctx, D, H, n_blocks, w, grads, taps, mask, B, T, P, d_block_out, dk, dv, saved,
                                     saved_bytes, block_hi, block_lo, static_cast<cudaStream_t>(stream));
}

int dfd_linear_f32_backward(dfd_ctx* ctx, const float* x, const float* W, const float* dy, const float* gelu_pre,
                            const float* dx_add, float* dx, float* dW, float* db, int B, int N, int K, void* workspace,
                            size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32_backward: ctx is NULL");
  if (B <= 0 || N <= 0 || K <= 0) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32_backward: bad shape");
  if (!x || !W || !dy) return dfd::fail(DFD_ERR_INVALID, "dfd_linear_f32_backward: null pointer");
  if (dx && (!workspace || workspace_bytes < dfd::linear_bwd_workspace_bytes(B, N, K)))
    return dfd::fail(DFD_ERR_WORKSPACE, "dfd_linear_f32_backward: workspace %zu < %zu bytes", workspace_bytes,
                     dfd::linear_bwd_workspace_bytes(B, N, K));
  return dfd::linear_f32_backward(ctx, x, W, dy, gelu_pre, dx_add, dx, dW, db, B, N, K,
                                  static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

size_t dfd_linear_f32_backward_workspace_bytes(int B, int N, int K) {
  if (B <= 0 || N <= 0 || K <= 0) return 0;
  return dfd::linear_bwd_workspace_bytes(B, N, K);
}

}  // extern "C"

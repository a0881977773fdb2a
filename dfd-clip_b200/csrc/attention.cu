// Encoder self-attention (src/clip/model.py:188-195): per frame and head, softmax_k((q/8).k) v over the
// L = P+1 tokens of ONE frame (no mask). One CTA per (frame, head); Q, K, V of that pair are staged once in
// swizzled shared memory; each warp owns one 16-row query tile and runs a flash-style pass over the keys in
// blocks of 64 (online softmax in registers, exp2 with the 1/8 scale folded into the exponent).
// Tensor work is issued with mma.sync (bf16, fp32 accumulate).
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace dfd {

namespace attn {
constexpr int DH = 64;
constexpr int ROW_BYTES = DH * 2;  // 128
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows][128 B] tile with XOR swizzle
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

template <int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
mha_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ mix, int L, int H) {
  using namespace attn;
  extern __shared__ __align__(128) uint8_t smem[];
  const int LP = (L + 15) & ~15;
  const int D = H * DH;
  const int h = blockIdx.x % H;
  const int f = blockIdx.x / H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sQ = smem_u32(smem), sK = sQ + LP * ROW_BYTES, sV = sK + LP * ROW_BYTES;

  // ---- stage Q, K, V (rows >= L are zero filled)
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(f) * L * 3 * D + h * DH;
  for (int t = threadIdx.x; t < 3 * LP * 8; t += blockDim.x) {
    const int which = t / (LP * 8);
    const int r = (t / 8) % LP, ch = t % 8;
    const bool valid = r < L;
    const __nv_bfloat16* src = base + static_cast<int64_t>(valid ? r : 0) * 3 * D + which * D + ch * 8;
    cp_async16(sQ + which * LP * ROW_BYTES + sw_off(r, ch), src, valid);
  }
  cp_async_wait_all();
  __syncthreads();

  const int q0 = warp * 16;
  if (q0 >= LP) return;

  // ---- Q fragments for the 4 k-steps over dh
  uint32_t qf[4][4];
  {
    const int r = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm_x4(sQ + sw_off(r, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = 0.125f * 1.4426950408889634f;  // (1/sqrt(dh)) * log2(e)

  for (int kb = 0; kb < LP; kb += 64) {
    const int ntile16 = min(4, (LP - kb) >> 4);  // 16-key tiles in this block (warp-uniform)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    // S = Q K^T
#pragma unroll
    for (int t16 = 0; t16 < 4; ++t16) {
      if (t16 < ntile16) {
        // lanes 0-7: keys 0-7 / d-chunk 0, 8-15: keys 0-7 / chunk 1, 16-23: keys 8-15 / chunk 0, 24-31: keys 8-15 / chunk 1
        const int r = kb + t16 * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sK + sw_off(r, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
          mma_bf16_16816(s[t16 * 2 + 0], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
          mma_bf16_16816(s[t16 * 2 + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
        }
      }
    }
    // mask keys beyond L, block row-max
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kb + j * 8 + (lane & 3) * 2;
      if (key >= L) s[j][0] = s[j][2] = -INFINITY;
      if (key + 1 >= L) s[j][1] = s[j][3] = -INFINITY;
      bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
      bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);  // finite: every block holds >= 1 valid key
    const float r0 = exp2f((m0 - mn0) * sc), r1 = exp2f((m1 - mn1) * sc);
    m0 = mn0;
    m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pf[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] * sc - mn0 * sc), p1 = exp2f(s[j][1] * sc - mn0 * sc);
      const float p2 = exp2f(s[j][2] * sc - mn1 * sc), p3 = exp2f(s[j][3] * sc - mn1 * sc);
      ps0 += p0 + p1;
      ps1 += p2 + p3;
      pf[j][0] = pack_bf16(p0, p1);
      pf[j][1] = pack_bf16(p2, p3);
    }
    l0 = l0 * r0 + ps0;
    l1 = l1 * r1 + ps1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= r0; o[j][1] *= r0; o[j][2] *= r1; o[j][3] *= r1;
    }
    // O += P V
#pragma unroll
    for (int t16 = 0; t16 < 4; ++t16) {
      if (t16 < ntile16) {
        // lanes 0-7: keys 0-7 / d-chunk c, 8-15: keys 8-15 / chunk c, 16-23: keys 0-7 / chunk c+1, 24-31: keys 8-15 / c+1
        const int r = kb + t16 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sV + sw_off(r, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
          mma_bf16_16816(o[dp * 2 + 0], pf[t16 * 2][0], pf[t16 * 2][1], pf[t16 * 2 + 1][0], pf[t16 * 2 + 1][1], b0, b1);
          mma_bf16_16816(o[dp * 2 + 1], pf[t16 * 2][0], pf[t16 * 2][1], pf[t16 * 2 + 1][0], pf[t16 * 2 + 1][1], b2, b3);
        }
      }
    }
  }

  // ---- finalize: full row sums across the quad, normalise, stage through this warp's Q rows, coalesced store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __syncwarp();
  {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // element (row, col = 8j + 2t): chunk j, byte offset 4t inside the chunk
      const uint32_t a0 = sQ + sw_off(q0 + g, j) + t * 4;
      const uint32_t a1 = sQ + sw_off(q0 + g + 8, j) + t * 4;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(pack_bf16(o[j][0] * i0, o[j][1] * i0)) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(pack_bf16(o[j][2] * i1, o[j][3] * i1)) : "memory");
    }
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = q0 + it * 4 + (lane >> 3), ch = lane & 7;
    if (r < L) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(sQ + sw_off(r, ch)));
      *reinterpret_cast<uint4*>(mix + (static_cast<int64_t>(f) * L + r) * D + h * DH + ch * 8) = v;
    }
  }
}

int mha_fwd_tc(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);
int mha_fwd_tc2(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);
int mha_fwd_tc3(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);

int mha_fwd(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream) {
  DFD_CHECK_ARG(n_frames >= 0 && L > 0 && H > 0, "mha_fwd: bad shape");
  DFD_CHECK_ARG(L <= 272, "mha_fwd: sequence length %d > 272 tokens per frame is not supported", L);
  if (n_frames == 0) return 0;
  DFD_CHECK_ARG(qkv && mix, "mha_fwd: null pointer");
  // up to 208 tokens per frame (ViT-B/16, B/32): tcgen05 kernel; longer sequences (ViT-L/14: 257): mma.sync kernel
  // 129..208 tokens (two 128-row query tiles): pipelined two-tile kernel with P in TMEM; up to 128: one-tile kernel
  if (L > 128 && L <= 208) return mha_fwd_tc2(ctx, qkv, mix, n_frames, L, H, stream);
  // DFD_MHA_SIMT=1 keeps the mma.sync kernel below for A/B runs
  static const bool simt_only = getenv("DFD_MHA_SIMT") && atoi(getenv("DFD_MHA_SIMT")) != 0;
  if (L > 208 && L <= 257 && !simt_only) return mha_fwd_tc3(ctx, qkv, mix, n_frames, L, H, stream);
  if (L <= 128) return mha_fwd_tc(ctx, qkv, mix, n_frames, L, H, stream);
  const int LP = (L + 15) & ~15;
  const int threads = (LP / 16) * 32;
  const size_t smem = static_cast<size_t>(3) * LP * attn::ROW_BYTES;
  DFD_CHECK_ARG(static_cast<int>(smem) <= ctx->smem_optin, "mha_fwd: needs %zu bytes of shared memory", smem);
  const unsigned grid = static_cast<unsigned>(n_frames) * H;
  if (threads <= 416) {
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_kernel<416>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mha_fwd_kernel<416><<<grid, threads, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                         static_cast<__nv_bfloat16*>(mix), L, H);
  } else {
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_kernel<544>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mha_fwd_kernel<544><<<grid, threads, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                         static_cast<__nv_bfloat16*>(mix), L, H);
  }
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

extern "C" int dfd_mha_fwd(dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_mha_fwd: ctx is NULL");
  return dfd::mha_fwd(ctx, qkv, mix, n_frames, L, H, static_cast<cudaStream_t>(stream));
}

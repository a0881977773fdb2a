// tcgen05 encoder self-attention for sequences of up to 208 tokens per frame (CLIP ViT-B/16: L = 197).
// Reference: MultiheadAttention.forward, src/clip/model.py:188-195 — softmax_k((q/8).k) v per frame and head, no mask.
//
// Work item = (frame, head, 128-row query tile); persistent CTAs (2 per SM), 192 threads each:
//   warp 0      TMA producer: Q tile [128 x 64], K [208 x 64], V [208 x 64] of the item through 3-D tensor maps over
//               the packed QKV buffer ([frame][token][3D]; tokens >= L are out of bounds => zero filled)
//   warp 1      MMA issuer:  S = Q K^T   (tcgen05.mma M=128 N=208 K=16 x4, both operands K-major, fp32 in TMEM)
//                            O = P V     (M=128 N=64 K=16 x13, A = P from smem (K-major), B = V MN-major)
//   warps 2..5  softmax: thread = query row. Two passes over the S row in TMEM (max, then exp2/sum), P written as
//               bf16 into 128B-swizzled smem (over the dead Q/K tiles), O epilogue: tcgen05.ld, 1/rowsum, bf16,
//               swizzled staging, TMA store through a 3-D map of the output (rows >= L clipped).
// TMEM: 256 columns per CTA (S in columns [0,208), O aliases columns [0,64) once S has been consumed).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

namespace attn_tc {
constexpr int QT = 128;             // query rows per item
constexpr int KP = 208;             // padded key count (MMA N of S, multiple of 16)
constexpr int DH = 64;
constexpr int THREADS = 192;
constexpr int P_BYTES = 4 * QT * 128;        // 4 key chunks of 64 keys: [128 rows][128 B] each
constexpr int Q_OFF = 0;                      // Q and K live inside the P region (dead before P is written)
constexpr int K_OFF = QT * 128;               // 16 KB
constexpr int V_OFF = P_BYTES;                // 64 KB
constexpr int V_BYTES = KP * 128;             // 26 KB
constexpr int O_OFF = V_OFF + V_BYTES;        // staging for the output tile: 4 warps x 4 KB
constexpr int O_BYTES = QT * 128;
constexpr int BAR_OFF = O_OFF + O_BYTES;
constexpr int NUM_BARS = 5;                   // load_full, s_full, p_full, o_full, o_empty
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int SMEM_BYTES = TMEM_PTR_OFF + 16 + 1024;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t LOAD_BYTES = QT * 128 + KP * 128 + KP * 128;
static_assert(K_OFF + KP * 128 <= P_BYTES, "Q and K must fit inside the P region");
}  // namespace attn_tc

__global__ void __launch_bounds__(attn_tc::THREADS, 2)
mha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                  const __grid_constant__ CUtensorMap tmO, int L, int H, int num_qt, int num_items) {
  using namespace attn_tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* load_full = bars + 0;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* o_empty = bars + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + TMEM_PTR_OFF);
  // warp-uniform values through a lane-0 shuffle: what derives from them stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int D = H * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(load_full, 1);
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
      mbar_init(o_full, 1);
      mbar_init(o_empty, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qt = item % num_qt, fh = item / num_qt;
        const int h = fh % H, f = fh / H;
        if (it > 0) mbar_wait(o_full, (it - 1) & 1);  // P / V of the previous item fully consumed by its PV MMAs
        mbar_arrive_expect_tx(load_full, LOAD_BYTES);
        tma_load_3d(&tmQ, load_full, smem + Q_OFF, h * DH, qt * QT, f, kEvictFirst);
        tma_load_3d(&tmKV, load_full, smem + K_OFF, D + h * DH, 0, f, kEvictNormal);
        tma_load_3d(&tmKV, load_full, smem + V_OFF, 2 * D + h * DH, 0, f, kEvictNormal);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------ MMA issuer
    // whole warp in the loop, ONE elected lane issues: the tcgen05 operands stay in uniform registers (see
    // attention_sm100_v2.cu for the measurement behind this)
    {
      const bool elected = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KP);
      constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DH, /*b_mn_major=*/true);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(load_full, ph);
        if (it > 0) mbar_wait(o_empty, (it - 1) & 1);  // O of the previous item has been read out of TMEM
        tc_fence_after();
        const uint64_t q_desc = umma_desc_sw128(smem + Q_OFF);
        const uint64_t k_desc = umma_desc_sw128(smem + K_OFF);
        if (elected) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, ph);
        tc_fence_after();
        const uint64_t v_desc = umma_desc_sw128_mn(smem + V_OFF);
        if (elected) {
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk) {
            // A: P chunk kk/4 (16 KB each), 32-byte step inside the swizzle atom; B: 16 keys = 2048 bytes of V rows
            const uint64_t p_desc = umma_desc_sw128(smem + (kk >> 2) * (QT * 128)) + 2 * (kk & 3);
            umma_bf16(tmem_base, p_desc, v_desc + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_o, kk != 0);
          }
          umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------------ softmax + epilogue warps
    const int q = warp & 3;               // TMEM lane quarter
    const int row = q * 32 + lane;        // query row inside the tile
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* obuf = smem + O_OFF + q * 4096;
    const float sc = 0.125f * 1.4426950408889634f;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int qt = item % num_qt, fh = item / num_qt;
      const int h = fh % H, f = fh / H;
      const bool warp_active = (qt * QT + q * 32) < L;  // this warp holds at least one real query row
      mbar_wait(s_full, ph);
      tc_fence_after();
      float inv_sum = 0.f;
      if (warp_active) {
        // pass 1: row max over the L real keys
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
          if (c0 + 32 <= KP) {
            uint32_t r[32];
            tmem_ld32(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < L) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
            uint32_t r[16];
            tmem_ld16(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < L) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        const float mo = mx * sc;
        // pass 2: p = exp2(s*sc - max*sc), row sum, bf16 P into swizzled smem (chunk = 64 keys)
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
          uint32_t pk[16];
          if (c0 + 32 <= KP) {
            uint32_t r[32];
            tmem_ld32(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = (c0 + j < L) ? fast_exp2(fmaf(__uint_as_float(r[j]), sc, -mo)) : 0.f;
              float p1 = (c0 + j + 1 < L) ? fast_exp2(fmaf(__uint_as_float(r[j + 1]), sc, -mo)) : 0.f;
              sum += p0 + p1;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
          } else {
            uint32_t r[16];
            tmem_ld16(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float p0 = (c0 + j < L) ? fast_exp2(fmaf(__uint_as_float(r[j]), sc, -mo)) : 0.f;
              float p1 = (c0 + j + 1 < L) ? fast_exp2(fmaf(__uint_as_float(r[j + 1]), sc, -mo)) : 0.f;
              sum += p0 + p1;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
#pragma unroll
            for (int j = 8; j < 16; ++j) pk[j] = 0u;
          }
          // 32 keys = 64 bytes = 4 sixteen-byte pieces at piece index (c0 % 64)/8 .. +3 of chunk c0/64
          uint8_t* prow = smem + (c0 >> 6) * (QT * 128) + row * 128;
          const int piece0 = (c0 & 63) >> 3;
          const int npieces = (c0 + 32 <= KP) ? 4 : 2;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            if (pc < npieces) {
              uint4 v = make_uint4(pk[pc * 4 + 0], pk[pc * 4 + 1], pk[pc * 4 + 2], pk[pc * 4 + 3]);
              *reinterpret_cast<uint4*>(prow + (((piece0 + pc) ^ (row & 7)) << 4)) = v;
            }
          }
        }
        inv_sum = 1.f / sum;
      }
      // P visible to the tensor core (async proxy), S fully read: hand over to the MMA warp
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);

      mbar_wait(o_full, ph);
      tc_fence_after();
      if (warp_active) {
        uint32_t o0[32], o1[32];
        tmem_ld32(t_row, o0);
        tmem_ld32(t_row + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(o_empty);
        if (lane == 0) tma_store_wait_read<0>();  // previous store out of this warp's staging buffer is done
        __syncwarp();
        uint8_t* dst = obuf + lane * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t* src = (ch < 4) ? (o0 + ch * 8) : (o1 + (ch - 4) * 8);
          uint4 v;
          v.x = pack_bf16(__uint_as_float(src[0]) * inv_sum, __uint_as_float(src[1]) * inv_sum);
          v.y = pack_bf16(__uint_as_float(src[2]) * inv_sum, __uint_as_float(src[3]) * inv_sum);
          v.z = pack_bf16(__uint_as_float(src[4]) * inv_sum, __uint_as_float(src[5]) * inv_sum);
          v.w = pack_bf16(__uint_as_float(src[6]) * inv_sum, __uint_as_float(src[7]) * inv_sum);
          *reinterpret_cast<uint4*>(dst + ((ch ^ (lane & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, obuf, h * DH, qt * QT + q * 32, f);
          tma_store_commit();
        }
      } else {
        tc_fence_before();
        mbar_arrive(o_empty);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int mha_fwd_tc(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream) {
  using namespace attn_tc;
  const int D = H * DH;
  CUtensorMap tmQ, tmKV, tmO;
  const uint64_t frame_ld = static_cast<uint64_t>(L) * 3 * D;
  DFD_TRY(make_tmap_3d(ctx, &tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, 3 * D, 3 * D, frame_ld, QT, DH));
  DFD_TRY(make_tmap_3d(ctx, &tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, 3 * D, 3 * D, frame_ld, KP, DH));
  DFD_TRY(make_tmap_3d(ctx, &tmO, mix, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, D, D,
                       static_cast<uint64_t>(L) * D, 32, DH));
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int num_qt = (L + QT - 1) / QT;
  const int num_items = n_frames * H * num_qt;
  const int max_ctas = 2 * ctx->num_sms;
  const int grid = num_items < max_ctas ? num_items : max_ctas;
  mha_fwd_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, L, H, num_qt, num_items);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

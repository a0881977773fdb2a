// tcgen05 encoder self-attention for 208 < L <= 257 tokens per frame (CLIP ViT-L/14 @ 224: L = 257 = 256 + 1).
// Reference: MultiheadAttention.forward, src/clip/model.py:188-195 — softmax_k((q/8).k) v, no mask.
//
// Same pipeline as the ViT-B/16 kernel (attention_sm100_v2.cu): work item = (frame, head), one persistent CTA per SM,
// warp 0 TMA producer (2-stage ring of Q/K/V tiles), warp 1 one-thread MMA issuer, warps 2..5 / 6..9 softmax +
// epilogue warpgroups of the two 128-row query tiles (thread = query row), S = Q K^T in TMEM, P written back into
// TMEM as bf16 pairs over the dead S columns, O = P V with A = P from TMEM.
//
// What L = 257 changes: the two tiles' fp32 S already fill all 512 TMEM columns at 256 keys, and a third query tile
// would hold one row. So rows / keys 0..255 go through the tensor core and TOKEN 256 is handled around it:
//   * the producer also copies the q / k / v rows of token 256 (3 x 128 B, cp.async.bulk) next to the barriers;
//   * as a KEY: every softmax thread adds s_last = q_row . k_256 (Q row from the swizzled smem tile, k_256 a
//     broadcast smem read) to its row max / row sum, and the epilogue adds p_last * v_256 to its O row;
//   * as a QUERY: warp 10 computes that one row in SIMT from the K / V tiles already in shared memory (lane = key:
//     8 keys per lane for the scores and for P.V with 64 per-lane channel accumulators, then a halving
//     transpose-reduce across the lanes) and stores its 64 outputs directly.
// For 208 < L < 257 (no 257th token) both extras are compiled out at run time and padded keys are masked as in v2.
// TMEM: tile X owns columns [256 X, 256 X + 256): S fp32 [0,256); P bf16x2 [0,128) once S is consumed; O fp32 [128,192).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

namespace attn3 {
constexpr int QT = 128;
constexpr int KP = 256;   // keys through the tensor core
constexpr int DH = 64;
constexpr int THREADS = 352;            // 11 warps
constexpr int TILE_BYTES = KP * 128;    // 32 KB: 256 rows x 64 bf16, 128-byte swizzle
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int Q_OFF = 0, K_OFF = TILE_BYTES, V_OFF = 2 * TILE_BYTES;
constexpr int O_OFF = 2 * STAGE_BYTES;  // 8 warps x 4 KB output staging
constexpr int O_BYTES = 8 * 4096;
constexpr int BAR_OFF = O_OFF + O_BYTES;
constexpr int NUM_BARS = 12;            // load_full[2], stage_empty[2], s_full[2], p_full[2], o_full[2], o_empty[2]
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int LAST_OFF = TMEM_PTR_OFF + 16;   // per stage: q_256, k_256, v_256 rows of the head (3 x 128 B)
constexpr int LAST_BYTES = 3 * 128;
constexpr int PS_OFF = LAST_OFF + 2 * LAST_BYTES;  // 256 fp32 scores of the SIMT query row
constexpr int SMEM_BYTES = PS_OFF + 1024 + 1024;
static_assert(LAST_OFF % 16 == 0, "bulk-copy destinations must be 16-byte aligned");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TILE_COLS = 256;
constexpr uint32_t O_COL = 128;
constexpr int NCHUNK = 8;               // 8 x 32 key columns
static_assert(SMEM_BYTES <= 227 * 1024, "attention v3 shared memory exceeds the per-CTA limit");

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void nbar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// dot product of two 64-element bf16 vectors: `a` from a 128-byte-swizzled smem tile row (row index r8 = row % 8
// selects the chunk permutation), `b` already in registers as 32 packed pairs
__device__ __forceinline__ float dot64_swz(const uint8_t* row_base, int r8, const uint32_t (&b)[32]) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 a = *reinterpret_cast<const uint4*>(row_base + ((c ^ r8) << 4));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc0 = fmaf(bf_lo(aw[j]), bf_lo(b[4 * c + j]), acc0);
      acc1 = fmaf(bf_hi(aw[j]), bf_hi(b[4 * c + j]), acc1);
    }
  }
  return acc0 + acc1;
}

// 64 bf16 (128 B) from shared memory, same address in every lane (broadcast)
__device__ __forceinline__ void lds_row64(const uint8_t* p, uint32_t (&r)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + 16 * c);
    r[4 * c + 0] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace attn3

// LAST = (L == 257): token 256 exists and every tensor-core key chunk is complete, so the masked code paths and
// the token-256 extras are compiled into separate instances (the kernel is instruction-cache sensitive: 11 warps in
// four different roles run through long unrolled bodies).
template <bool LAST>
__global__ void __launch_bounds__(attn3::THREADS, 1)
mha_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmO,
                   const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ mix, int L, int H,
                   int num_items) {
  using namespace attn3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* load_full = bars + 0;
  uint64_t* stage_empty = bars + 2;
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 10;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + TMEM_PTR_OFF);
  // warp-uniform values through a lane-0 shuffle: what derives from them stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int D = H * DH;
  constexpr bool has_last = LAST;               // token 256 exists
  const int Lk = has_last ? KP : L;             // keys held by the tensor-core tiles
  const int n_my = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&load_full[i], 1);
        mbar_init(&stage_empty[i], has_last ? 2 : 1);  // MMA commit (+ the last-row warp)
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 128);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_empty[i], 128);
      }
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
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int h = item % H, f = item / H;
        const int st = it & 1;
        if (it >= 2) mbar_wait(&stage_empty[st], ((it >> 1) & 1) ^ 1);
        uint8_t* sb = smem + st * STAGE_BYTES;
        mbar_arrive_expect_tx(&load_full[st], STAGE_BYTES + (has_last ? LAST_BYTES : 0));
        tma_load_3d(&tmIn, &load_full[st], sb + Q_OFF, h * DH, 0, f, kEvictFirst);
        tma_load_3d(&tmIn, &load_full[st], sb + K_OFF, D + h * DH, 0, f, kEvictFirst);
        tma_load_3d(&tmIn, &load_full[st], sb + V_OFF, 2 * D + h * DH, 0, f, kEvictFirst);
        if (has_last) {  // token 256 of this head: three 128-byte rows next to the barriers
          const __nv_bfloat16* last = qkv + (static_cast<int64_t>(f) * L + KP) * 3 * D + h * DH;
          uint8_t* lb = smem + LAST_OFF + st * LAST_BYTES;
          bulk_g2s(lb, last, 128, &load_full[st]);
          bulk_g2s(lb + 128, last + D, 128, &load_full[st]);
          bulk_g2s(lb + 256, last + 2 * D, 128, &load_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop and ONE elected lane issues the tcgen05 instructions, so their operands stay in
    // uniform registers (under `if (lane == 0)` every MMA paid an ELECT / 4 x R2UR / retry-branch sequence of ~95
    // cycles — three times the tensor-core time of a P.V MMA; measured on the ViT-B/16 kernel, attention_sm100_v2.cu).
    if (n_my > 0) {
      const bool elected = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KP);
      constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DH, /*b_mn_major=*/true);
      auto issue_s = [&](int x, int st) {
        const uint8_t* sb = smem + st * STAGE_BYTES;
        const uint64_t q_desc = umma_desc_sw128(sb + Q_OFF + x * (QT * 128));
        const uint64_t k_desc = umma_desc_sw128(sb + K_OFF);
        if (elected) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem_base + x * TILE_COLS, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[x]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int x, int st) {
        const uint8_t* sb = smem + st * STAGE_BYTES;
        const uint64_t v_desc = umma_desc_sw128_mn(sb + V_OFF);
        const uint32_t t_tile = tmem_base + x * TILE_COLS;
        if (elected) {
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk)
            umma_bf16_ts(t_tile + O_COL, t_tile + kk * 8, v_desc + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_o,
                         kk != 0);
          umma_commit(&o_full[x]);
        }
        __syncwarp();
      };
      // issue order PV_A(i), S_A(i+1), PV_B(i), S_B(i+1), as in the v2 kernel
      mbar_wait(&load_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int it = 0; it < n_my; ++it) {
        const int st = it & 1;
        const uint32_t ph = it & 1;
        const bool nxt = it + 1 < n_my;
        const int nst = (it + 1) & 1;
        mbar_wait(&p_full[0], ph);
        tc_fence_after();
        issue_pv(0, st);
        if (nxt) {
          mbar_wait(&load_full[nst], ((it + 1) >> 1) & 1);
          mbar_wait(&o_empty[0], ph);
          tc_fence_after();
          issue_s(0, nst);
        }
        mbar_wait(&p_full[1], ph);
        tc_fence_after();
        issue_pv(1, st);
        if (elected) umma_commit(&stage_empty[st]);
        __syncwarp();
        if (nxt) {
          mbar_wait(&o_empty[1], ph);
          tc_fence_after();
          issue_s(1, nst);
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------------------ query row 256 (SIMT)
    // lane = key (keys lane, lane + 32, ..., lane + 224) for the scores AND for P.V: every lane accumulates all 64
    // channels over its 8 keys (64 independent FMA chains), then a halving transpose-reduce leaves channels
    // 2*lane, 2*lane + 1 in lane `lane`. The two key loops are rolled (scores parked in shared memory): this role's
    // code must stay small next to the unrolled softmax bodies.
    if (has_last) {
      const float sc = 0.125f * 1.4426950408889634f;
      float* psc = reinterpret_cast<float*>(smem + PS_OFF) + lane;
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int h = item % H, f = item / H;
        const int st = it & 1;
        mbar_wait(&load_full[st], (it >> 1) & 1);
        const uint8_t* sb = smem + st * STAGE_BYTES;
        const uint8_t* lb = smem + LAST_OFF + st * LAST_BYTES;
        float s_last, mx = -INFINITY;
        {
          uint32_t qv[32];
          lds_row64(lb, qv);          // q_256
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int j = lane + 32 * i;
            const float sv = dot64_swz(sb + K_OFF + (j >> 3) * 1024 + (j & 7) * 128, j & 7, qv);
            psc[32 * i] = sv;
            mx = fmaxf(mx, sv);
          }
          s_last = dot64_swz(lb + 128, 0, qv);   // k_256 (plain row: chunk permutation 0)
        }
        mx = fmaxf(warp_max(mx), s_last);
        const float mo = mx * sc;
        const float p_last = fast_exp2(fmaf(s_last, sc, -mo));
        float sum = 0.f;
        float acc[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.f;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int j = lane + 32 * i;
          const float pj = fast_exp2(fmaf(psc[32 * i], sc, -mo));
          sum += pj;
          const uint8_t* vrow = sb + V_OFF + (j >> 3) * 1024 + (j & 7) * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(vrow + ((c ^ (j & 7)) << 4));
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              acc[8 * c + 2 * w] = fmaf(pj, bf_lo(aw[w]), acc[8 * c + 2 * w]);
              acc[8 * c + 2 * w + 1] = fmaf(pj, bf_hi(aw[w]), acc[8 * c + 2 * w + 1]);
            }
          }
        }
        sum = warp_sum(sum) + p_last;
        const uint32_t vl = *reinterpret_cast<const uint32_t*>(lb + 256 + 4 * lane);  // v_256 channel pair
        __syncwarp();
        if (lane == 0) mbar_arrive(&stage_empty[st]);  // this warp no longer reads the stage
        // transpose-reduce: in the round with offset o a lane keeps the half of its values selected by (lane & o),
        // so bit k of the lane index picks the upper half at offset 2^k and lane l ends with channels 2l, 2l + 1
#pragma unroll
        for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int c = 0; c < n; ++c) {
            const float send = up ? acc[c] : acc[c + n];
            const float keep = up ? acc[c + n] : acc[c];
            acc[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        const float inv = 1.f / sum;
        const float o0 = fmaf(p_last, bf_lo(vl), acc[0]) * inv;
        const float o1 = fmaf(p_last, bf_hi(vl), acc[1]) * inv;
        reinterpret_cast<uint32_t*>(mix + (static_cast<int64_t>(f) * L + KP) * D + h * DH)[lane] = pack_bf16(o0, o1);
      }
    }
  } else {
    // ------------------------------------------------------------------------------ softmax + epilogue warps
    const int x = (warp - 2) >> 2;         // query tile of this warpgroup
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int row0 = x * QT + q * 32;      // first query row of this warp inside the frame
    const int row = row0 + lane;
    const bool warp_active = row0 < L;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + x * TILE_COLS;
    uint8_t* obuf = smem + O_OFF + (warp - 2) * 4096;
    const float sc = 0.125f * 1.4426950408889634f;
    int h = static_cast<int>(blockIdx.x) % H, f = static_cast<int>(blockIdx.x) / H;
    const int dh_step = static_cast<int>(gridDim.x) % H, df_step = static_cast<int>(gridDim.x) / H;
    // MUFU ping-pong between the two warpgroups (named barriers 1 and 2, 256 threads each), as in the v2 kernel
    if (x == 1 && n_my > 0) nbar_arrive(1, 256);
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ph = it & 1;
      const int st = it & 1;
      const uint8_t* lb = smem + LAST_OFF + st * LAST_BYTES;
      mbar_wait(&s_full[x], ph);
      tc_fence_after();
      float inv_sum = 0.f, mo = 0.f, s_last = 0.f, p_last = 0.f;
      if (warp_active) {
        uint32_t ra[32], rb[32];
        // ---- pass 1: row max over the keys. The chunk loop is rolled in bodies of two 32-column chunks (double
        // buffer ra / rb: the TMEM load of the next chunk is in flight while the current one is reduced).
        float mx0 = -INFINITY, mx1 = -INFINITY;
        tmem_ld32(t_row, ra);
        if (has_last) {
          // the 257th key: s_last = q_row . k_256 (runs under the latency of the first TMEM load)
          uint32_t kl[32];
          lds_row64(lb + 128, kl);
          s_last = dot64_swz(smem + st * STAGE_BYTES + Q_OFF + (row >> 3) * 1024 + (row & 7) * 128, row & 7, kl);
          mx0 = s_last;
        }
        auto max_chunk = [&](const uint32_t (&cur)[32], int c0) {
          if (LAST || c0 + 32 <= Lk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              mx0 = max3(mx0, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
              mx1 = max3(mx1, __uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < Lk) mx0 = fmaxf(mx0, __uint_as_float(cur[j]));
          }
        };
#pragma unroll 1
        for (int c = 0; c < NCHUNK; c += 2) {
          tmem_ld_wait();
          tmem_ld32(t_row + (c + 1) * 32, rb);
          max_chunk(ra, c * 32);
          tmem_ld_wait();
          if (c + 2 < NCHUNK) tmem_ld32(t_row + (c + 2) * 32, ra);
          max_chunk(rb, (c + 1) * 32);
        }
        mo = fmaxf(mx0, mx1) * sc;
      }
      nbar_sync(1 + x, 256);  // wait for this tile's turn on the MUFU pipe
      if (warp_active) {
        uint32_t ra[32], rb[32];
        // ---- pass 2: p = exp2(s*sc - max*sc), row sum, bf16 pairs back into TMEM columns [16c, 16c+16).
        // Rolled in bodies of two chunks and software pipelined: the exp2 of a chunk are issued (MUFU) before the
        // sums / packing / tcgen05.st of the chunk before it, so the MUFU pipe always has independent work queued.
        float sum0 = 0.f, sum1 = 0.f;
        float ea[32], eb[32];
        tmem_ld32(t_row, ra);
        if (has_last) {
          p_last = fast_exp2(fmaf(s_last, sc, -mo));
          sum0 = p_last;
        }
        auto exp_chunk = [&](const uint32_t (&cur)[32], float (&e)[32], int c0) {
          if (LAST || c0 + 32 <= Lk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(cur[j]), sc, -mo));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              e[j] = (c0 + j < Lk) ? fast_exp2(fmaf(__uint_as_float(cur[j]), sc, -mo)) : 0.f;
          }
        };
        auto finish_chunk = [&](const float (&e)[32], int cp) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            sum0 += e[j];
            sum1 += e[j + 1];
            pk[j >> 1] = pack_bf16(e[j], e[j + 1]);
          }
          tmem_st16(t_row + cp * 16, pk);
        };
#pragma unroll 1
        for (int c = 0; c < NCHUNK; c += 2) {
          tmem_ld_wait();
          tmem_ld32(t_row + (c + 1) * 32, rb);
          exp_chunk(ra, ea, c * 32);
          if (c > 0) finish_chunk(eb, c - 1);
          tmem_ld_wait();
          if (c + 2 < NCHUNK) tmem_ld32(t_row + (c + 2) * 32, ra);
          exp_chunk(rb, eb, (c + 1) * 32);
          finish_chunk(ea, c);
        }
        finish_chunk(eb, NCHUNK - 1);
        tmem_st_wait();
        inv_sum = 1.f / (sum0 + sum1);
      }
      // pass the MUFU turn to the other tile (tile B does not hand back after its last item)
      if (x == 0 || it + 1 < n_my) nbar_arrive(2 - x, 256);
      // v_256 for the rank-1 update of the epilogue. Read BEFORE handing the tile to the MMA warp: once P.V of both
      // tiles has retired the producer may refill this stage's token-256 rows.
      uint32_t vl[32];
      if (has_last && warp_active) lds_row64(lb + 256, vl);
      // P is in TMEM, S fully consumed: hand the tile to the MMA warp
      tc_fence_before();
      mbar_arrive(&p_full[x]);

      mbar_wait(&o_full[x], ph);
      tc_fence_after();
      if (warp_active) {
        uint32_t o0[32], o1[32];
        tmem_ld32(t_row + O_COL, o0);
        tmem_ld32(t_row + O_COL + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_empty[x]);
        if (lane == 0) tma_store_wait_read<0>();  // previous store out of this warp's staging buffer is done
        __syncwarp();
        uint8_t* dst = obuf + lane * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t* src = (ch < 4) ? (o0 + ch * 8) : (o1 + (ch - 4) * 8);
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(src[j]);
          if (has_last) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[2 * j] = fmaf(p_last, bf_lo(vl[ch * 4 + j]), v[2 * j]);
              v[2 * j + 1] = fmaf(p_last, bf_hi(vl[ch * 4 + j]), v[2 * j + 1]);
            }
          }
          uint4 o;
          o.x = pack_bf16(v[0] * inv_sum, v[1] * inv_sum);
          o.y = pack_bf16(v[2] * inv_sum, v[3] * inv_sum);
          o.z = pack_bf16(v[4] * inv_sum, v[5] * inv_sum);
          o.w = pack_bf16(v[6] * inv_sum, v[7] * inv_sum);
          *reinterpret_cast<uint4*>(dst + ((ch ^ (lane & 7)) << 4)) = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, obuf, h * DH, row0, f);
          tma_store_commit();
        }
      } else {
        tc_fence_before();
        mbar_arrive(&o_empty[x]);
      }
      h += dh_step;
      f += df_step;
      if (h >= H) {
        h -= H;
        ++f;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int mha_fwd_tc3(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream) {
  using namespace attn3;
  DFD_CHECK_ARG(L > 208 && L <= KP + 1, "mha_fwd_tc3: needs 208 < L <= 257, got %d", L);
  const int D = H * DH;
  CUtensorMap tmIn, tmO;
  const uint64_t frame_ld = static_cast<uint64_t>(L) * 3 * D;
  // rows beyond min(L, 256) - 1 of a frame are out of bounds for the tile loads (zero-filled); token 256 is read
  // straight from global memory by the threads that need it
  DFD_TRY(make_tmap_3d(ctx, &tmIn, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L > KP ? KP : L, 3 * D, 3 * D,
                       frame_ld, KP, DH));
  DFD_TRY(make_tmap_3d(ctx, &tmO, mix, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L > KP ? KP : L, D, D,
                       static_cast<uint64_t>(L) * D, 32, DH));
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int num_items = n_frames * H;
  const int grid = num_items < ctx->num_sms ? num_items : ctx->num_sms;
  if (L > KP)
    mha_fwd_tc3_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmIn, tmO, static_cast<const __nv_bfloat16*>(qkv),
                                                                    static_cast<__nv_bfloat16*>(mix), L, H, num_items);
  else
    mha_fwd_tc3_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmIn, tmO, static_cast<const __nv_bfloat16*>(qkv),
                                                                     static_cast<__nv_bfloat16*>(mix), L, H, num_items);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

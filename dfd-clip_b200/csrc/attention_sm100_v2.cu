// tcgen05 encoder self-attention, pipelined two-tile kernel for 128 < L <= 208 tokens per frame (CLIP ViT-B/16:
// L = 197). Reference: MultiheadAttention.forward, src/clip/model.py:188-195 — softmax_k((q/8).k) v, no mask.
//
// Work item = (frame, head); one persistent CTA per SM, 352 threads:
//   warp 0        TMA producer: Q [256 x 64] (rows >= L zero-filled), K [208 x 64], V [208 x 64] of the item into a
//                 2-stage smem ring, so the loads of item i+1 overlap all compute of item i.
//   warp 1 / 10   MMA issuer of query tile A / B (128 rows each). The whole warp walks the loop with uniform control
//                 flow and ONE elected lane issues, so TMEM addresses and smem descriptors stay in uniform registers
//                 (under `if (lane == 0)` every tcgen05.mma paid an ELECT / 4 x R2UR / retry-branch sequence of ~95
//                 cycles: the 13 P.V MMAs of ~32 tensor-core cycles each took 1 250 cycles on the critical chain).
//                   S_X = Q_X K^T        tcgen05.mma M=128 N=208 K=16 x4, operands in smem, fp32 S in TMEM
//                   O_X = P_X V          tcgen05.mma M=128 N=64  K=16 x13, A = P from TMEM (bf16 pairs written by the
//                                        softmax warps over the dead S columns), B = V from smem (MN-major)
//                 Per tile: PV(i) as soon as its P is complete, S(i+1) as soon as its O is drained; the two tiles'
//                 chains only meet at the smem stage release (both commit to stage_empty).
//   warps 2..5    softmax + epilogue warpgroup of tile A (thread = query row), warps 6..9 the same for tile B:
//                 pass 1 row max out of TMEM, pass 2 exp2 / row sum / bf16 P back into TMEM (tcgen05.st), then
//                 O: tcgen05.ld, 1/rowsum, bf16, swizzled smem staging, TMA store (rows >= L clipped by the map).
//                 The MUFU-bound pass 2 is ping-ponged between the two warpgroups with named barriers, so the tiles
//                 settle half a period apart. For L = 197 (compile-time instance) pass 2 runs in 16-key steps with
//                 the next step's arguments and the previous step's sums / packing interleaved with the MUFU stream
//                 (softmax_pass2_scheduled).
// TMEM (512 columns): tile X owns columns [256 X, 256 X + 208): S fp32 [0,208); P bf16x2 [0,104) once S is consumed;
// O fp32 [128,192) (S columns that pass 2 has already read).
// Round-2 timeline at C2 (clock64, profiles/r2_mha_timeline.md): per tile max 500 + exp2 2 430 + P.V 650 + drain 700 +
// next S 300-800 cycles; 152 -> 132 us per launch (512 frames x 12 heads).
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace dfd {

#ifdef DFD_MHA_TRACE
__device__ long long g_mha_trace[64 * 16];
#define TRACE(slot) do { if (blockIdx.x == 0 && it < 64 && (threadIdx.x & 31) == 0) g_mha_trace[it * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace attn2 {
constexpr int QT = 128;
constexpr int KP = 208;
constexpr int DH = 64;
constexpr int THREADS = 352;  // producer, MMA issuer of tile A, 2 x 4 softmax warps, MMA issuer of tile B
constexpr int Q_BYTES = 2 * QT * 128;   // 32 KB
constexpr int KV_BYTES = KP * 128;      // 26 KB
constexpr int STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;
constexpr int Q_OFF = 0, K_OFF = Q_BYTES, V_OFF = Q_BYTES + KV_BYTES;
constexpr int O_OFF = 2 * STAGE_BYTES;  // 8 warps x 4 KB output staging
constexpr int O_BYTES = 8 * 4096;
constexpr int BAR_OFF = O_OFF + O_BYTES;
// barriers: load_full[2], stage_empty[2], s_full[2], p_full[2], o_full[2], o_empty[2]
constexpr int NUM_BARS = 12;
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int SMEM_BYTES = TMEM_PTR_OFF + 16 + 1024;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TILE_COLS = 256;
constexpr uint32_t O_COL = 128;
constexpr uint32_t LOAD_BYTES = STAGE_BYTES;
constexpr int NCHUNK = 7;  // 6 x 32 + 16 key columns
static_assert(STAGE_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0, "tiles must stay 1024-byte aligned");
}  // namespace attn2

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x (K/2) 32-bit columns of packed bf16 pairs.
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

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (one row per thread).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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

// chunk c of the S row: 32 fp32 columns (c < 6) or 16 (c == 6) into r[0..)
__device__ __forceinline__ void ld_chunk(uint32_t t_s, int c, uint32_t (&r)[32]) {
  if (c < 6) {
    tmem_ld32(t_s + c * 32, r);
  } else {
    uint32_t (&h)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
    tmem_ld16(t_s + c * 32, h);
  }
}

// e[j] = 2^(s_j * sc - mo) for j < w (w = 32 or 16), element j on the pipe softmax_exp2<j> selects
template <int J>
__device__ __forceinline__ void exp2_chunk(float (&e)[32], const uint32_t (&cur)[32], float sc, float mo, int w) {
  if constexpr (J < 32) {
    if (J < 16 || w == 32) e[J] = softmax_exp2<J>(fmaf(__uint_as_float(cur[J]), sc, -mo));
    exp2_chunk<J + 1>(e, cur, sc, mo, w);
  }
}

// Softmax pass 2 with a compile-time sequence length, in steps of 16 keys, software-pipelined so that every MUFU.EX2
// has independent FMA-pipe work next to it. Why: with one softmax warp per sub-partition the pass is a single in-order
// instruction stream; MUFU.EX2 issues 4 lanes per clock (8 cycles per warp instruction), and in the chunked version
// above ptxas clusters 24-32 MUFUs back to back (nothing else is ready: the next chunk's arguments wait on its TMEM
// load, the sums wait on the MUFU results), then runs the FFMA / FADD / F2FP bursts with the MUFU pipe idle — ncu
// source view (profiles/r2_mha_sass_schedule.md): ~14 cycles per element against the pipe's 8. Here step k issues, per
// key j:   e_k[j] = ex2(arg_k[j])   |   arg_{k+1}[j] = s_{k+1}[j] * sc - max   |   sum += e_{k-1}[j], pack e_{k-1}
// — three independent streams (S of step k+1 was loaded during step k-1), so the scheduler can put ~3 FMA-pipe
// instructions into each MUFU's 8-cycle shadow. Keys >= LCT are compile-time zeros (no MUFU, no select).
template <int LCT>
__device__ __forceinline__ float softmax_pass2_scheduled(uint32_t t_row, float sc, float mo) {
  constexpr int NSTEP = 13;  // 208 / 16
  uint32_t sa[16], sb[16];   // raw S of steps k+1 / k+2 (alternating)
  float arga[16], argb[16];  // exp2 arguments of steps k / k+1
  float ea[16], eb[16];      // exp2 results of steps k / k-1
  float sum0 = 0.f, sum1 = 0.f;
  // prologue: S of step 0 -> arguments of step 0; S of step 1 in flight
  tmem_ld16(t_row, sa);
  tmem_ld_wait();
  tmem_ld16(t_row + 16, sb);
#pragma unroll
  for (int j = 0; j < 16; ++j) arga[j] = fmaf(__uint_as_float(sa[j]), sc, -mo);
#pragma unroll
  for (int k = 0; k <= NSTEP; ++k) {
    // S of step k+1 has landed (issued one step ago); start the load of step k+2 into the other buffer
    uint32_t (&s_next)[16] = (k & 1) ? sa : sb;    // step k+1
    uint32_t (&s_after)[16] = (k & 1) ? sb : sa;   // step k+2
    float (&arg_cur)[16] = (k & 1) ? argb : arga;
    float (&arg_next)[16] = (k & 1) ? arga : argb;
    float (&e_cur)[16] = (k & 1) ? eb : ea;
    float (&e_prev)[16] = (k & 1) ? ea : eb;
    if (k + 1 < NSTEP) tmem_ld_wait();
    if (k + 2 < NSTEP) tmem_ld16(t_row + (k + 2) * 16, s_after);
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (k < NSTEP) e_cur[j] = (k * 16 + j < LCT) ? fast_exp2(arg_cur[j]) : 0.f;
      if (k + 1 < NSTEP && (k + 1) * 16 + j < LCT) arg_next[j] = fmaf(__uint_as_float(s_next[j]), sc, -mo);
      if (k > 0) {
        if (j & 1) {
          sum1 += e_prev[j];
          pk[j >> 1] = pack_bf16(e_prev[j - 1], e_prev[j]);
        } else {
          sum0 += e_prev[j];
        }
      }
    }
    if (k > 0) tmem_st8(t_row + (k - 1) * 8, pk);
  }
  tmem_st_wait();
  return sum0 + sum1;
}

template <int LCT>
__global__ void __launch_bounds__(attn2::THREADS, 1)
mha_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmO, int L_rt, int H, int num_items) {
  using namespace attn2;
  const int L = LCT > 0 ? LCT : L_rt;
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
  // warp-uniform by construction: tell the compiler so (values that pass through shfl from lane 0 are treated as
  // uniform, which keeps everything derived from them — TMEM addresses, smem descriptors — in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int D = H * DH;
  const int n_my = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&load_full[i], 1);
        mbar_init(&stage_empty[i], 2);  // both tiles' MMA issuers commit their last read of the stage
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
        mbar_arrive_expect_tx(&load_full[st], LOAD_BYTES);
        tma_load_3d(&tmQ, &load_full[st], sb + Q_OFF, h * DH, 0, f, kEvictFirst);
        tma_load_3d(&tmKV, &load_full[st], sb + K_OFF, D + h * DH, 0, f, kEvictFirst);
        tma_load_3d(&tmKV, &load_full[st], sb + V_OFF, 2 * D + h * DH, 0, f, kEvictFirst);
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------------------------ MMA issuers
    // One issuing thread PER TILE (warp 1: tile A, warp 10: tile B). With a single thread serving both tiles in a
    // fixed order, a tile whose P was ready had to wait until the thread had finished waiting for the OTHER tile's
    // epilogue (clock64 timeline, profiles/r2_mha_timeline.md: P written -> O ready 1 150 - 1 450 cycles for 13 MMAs
    // of ~35 cycles): head-of-line blocking on the critical chain S -> max -> exp -> PV -> drain of each tile.
    // The whole warp walks the loop (uniform control flow, every lane polls the barriers) and ONE elected lane
    // executes the tcgen05 instructions: with the loop under `if (lane == 0)` every operand lived in a vector register
    // and each MMA cost an ELECT / 4 x R2UR / retry-branch sequence (~95 cycles per issue in the clock64 timeline:
    // the 13 P.V MMAs of ~32 tensor-core cycles each took 1 250 cycles).
    const int x = warp == 1 ? 0 : 1;
    if (n_my > 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KP);
      constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DH, /*b_mn_major=*/true);
      const uint32_t t_tile = tmem_base + x * TILE_COLS;
      auto issue_s = [&](int st) {
        const uint8_t* sb = smem + st * STAGE_BYTES;
        const uint64_t q_desc = umma_desc_sw128(sb + Q_OFF + x * (QT * 128));
        const uint64_t k_desc = umma_desc_sw128(sb + K_OFF);
        if (leader) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(t_tile, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[x]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int st) {
        const uint8_t* sb = smem + st * STAGE_BYTES;
        const uint64_t v_desc = umma_desc_sw128_mn(sb + V_OFF);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk)
            umma_bf16_ts(t_tile + O_COL, t_tile + kk * 8, v_desc + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_o,
                         kk != 0);
          umma_commit(&o_full[x]);
          umma_commit(&stage_empty[st]);  // this tile's last read of the stage (the producer waits for both tiles)
        }
        __syncwarp();
      };
      mbar_wait(&load_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int it = 0; it < n_my; ++it) {
        const int st = it & 1;
        const uint32_t ph = it & 1;
        mbar_wait(&p_full[x], ph);
        tc_fence_after();
        if (x == 0) TRACE(0);
        issue_pv(st);
        if (it + 1 < n_my) {
          const int nst = (it + 1) & 1;
          mbar_wait(&load_full[nst], ((it + 1) >> 1) & 1);
          if (x == 0) TRACE(1);
          mbar_wait(&o_empty[x], ph);    // O (aliasing S columns) drained by this tile's epilogue
          tc_fence_after();
          if (x == 0) TRACE(2);
          issue_s(nst);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------ softmax + epilogue warps
    const int x = (warp - 2) >> 2;         // query tile of this warpgroup (warps 2..5: A, 6..9: B)
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int row0 = x * QT + q * 32;      // first query row of this warp inside the frame
    const bool warp_active = row0 < L;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + x * TILE_COLS;
    uint8_t* obuf = smem + O_OFF + (warp - 2) * 4096;
    const float sc = 0.125f * 1.4426950408889634f;
    int h = static_cast<int>(blockIdx.x) % H, f = static_cast<int>(blockIdx.x) / H;
    const int dh_step = static_cast<int>(gridDim.x) % H, df_step = static_cast<int>(gridDim.x) / H;
    // Ping-pong of the MUFU-bound pass 2 between the two warpgroups (named barriers 1 and 2, 256 threads each):
    // tile A's exp2 phase runs while tile B is in its MMA / max / epilogue phases and vice versa, so the two tiles
    // never halve each other's MUFU rate and settle half a period apart.
    if (x == 1 && n_my > 0) nbar_arrive(1, 256);
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(&s_full[x], ph);
      tc_fence_after();
      if (warp == 2) TRACE(5);
      if (warp == 6) TRACE(10);
      float inv_sum = 0.f, mo = 0.f;
      if (warp_active) {
        uint32_t ra[32], rb[32];
        // ---- pass 1: row max over the L real keys (chunk c+1 is in flight while chunk c is reduced)
        float mx0 = -INFINITY, mx1 = -INFINITY;
        ld_chunk(t_row, 0, ra);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t (&cur)[32] = (c & 1) ? rb : ra;
          uint32_t (&nx)[32] = (c & 1) ? ra : rb;
          tmem_ld_wait();
          if (c + 1 < NCHUNK) ld_chunk(t_row, c + 1, nx);
          const int c0 = c * 32, w = (c < 6) ? 32 : 16;
          if (c < 4 || c0 + w <= L) {  // this kernel runs for L > 128: chunks 0..3 are always complete
#pragma unroll
            for (int j = 0; j < w; j += 4) {
              mx0 = max3(mx0, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
              mx1 = max3(mx1, __uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < w; ++j)
              if (c0 + j < L) mx0 = fmaxf(mx0, __uint_as_float(cur[j]));
          }
        }
        mo = fmaxf(mx0, mx1) * sc;
      }
      nbar_sync(1 + x, 256);  // wait for this tile's turn on the MUFU pipe
      if (warp == 2) TRACE(6);
      if (warp == 6) TRACE(11);
      if (warp_active) {
        if constexpr (LCT > 0) {
          inv_sum = 1.f / softmax_pass2_scheduled<LCT>(t_row, sc, mo);
        } else {
          uint32_t ra[32], rb[32];
          // ---- pass 2: p = exp2(s*sc - max*sc), row sum, bf16 pairs back into TMEM columns [16c, 16c+16).
          // Software pipelined: the exp2 of chunk c are issued (MUFU) before the sums / packing / tcgen05.st of
          // chunk c-1, so the MUFU pipe always has independent work queued behind it.
          float sum0 = 0.f, sum1 = 0.f;
          float ea[32], eb[32];
          ld_chunk(t_row, 0, ra);
  #pragma unroll
          for (int c = 0; c <= NCHUNK; ++c) {
            if (c < NCHUNK) {
              uint32_t (&cur)[32] = (c & 1) ? rb : ra;
              uint32_t (&nx)[32] = (c & 1) ? ra : rb;
              float (&e)[32] = (c & 1) ? eb : ea;
              tmem_ld_wait();
              if (c + 1 < NCHUNK) ld_chunk(t_row, c + 1, nx);
              const int c0 = c * 32, w = (c < 6) ? 32 : 16;
              if (c < 4 || c0 + w <= L) {
                exp2_chunk<0>(e, cur, sc, mo, w);
              } else {
  #pragma unroll
                for (int j = 0; j < w; ++j)
                  e[j] = (c0 + j < L) ? fast_exp2(fmaf(__uint_as_float(cur[j]), sc, -mo)) : 0.f;
              }
            }
            if (c > 0) {
              const int cp = c - 1, wp = (cp < 6) ? 32 : 16;
              float (&e)[32] = (cp & 1) ? eb : ea;
              uint32_t pk[16];
  #pragma unroll
              for (int j = 0; j < wp; j += 2) {
                sum0 += e[j];
                sum1 += e[j + 1];
                pk[j >> 1] = pack_bf16(e[j], e[j + 1]);
              }
              if (cp < 6)
                tmem_st16(t_row + cp * 16, pk);
              else
                tmem_st8(t_row + cp * 16, pk);
            }
          }
          tmem_st_wait();
          inv_sum = 1.f / (sum0 + sum1);
        }
      }
      // pass the MUFU turn to the other tile (tile B does not hand back after its last item)
      if (x == 0 || it + 1 < n_my) nbar_arrive(2 - x, 256);
      // P is in TMEM, S fully consumed: hand the tile to the MMA warp
      tc_fence_before();
      if (warp == 2) TRACE(7);
      if (warp == 6) TRACE(12);
      mbar_arrive(&p_full[x]);

      mbar_wait(&o_full[x], ph);
      tc_fence_after();
      if (warp == 2) TRACE(8);
      if (warp == 6) TRACE(13);
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
          tma_store_3d(&tmO, obuf, h * DH, row0, f);
          tma_store_commit();
        }
      } else {
        tc_fence_before();
        mbar_arrive(&o_empty[x]);
      }
      if (warp == 2) TRACE(9);
      if (warp == 6) TRACE(14);
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

int mha_fwd_tc2(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream) {
  using namespace attn2;
  DFD_CHECK_ARG(L > QT && L <= KP, "mha_fwd_tc2: needs 128 < L <= 208, got %d", L);
  const int D = H * DH;
  CUtensorMap tmQ, tmKV, tmO;
  const uint64_t frame_ld = static_cast<uint64_t>(L) * 3 * D;
  DFD_TRY(make_tmap_3d(ctx, &tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, 3 * D, 3 * D, frame_ld, 2 * QT, DH));
  DFD_TRY(make_tmap_3d(ctx, &tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, 3 * D, 3 * D, frame_ld, KP, DH));
  DFD_TRY(make_tmap_3d(ctx, &tmO, mix, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_frames, L, D, D,
                       static_cast<uint64_t>(L) * D, 32, DH));
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DFD_CUDA_OK(cudaFuncSetAttribute(mha_fwd_tc2_kernel<197>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int num_items = n_frames * H;
  const int grid = num_items < ctx->num_sms ? num_items : ctx->num_sms;
  // L = 197 (every 224-pixel / patch-16 CLIP ViT): the instance with the sequence length at compile time and the
  // scheduled softmax pass; DFD_MHA_GENERIC=1 keeps the run-time-L instance for A/B runs
  static const bool generic = getenv("DFD_MHA_GENERIC") && atoi(getenv("DFD_MHA_GENERIC")) != 0;
  if (L == 197 && !generic)
    mha_fwd_tc2_kernel<197><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, L, H, num_items);
  else
    mha_fwd_tc2_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, L, H, num_items);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfd

#ifdef DFD_MHA_TRACE
extern "C" int dfd_debug_mha_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, dfd::g_mha_trace, sizeof(long long) * 64 * 16);
}
#endif

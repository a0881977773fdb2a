// tcgen05 GEMM for the encoder's dense contractions (patch embedding, QKV, out-proj, MLP):
//   out (epilogue) A[M,K] * W[N,K]^T,  bf16 operands, fp32 accumulation in TMEM.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: A and W tiles -> 4-stage smem ring (128-byte swizzle), mbarrier expect_tx
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16) on the smem descriptors,
//               tcgen05.commit releases smem stages and publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator (thread = row), bias / QuickGELU, stage the tile in
//               swizzled smem and write it with TMA (store, or reduce-add for the in-place residual)
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.
//
// Reference ops replaced: F.linear / nn.Linear / nn.Conv2d in src/clip/model.py:186,197,209,211,277.
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace dfd {

namespace gemm {
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_STAGE = BM * BK * 2;           // 16 KB
constexpr int B_STAGE = BN * BK * 2;           // 32 KB
constexpr int OUT_BUF = 32 * 128;              // one epilogue staging box: 32 rows x 128 B
constexpr int OUT_BUFS_PER_WARP = 2;
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
constexpr int OFF_OUT = OFF_B + STAGES * B_STAGE;
constexpr int OFF_BAR = OFF_OUT + EPI_WARPS * OUT_BUFS_PER_WARP * OUT_BUF;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;  // + slack for manual 1024 B alignment
constexpr uint32_t TMEM_COLS = 512;
}  // namespace gemm

// ---------------------------------------------------------------------------------------------------------
// Epilogue shared by the 1-SM and the SM-pair kernel: one warp drains `box_count` boxes (32 rows x 128 bytes of
// output each: 64 bf16 or 32 fp32 columns) of its TMEM lane quarter, starting at box `box_begin` of the tile.
//   tcgen05.ld -> bias / LayerNorm fold / QuickGELU / erf-GELU -> 128-byte-swizzled smem box -> TMA store or reduce-add
// `release()` is called by every lane once the warp's last TMEM read has completed (the accumulator can be reused).
namespace gemm_epi {
constexpr int OUT_BUF = 32 * 128;
constexpr int OUT_BUFS_PER_WARP = 2;
}  // namespace gemm_epi

template <int EPI, bool kLnFold, class Release>
__device__ __forceinline__ void epilogue_boxes(uint32_t t_row, int col0, int row0, int M, int box_begin, int box_count,
                                               const float* __restrict__ bias, const float* __restrict__ colsum,
                                               float ln_mu, float ln_rstd, uint8_t* obuf, int& buf,
                                               const CUtensorMap* tmC, Release&& release) {
  using namespace gemm_epi;
  constexpr bool kOutF32 = (EPI == DFD_EPI_STORE_F32 || EPI == DFD_EPI_ADD_F32);
  constexpr int COLS_PER_BOX = kOutF32 ? 32 : 64;
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int box = box_begin; box < box_begin + box_count; ++box) {
    const int c = col0 + box * COLS_PER_BOX;
    uint32_t packed[32];  // 128 bytes of output for this thread's row
    if constexpr (kOutF32) {
      uint32_t r[32];
      tmem_ld32(t_row + box * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c + j)) : make_float4(0, 0, 0, 0);
        packed[j + 0] = __float_as_uint(__uint_as_float(r[j + 0]) + b4.x);
        packed[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + b4.y);
        packed[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + b4.z);
        packed[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + b4.w);
      }
    } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(t_row + box * 64 + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4 =
              bias ? __ldg(reinterpret_cast<const float4*>(bias + c + half * 32 + j)) : make_float4(0, 0, 0, 0);
          float v0, v1, v2, v3;
          if constexpr (kLnFold) {
            // LayerNorm folded into this GEMM: out = rstd * (acc - mu * colsum[n]) + bias'[n]
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(colsum + c + half * 32 + j));
            v0 = fmaf(ln_rstd, fmaf(-ln_mu, c4.x, __uint_as_float(r[j + 0])), b4.x);
            v1 = fmaf(ln_rstd, fmaf(-ln_mu, c4.y, __uint_as_float(r[j + 1])), b4.y);
            v2 = fmaf(ln_rstd, fmaf(-ln_mu, c4.z, __uint_as_float(r[j + 2])), b4.z);
            v3 = fmaf(ln_rstd, fmaf(-ln_mu, c4.w, __uint_as_float(r[j + 3])), b4.w);
          } else {
            v0 = __uint_as_float(r[j + 0]) + b4.x;
            v1 = __uint_as_float(r[j + 1]) + b4.y;
            v2 = __uint_as_float(r[j + 2]) + b4.z;
            v3 = __uint_as_float(r[j + 3]) + b4.w;
          }
          if constexpr (EPI == DFD_EPI_STORE_BF16_QGELU || EPI == DFD_EPI_STORE_BF16_QGELU_LNFOLD) {
            v0 = quick_gelu_fast(v0);
            v1 = quick_gelu_fast(v1);
            v2 = quick_gelu_fast(v2);
            v3 = quick_gelu_fast(v3);
          } else if constexpr (EPI == DFD_EPI_STORE_BF16_GELU) {
            v0 = gelu_erf(v0);
            v1 = gelu_erf(v1);
            v2 = gelu_erf(v2);
            v3 = gelu_erf(v3);
          }
          packed[half * 16 + j / 2 + 0] = pack_bf16(v0, v1);
          packed[half * 16 + j / 2 + 1] = pack_bf16(v2, v3);
        }
      }
    }
    if (box == box_begin + box_count - 1) release();  // all TMEM reads of this accumulator are done in this warp
    // staging buffer `buf` was last read by the TMA store issued two boxes ago
    if (lane == 0) tma_store_wait_read<OUT_BUFS_PER_WARP - 1>();
    __syncwarp();
    uint8_t* dst = obuf + buf * OUT_BUF + lane * 128;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      // 128-byte swizzle: 16-byte chunk index XOR (row & 7); box base is 1024-byte aligned
      uint4 v = make_uint4(packed[ch * 4 + 0], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
      *reinterpret_cast<uint4*>(dst + ((ch ^ (lane & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && row0 < M) {
      if constexpr (EPI == DFD_EPI_ADD_F32 || EPI == DFD_EPI_ADD_BF16)
        tma_reduce_add_2d(tmC, obuf + buf * OUT_BUF, c, row0);
      else
        tma_store_2d(tmC, obuf + buf * OUT_BUF, c, row0);
    }
    if (lane == 0) tma_store_commit();
    buf ^= 1;
  }
}

template <int EPI>
__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int M, int N, int K) {
  using namespace gemm;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);

  // warp-uniform values go through a lane-0 shuffle: the compiler then keeps what is derived from them (TMEM
  // addresses, tile indices) in uniform registers, which the tcgen05 / TMA instructions take directly
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], EPI_WARPS * 32);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_STAGE + B_STAGE);
          tma_load_2d(&tmA, &full_bar[stage], smem + OFF_A + stage * A_STAGE, kb * BK, m_blk * BM, kEvictNormal);
          tma_load_2d(&tmB, &full_bar[stage], smem + OFF_B + stage * B_STAGE, kb * BK, n_blk * BN, kEvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (uniform control flow, every lane polls the barriers) and ONE elected lane issues
    // the tcgen05 instructions: under `if (lane == 0)` every operand lived in a vector register and each MMA paid an
    // ELECT / R2UR.BROADCAST x 4 / retry-branch sequence (~95 cycles per issue, measured on the attention kernel).
    {
      const bool elected = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem + OFF_A + stage * A_STAGE);
          const uint64_t b_desc = umma_desc_sw128(smem + OFF_B + stage * B_STAGE);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance both descriptors by k*32 bytes inside the 128-byte swizzle atom
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elected) umma_commit(&tmem_full[acc]);  // accumulator complete
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may read: lanes [32q, 32q+32)
    const int ew = warp - 2;
    uint8_t* obuf = smem + OFF_OUT + ew * (OUT_BUFS_PER_WARP * OUT_BUF);
    constexpr bool kOutF32 = (EPI == DFD_EPI_STORE_F32 || EPI == DFD_EPI_ADD_F32);
    constexpr int NUM_BOX = BN / (kOutF32 ? 32 : 64);
    static_assert(OUT_BUF == gemm_epi::OUT_BUF && OUT_BUFS_PER_WARP == gemm_epi::OUT_BUFS_PER_WARP, "epilogue staging");
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row0 = m_blk * BM + q * 32;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      epilogue_boxes<EPI, false>(t_row, n_blk * BN, row0, M, 0, NUM_BOX, bias, nullptr, 0.f, 0.f, obuf, buf, &tmC, [&]() {
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);  // hand the accumulator back to the MMA warp (128 arrivals)
      });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// =====================================================================================================
// 2-SM variant (tcgen05 cta_group::2): a cluster of two CTAs (one SM pair) computes a 256 x 256 tile.
// Each CTA loads its own 128 rows of A and its own 128-row half of the W tile (32 KB per stage instead of
// 48 KB: one third less L2->SM traffic and one third less shared-memory read traffic per FLOP), the leader
// CTA's MMA thread issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16) which reads both CTAs' shared
// memory and writes rows [0,128) of the accumulator to the leader's TMEM and rows [128,256) to the peer's.
// Each CTA drains its own TMEM with the same epilogue as the 1-SM kernel.
// Cross-CTA signalling: TMA loads of both CTAs complete_tx on the LEADER's full barrier; tcgen05.commit
// multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps
// arrive remotely on the leader's "accumulator drained" barrier.
constexpr int DFD_EPI_RESID_LN_F32_X2 = 9;  // internal second instance of DFD_EPI_RESID_LN_F32 (see gemm2::Cfg)

namespace gemm2 {
constexpr int BM = 128;            // rows per CTA (256 per cluster)
constexpr int BN = 256;
constexpr int BNH = 128;           // W rows loaded by each CTA
constexpr int BK = 64;
constexpr int STAGES = 5;
constexpr int UMMA_K = 16;
constexpr int A_STAGE = BM * BK * 2;   // 16 KB
constexpr int B_STAGE = BNH * BK * 2;  // 16 KB
constexpr int OUT_BUF = 32 * 128;
constexpr int OUT_BUFS_PER_WARP = 2;
constexpr int EPI_WARPS = 8;      // two per TMEM lane quarter, each draining one 128-column half of the tile
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr uint32_t TMEM_COLS = 512;

// Shared-memory plan. The residual + LayerNorm-statistics epilogue stages fp32 boxes of x (TMA load, update in place,
// TMA store) and one bf16 box per epilogue warp, with one load barrier per fp32 box. Two instances, picked by K:
//   DFD_EPI_RESID_LN_F32      5-stage operand ring, ONE fp32 box: for long K (c_proj, K = 4D) the epilogue has slack
//                             but a 4-stage ring cost the mainloop 12 %;
//   DFD_EPI_RESID_LN_F32_X2   4-stage ring, TWO fp32 boxes (the next box of x is in flight while one is updated): for
//                             short K (out_proj) the epilogue is the critical path (single box: 245 vs 208 us).
template <int EPI>
struct Cfg {
  static constexpr bool kResidLn = (EPI == DFD_EPI_RESID_LN_F32 || EPI == DFD_EPI_RESID_LN_F32_X2);
  static constexpr int NX = (EPI == DFD_EPI_RESID_LN_F32_X2) ? 2 : 1;   // fp32 boxes of x per epilogue warp
  static constexpr int STAGES = (EPI == DFD_EPI_RESID_LN_F32_X2) ? 4 : gemm2::STAGES;
  static constexpr int OUT_BUFS_PER_WARP = kResidLn ? NX + 1 : gemm2::OUT_BUFS_PER_WARP;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
  static constexpr int OFF_OUT = OFF_B + STAGES * B_STAGE;
  static constexpr int OFF_BAR = OFF_OUT + EPI_WARPS * OUT_BUFS_PER_WARP * OUT_BUF;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + (kResidLn ? 2 * EPI_WARPS : 0);  // + x-load barriers
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "2-SM GEMM shared memory exceeds the per-CTA limit");
};
}  // namespace gemm2

// Arguments of the LayerNorm-folding epilogues (device pointers; see dfd_gemm_ln_args in the header).
struct GemmLnArgs {
  const float* stats_in = nullptr;  // [M, slots, 2] partial (sum, sum of squares) of the rows whose bf16 copy is A
  const float* colsum = nullptr;    // [N] c[n] = sum_k W'[n, k]
  float* stats_out = nullptr;       // RESID_LN: [M, 2 * N / 256, 2]
  const float* x = nullptr;         // RESID_LN: the fp32 rows being updated (= out), read with plain loads
  int64_t ldx = 0;
  int slots = 0;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0,
                                                int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm2::THREADS, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmU,
                     const float* __restrict__ bias, const GemmLnArgs ln, int M, int N, int K) {
  using namespace gemm2;
  using C = Cfg<EPI>;
  constexpr int STAGES = C::STAGES, OUT_BUFS_PER_WARP = C::OUT_BUFS_PER_WARP;
  constexpr int OFF_A = C::OFF_A, OFF_B = C::OFF_B, OFF_OUT = C::OFF_OUT, OFF_BAR = C::OFF_BAR,
                OFF_TMEM_PTR = C::OFF_TMEM_PTR;
  constexpr bool kLnFold = (EPI == DFD_EPI_STORE_BF16_LNFOLD || EPI == DFD_EPI_STORE_BF16_QGELU_LNFOLD);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);

  // warp-uniform values go through a lane-0 shuffle: the compiler then keeps what is derived from them (TMEM
  // addresses, tile indices) in uniform registers, which the tcgen05 / TMA instructions take directly
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool leader = rank == 0;

  const int num_m2 = (M + 2 * BM - 1) / (2 * BM);
  const int num_n = N / BN;
  const int num_tiles = num_m2 * num_n;
  const int num_kb = (K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (C::kResidLn) tma_prefetch_desc(&tmU);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 2 * EPI_WARPS);  // one elected arrival per epilogue warp of both CTAs
      }
      if constexpr (C::kResidLn)
        for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&tmem_empty[2 + i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int a_row = m_blk * 2 * BM + static_cast<int>(rank) * BM;
        const int b_row = n_blk * BN + static_cast<int>(rank) * BNH;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE + B_STAGE));
          const uint32_t bar = mapa_u32(&full_bar[stage], 0);
          tma_load_2d_2sm(&tmA, bar, smem + OFF_A + stage * A_STAGE, kb * BK, a_row, kEvictNormal);
          tma_load_2d_2sm(&tmB, bar, smem + OFF_B + stage * B_STAGE, kb * BK, b_row, kEvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // whole warp in the loop, one elected lane issues (see the 1-SM kernel): operands stay in uniform registers
    if (leader) {
      const bool elected = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem + OFF_A + stage * A_STAGE);
          const uint64_t b_desc = umma_desc_sw128(smem + OFF_B + stage * B_STAGE);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);  // both CTAs' stage reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elected) umma_commit_2sm(&tmem_full[acc]);  // accumulator complete in both CTAs' TMEM
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    const int ew = warp - 2;
    uint8_t* obuf = smem + OFF_OUT + ew * (OUT_BUFS_PER_WARP * OUT_BUF);
    if constexpr (C::kResidLn) {
      // ---- x = x + acc + bias with the FULL value in hand (instead of a blind reduce-add), so that the same pass
      // also emits the bf16 copy of x (the A operand of the next, LayerNorm-folded GEMM) and each row's partial
      // (sum, sum of squares) over this warp's 128 columns. LayerNorm itself then costs no pass over x at all:
      // LN(x) W^T = rstd (x (gamma.W)^T - mu c) + (b + W beta) is finished in the consumer GEMM's epilogue.
      const int half = ew >> 2;
      constexpr int NX = C::NX;
      uint8_t* xbuf = obuf;                   // NX fp32 boxes (32 rows x 32 cols): TMA load of x, update in place, TMA store
      uint8_t* ubuf = obuf + NX * OUT_BUF;    // bf16 staging box (32 rows x 64 cols), filled by two fp32 boxes
      uint64_t* lbar = tmem_empty + 2 + 2 * ew;
      uint32_t lphase = 0;                    // bit i = parity of lbar[i]
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int row0 = m_blk * 2 * BM + static_cast<int>(rank) * BM + q * 32;
        const int row = row0 + lane;
        const bool live = row0 < M;          // a warp whose 32 rows lie beyond M only drains its accumulator
        const int col_base = n_blk * BN + half * 128;
        // the first box(es) of x are fetched while the MMAs of this tile are still running
        if (lane == 0 && live) {
          tma_store_wait_read<0>();          // the previous tile's last stores have left shared memory
#pragma unroll
          for (int b = 0; b < NX; ++b) {
            mbar_arrive_expect_tx(&lbar[b], OUT_BUF);
            tma_load_2d(&tmC, &lbar[b], xbuf + b * OUT_BUF, col_base + b * 32, row0, kEvictFirst);
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * 128;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xi = b % NX;
          const int c = col_base + b * 32;
          uint32_t r[32];
          tmem_ld32(t_row + b * 32, r);
          uint8_t* xdst = xbuf + xi * OUT_BUF + lane * 128;
          uint8_t* udst = ubuf + lane * 128;
          float v[32];
          if (live) {
            mbar_wait(&lbar[xi], (lphase >> xi) & 1u);
            lphase ^= 1u << xi;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const float4 xv = *reinterpret_cast<const float4*>(xdst + ((ch ^ (lane & 7)) << 4));
              v[4 * ch] = xv.x; v[4 * ch + 1] = xv.y; v[4 * ch + 2] = xv.z; v[4 * ch + 3] = xv.w;
            }
          }
          tmem_ld_wait();
          if (b == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));
          }
          if (live) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c) + j) : make_float4(0, 0, 0, 0);
              v[4 * j + 0] += __uint_as_float(r[4 * j + 0]) + b4.x;
              v[4 * j + 1] += __uint_as_float(r[4 * j + 1]) + b4.y;
              v[4 * j + 2] += __uint_as_float(r[4 * j + 2]) + b4.z;
              v[4 * j + 3] += __uint_as_float(r[4 * j + 3]) + b4.w;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s1 += v[j];
              s2 = fmaf(v[j], v[j], s2);
            }
            // lane 0 has waited for every earlier store group before it issued a refill: past this point ubuf (last
            // stored one box ago) and this fp32 box are free
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 f = make_uint4(__float_as_uint(v[4 * ch]), __float_as_uint(v[4 * ch + 1]),
                                         __float_as_uint(v[4 * ch + 2]), __float_as_uint(v[4 * ch + 3]));
              *reinterpret_cast<uint4*>(xdst + ((ch ^ (lane & 7)) << 4)) = f;
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const uint4 h = make_uint4(pack_bf16(v[8 * ch], v[8 * ch + 1]), pack_bf16(v[8 * ch + 2], v[8 * ch + 3]),
                                         pack_bf16(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16(v[8 * ch + 6], v[8 * ch + 7]));
              *reinterpret_cast<uint4*>(udst + ((((b & 1) * 4 + ch) ^ (lane & 7)) << 4)) = h;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, xbuf + xi * OUT_BUF, c, row0);
              if (b & 1) tma_store_2d(&tmU, ubuf, col_base + (b >> 1) * 64, row0);
              tma_store_commit();
              // refill this fp32 box with box b + NX once its store has read it; with two boxes the wait after the odd
              // box is still needed: it frees ubuf for the next even box
              if (b + NX < 4 || (NX == 2 && b == 1)) tma_store_wait_read<0>();
              if (b + NX < 4) {
                mbar_arrive_expect_tx(&lbar[xi], OUT_BUF);
                tma_load_2d(&tmC, &lbar[xi], xbuf + xi * OUT_BUF, c + NX * 32, row0, kEvictFirst);
              }
            }
          }
        }
        if (live && row < M)
          *reinterpret_cast<float2*>(ln.stats_out + (static_cast<int64_t>(row) * ln.slots + n_blk * 2 + half) * 2) =
              make_float2(s1, s2);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (lane == 0) tma_store_wait_all<0>();
    } else {
    constexpr bool kOutF32 = (EPI == DFD_EPI_STORE_F32 || EPI == DFD_EPI_ADD_F32);
    constexpr int NUM_BOX = BN / (kOutF32 ? 32 : 64);
    constexpr int BOX_PER_WARP = NUM_BOX / 2;
    static_assert(OUT_BUF == gemm_epi::OUT_BUF && OUT_BUFS_PER_WARP == gemm_epi::OUT_BUFS_PER_WARP, "epilogue staging");
    const int box_begin = (ew >> 2) * BOX_PER_WARP;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    float2 ln_pref[8];  // (sum, sum of squares) partials of one row, at most 8 blocks of 128 columns (D <= 1024)
    auto load_ln_partials = [&](int row) {
#pragma unroll
      for (int sl = 0; sl < 8; ++sl)
        ln_pref[sl] = (kLnFold && row < M && sl < ln.slots)
                          ? reinterpret_cast<const float2*>(ln.stats_in)[static_cast<int64_t>(row) * ln.slots + sl]
                          : make_float2(0.f, 0.f);
    };
    if constexpr (kLnFold) {
      if (cluster_id < num_tiles)
        load_ln_partials((cluster_id / num_n) * 2 * BM + static_cast<int>(rank) * BM + q * 32 + lane);
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row0 = m_blk * 2 * BM + static_cast<int>(rank) * BM + q * 32;
      const int col0 = n_blk * BN;
      // LayerNorm folded into this GEMM: A is the bf16 copy of the un-normalised rows and W carries gamma, so
      // out = rstd * (acc - mu * colsum[n]) + bias'[n]; mu / rstd come from the producer's per-row partial sums.
      float ln_mu = 0.f, ln_rstd = 0.f;
      if constexpr (kLnFold) {
        // the partial sums of this tile's row were requested one tile ago (ln_pref); request the next tile's now, so
        // that the dependent global loads never sit on the epilogue's critical path
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          a1 += ln_pref[sl].x;
          a2 += ln_pref[sl].y;
        }
        const float inv_k = 1.0f / static_cast<float>(K);
        ln_mu = a1 * inv_k;
        ln_rstd = rsqrtf(fmaxf(a2 * inv_k - ln_mu * ln_mu, 0.f) + 1e-5f);
        const int next = tile + num_clusters;
        if (next < num_tiles) load_ln_partials((next / num_n) * 2 * BM + static_cast<int>(rank) * BM + q * 32 + lane);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      epilogue_boxes<EPI, kLnFold>(t_row, col0, row0, M, box_begin, BOX_PER_WARP, bias, ln.colsum, ln_mu, ln_rstd, obuf,
                                   buf, &tmC, [&]() {
                                     // tell the leader's MMA thread (one elected arrival per epilogue warp)
                                     tc_fence_before();
                                     __syncwarp();
                                     if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));
                                   });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
    }  // generic epilogues
  }

  // no CTA of the pair may exit (or free TMEM) while the other can still signal it or read its shared memory
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// --------------------------------------------------------------------------------------------- host side
template <int EPI>
static int launch(const dfd_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                  const float* bias, int M, int N, int K, cudaStream_t stream) {
  using namespace gemm;
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int num_tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = num_tiles < ctx->num_sms ? num_tiles : ctx->num_sms;
  gemm_bf16_kernel<EPI><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, bias, M, N, K);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int EPI>
static int launch2(const dfd_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                   const float* bias, int M, int N, int K, cudaStream_t stream, const CUtensorMap* tmU = nullptr,
                   const GemmLnArgs& ln = GemmLnArgs()) {
  using namespace gemm2;
  constexpr int SMEM = Cfg<EPI>::SMEM_BYTES;
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_2sm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured[ctx->device & 63] = true;
  }
  const int num_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / BN);
  const int max_clusters = ctx->num_sms / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  gemm_bf16_2sm_kernel<EPI><<<2 * clusters, THREADS, SMEM, stream>>>(tmA, tmB, tmC, tmU ? *tmU : tmC, bias, ln, M, N, K);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// DFD_GEMM_2SM=0 selects the 1-SM kernel (A/B comparisons); default: SM-pair kernel whenever there are >= 2 row blocks.
static bool use_2sm(int M) {
  static const int mode = []() {
    const char* e = getenv("DFD_GEMM_2SM");
    return e ? atoi(e) : 1;
  }();
  return mode != 0 && M > gemm::BM;
}

int gemm_bf16(const dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
              void* out, int64_t ldo, int M, int N, int K, int epilogue, cudaStream_t stream) {
  using namespace gemm;
  DFD_CHECK_ARG(ctx && A && W && out, "gemm: null pointer");
  DFD_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  DFD_CHECK_ARG(N % BN == 0, "gemm: N=%d must be a multiple of %d", N, BN);
  DFD_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16-byte rows)");
  DFD_CHECK_ARG(lda >= K && ldw >= K && ldo >= N, "gemm: leading dimension smaller than the row length");
  const bool f32 = (epilogue == DFD_EPI_STORE_F32 || epilogue == DFD_EPI_ADD_F32);
  DFD_CHECK_ARG(ldo % (f32 ? 4 : 8) == 0, "gemm: ldo must give 16-byte aligned rows");
  DFD_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "gemm: operands must be 16-byte aligned");
  DFD_CHECK_ARG(bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0, "gemm: bias must be 16-byte aligned");

  CUtensorMap tmA, tmB, tmC;
  DFD_TRY(make_tmap_2d(ctx, &tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK));
  const bool two_sm = use_2sm(M);
  DFD_TRY(make_tmap_2d(ctx, &tmB, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldw, two_sm ? gemm2::BNH : BN, BK));
  if (f32)
    DFD_TRY(make_tmap_2d(ctx, &tmC, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldo, 32, 32));
  else
    DFD_TRY(make_tmap_2d(ctx, &tmC, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldo, 32, 64));

  if (two_sm) {
    switch (epilogue) {
      case DFD_EPI_STORE_BF16:
        return launch2<DFD_EPI_STORE_BF16>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      case DFD_EPI_STORE_BF16_QGELU:
        return launch2<DFD_EPI_STORE_BF16_QGELU>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      case DFD_EPI_STORE_F32:
        return launch2<DFD_EPI_STORE_F32>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      case DFD_EPI_ADD_F32:
        return launch2<DFD_EPI_ADD_F32>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      case DFD_EPI_ADD_BF16:
        return launch2<DFD_EPI_ADD_BF16>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      case DFD_EPI_STORE_BF16_GELU:
        return launch2<DFD_EPI_STORE_BF16_GELU>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
      default:
        return fail(DFD_ERR_INVALID, "gemm: unknown epilogue %d", epilogue);
    }
  }
  switch (epilogue) {
    case DFD_EPI_STORE_BF16:
      return launch<DFD_EPI_STORE_BF16>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    case DFD_EPI_STORE_BF16_QGELU:
      return launch<DFD_EPI_STORE_BF16_QGELU>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    case DFD_EPI_STORE_F32:
      return launch<DFD_EPI_STORE_F32>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    case DFD_EPI_ADD_F32:
      return launch<DFD_EPI_ADD_F32>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    case DFD_EPI_ADD_BF16:
      return launch<DFD_EPI_ADD_BF16>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    case DFD_EPI_STORE_BF16_GELU:
      return launch<DFD_EPI_STORE_BF16_GELU>(ctx, tmA, tmB, tmC, bias, M, N, K, stream);
    default:
      return fail(DFD_ERR_INVALID, "gemm: unknown epilogue %d", epilogue);
  }
}

// GEMMs with LayerNorm folded around them (2-SM kernel only, M > 128):
//   DFD_EPI_RESID_LN_F32            out_f32 (= ln->x) += acc + bias; also writes bf16(out) and per-row partial statistics
//   DFD_EPI_STORE_BF16[_QGELU]_LNFOLD  out_bf16 = [quickgelu] (rstd * (acc - mu * colsum[n]) + bias[n])
int gemm_bf16_ln(const dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                 void* out, int64_t ldo, int M, int N, int K, int epilogue, const dfd_gemm_ln_args* a,
                 cudaStream_t stream) {
  using namespace gemm;
  DFD_CHECK_ARG(ctx && A && W && out && a, "gemm_ln: null pointer");
  DFD_CHECK_ARG(M > 0, "gemm_ln: empty problem");
  DFD_CHECK_ARG(N > 0 && K > 0 && N % BN == 0, "gemm_ln: N=%d must be a positive multiple of %d", N, BN);
  DFD_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K && ldo >= N,
                "gemm_ln: bad K / leading dimensions");
  DFD_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "gemm_ln: operands must be 16-byte aligned");
  DFD_CHECK_ARG(bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0, "gemm_ln: bias must be 16-byte aligned");
  CUtensorMap tmA, tmB, tmC, tmU;
  DFD_TRY(make_tmap_2d(ctx, &tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK));
  DFD_TRY(make_tmap_2d(ctx, &tmB, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldw, gemm2::BNH, BK));
  GemmLnArgs ln;
  if (epilogue == DFD_EPI_RESID_LN_F32) {
    DFD_CHECK_ARG(a->stats_out && a->bf16_out, "gemm_ln: RESID_LN needs stats_out and bf16_out");
    DFD_CHECK_ARG(ldo % 4 == 0 && a->ld_bf16 % 8 == 0 && a->ld_bf16 >= N &&
                      reinterpret_cast<uintptr_t>(a->bf16_out) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(a->stats_out) % 8 == 0,
                  "gemm_ln: RESID_LN alignment");
    DFD_TRY(make_tmap_2d(ctx, &tmC, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldo, 32, 32));
    DFD_TRY(make_tmap_2d(ctx, &tmU, a->bf16_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, a->ld_bf16, 32, 64));
    ln.stats_out = a->stats_out;
    ln.x = static_cast<const float*>(out);
    ln.ldx = ldo;
    ln.slots = 2 * (N / BN);
    if (K <= 1024) return launch2<DFD_EPI_RESID_LN_F32_X2>(ctx, tmA, tmB, tmC, bias, M, N, K, stream, &tmU, ln);
    return launch2<DFD_EPI_RESID_LN_F32>(ctx, tmA, tmB, tmC, bias, M, N, K, stream, &tmU, ln);
  }
  DFD_CHECK_ARG(epilogue == DFD_EPI_STORE_BF16_LNFOLD || epilogue == DFD_EPI_STORE_BF16_QGELU_LNFOLD,
                "gemm_ln: unknown epilogue %d", epilogue);
  DFD_CHECK_ARG(a->stats_in && a->colsum && a->slots > 0 && a->slots <= 8,
                "gemm_ln: LNFOLD needs stats_in, colsum and 1..8 slots");
  DFD_CHECK_ARG(ldo % 8 == 0 && reinterpret_cast<uintptr_t>(a->colsum) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(a->stats_in) % 8 == 0,
                "gemm_ln: LNFOLD alignment");
  DFD_TRY(make_tmap_2d(ctx, &tmC, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldo, 32, 64));
  ln.stats_in = a->stats_in;
  ln.colsum = a->colsum;
  ln.slots = a->slots;
  if (epilogue == DFD_EPI_STORE_BF16_LNFOLD)
    return launch2<DFD_EPI_STORE_BF16_LNFOLD>(ctx, tmA, tmB, tmC, bias, M, N, K, stream, nullptr, ln);
  return launch2<DFD_EPI_STORE_BF16_QGELU_LNFOLD>(ctx, tmA, tmB, tmC, bias, M, N, K, stream, nullptr, ln);
}

}  // namespace dfd

extern "C" int dfd_gemm_bf16_ln(dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw,
                                const float* bias, void* out, int64_t ldo, int M, int N, int K, int epilogue,
                                const dfd_gemm_ln_args* args, void* stream) {
  dfd::clear_error();
  return dfd::gemm_bf16_ln(ctx, A, lda, W, ldw, bias, out, ldo, M, N, K, epilogue, args,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int dfd_gemm_bf16(dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             void* out, int64_t ldo, int M, int N, int K, int epilogue, void* stream) {
  dfd::clear_error();
  return dfd::gemm_bf16(ctx, A, lda, W, ldw, bias, out, ldo, M, N, K, epilogue, static_cast<cudaStream_t>(stream));
}

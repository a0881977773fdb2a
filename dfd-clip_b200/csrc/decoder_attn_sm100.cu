// Decoder cross-attention (src/models.py:99-146, pos-emb add :326-329, mask :324), streaming version.
//
//   mix[b,h,:] = sum_s 1/2 (a0 + a1)_{s,h} (V_s + pe_t)          s = (frame t, patch p), one query per clip
//   a0 = softmax_s(q0.(K_s + pe_t) / 8)          (masked frames: -inf)
//   a1 = tanh(q1.(K_s + pe_t) / 8) * 2 sigmoid(-|q1 - K_s - pe_t|_1 / 8)      (masked frames: 0)
//
// HBM-bound: every K and V row of the tapped layer (bf16, read in place from the encoder's packed QKV buffer through
// strides) is needed exactly once: 2 * S * D * 2 bytes per clip and block. The kernel is a persistent producer /
// consumer pipeline, one CTA per SM:
//   * work unit = (clip, frame, half of the frame's patches); units are handed out round-robin to the CTAs
//     (B*T*2 units over 148 SMs: ~1 % imbalance at 64 clips x 8 frames);
//   * warp 0 streams the unit's token rows into a 3-slot shared-memory ring with bulk async copies
//     (cp.async.bulk, one K row and one V row of H*128 bytes per token, 4*KS tokens per slot, mbarrier expect_tx),
//     so ~130 KB per SM are in flight independently of what the compute warps are doing;
//   * HG*KS consumer warps (HG = H/4 head groups, KS key subsets). A warp works on a PAIR of tokens at a time:
//     lane = (token of the pair, head within the group, channel quarter), 16 channels per lane, so the three
//     per-(token, head) scalars need only two shuffle rounds and the scalar softmax / CoDA math is replicated 4x
//     instead of 8x. K/V pieces come from shared memory with conflict-free 16-byte loads; dot products, the L1
//     distance and the accumulator updates use packed fp32x2 instructions (FFMA2) on (softmax, coda) pairs;
//     online softmax + plain CoDA accumulation in registers;
//   * the temporal position embedding never touches the inner loop: q.(K+pe) = q.K + q.pe (a per-unit constant),
//     |q1 - K - pe| = |(q1 - pe) - K|, and sum a (V + pe) = sum a V + (sum a) pe is applied once per unit;
//   * per unit the KS key subsets are merged through shared memory into one (m, l, acc0[64], acc1[64]) record per
//     head; dec_attn_combine_kernel merges the records of a clip in a fixed order (deterministic, no atomics).
#include "common.cuh"
#include "host_common.h"

namespace dfd {

constexpr int DAS_REC = 130;   // m, l, acc0[64], acc1[64] — same record as dec_attn_combine_kernel reads
constexpr int DAS_SLOTS = 3;
constexpr int DAS_SPLIT = 2;   // units per (clip, frame)

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void bf16x8_unpack(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i + 0] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ float tanh_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int H, int KS>
struct DasCfg {
  static constexpr int HG = H / 4;
  static constexpr int TOK = 4 * KS;                 // tokens per ring slot: 2 token pairs per consumer warp
  static constexpr int ROW = 2 * H * 128;            // bytes per token in the ring: K row then V row
  static constexpr int SLOT = TOK * ROW;
  static constexpr int CONSUMERS = HG * KS;          // warps
  static constexpr int THREADS = 32 * (1 + CONSUMERS);
  static constexpr int RING_BYTES = DAS_SLOTS * SLOT;
  static constexpr int COMB_BYTES = 2 * KS * H * DAS_REC * 4;  // double-buffered subset records
  static constexpr int BAR_OFF = RING_BYTES + COMB_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 2 * DAS_SLOTS * 8 + 128;
};

template <int H, int KS>
__global__ void __launch_bounds__(DasCfg<H, KS>::THREADS, 1)
dec_attn_stream_kernel(const float* __restrict__ qs, const __nv_bfloat16* __restrict__ kbase,
                       const __nv_bfloat16* __restrict__ vbase, int64_t stride_b, int64_t stride_t, int64_t stride_p,
                       const float* __restrict__ pos_emb, const uint8_t* __restrict__ mask, int T, int P, int num_units,
                       int kv_contig, float* __restrict__ part) {
  using C = DasCfg<H, KS>;
  extern __shared__ __align__(128) uint8_t das_smem[];
  uint8_t* ring = das_smem;
  float* comb = reinterpret_cast<float*>(das_smem + C::RING_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(das_smem + C::BAR_OFF);
  uint64_t* empty = full + DAS_SLOTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p_half = (P + DAS_SPLIT - 1) / DAS_SPLIT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DAS_SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C::CONSUMERS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ---------------------------------------------------------------------------------- producer
    // lane j issues the copies of token j of the slot (one K row + one V row, or one copy when V follows K in memory)
    int slot = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int bt = u / DAS_SPLIT, half = u % DAS_SPLIT;
      if (mask[bt] == 0) continue;
      const int b = bt / T, t = bt % T;
      const int p_beg = half * p_half, p_end = min(P, p_beg + p_half);
      const int64_t off = b * stride_b + t * stride_t;
      for (int p0 = p_beg; p0 < p_end; p0 += C::TOK) {
        const int n = min(C::TOK, p_end - p0);
        mbar_wait(&empty[slot], phase ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(n) * C::ROW);
        __syncwarp();
        if (lane < n) {
          uint8_t* dst = ring + slot * C::SLOT + lane * C::ROW;
          const int64_t o = off + static_cast<int64_t>(p0 + lane) * stride_p;
          if (kv_contig) {
            bulk_g2s(dst, kbase + o, 2 * H * 128, &full[slot]);
          } else {
            bulk_g2s(dst, kbase + o, H * 128, &full[slot]);
            bulk_g2s(dst + H * 128, vbase + o, H * 128, &full[slot]);
          }
        }
        if (++slot == DAS_SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------------------------- consumers
    const int cw = warp - 1;
    const int hg = cw % C::HG, ks = cw / C::HG;
    const int tsel = lane >> 4;                         // which token of the pair
    const int head = hg * 4 + ((lane >> 2) & 3);
    const int qd = lane & 3;                            // channels 8qd..8qd+7 and 32+8qd..32+8qd+7
    const float kLog2e = 1.4426950408889634f;
    // a quarter-warp (two heads x four channel quarters) must cover all 32 banks: odd heads take the upper 64 bytes
    // of their 128-byte row with the first load, even heads the lower 64 bytes
    const int hsw = (lane >> 2) & 1;
    auto chan = [&](int e) { return ((e < 8) != (hsw != 0) ? 0 : 32) + 8 * qd + (e & 7); };
    int slot = 0;
    uint32_t phase = 0;
    int parity = 0;  // combine buffer of this unit
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int bt = u / DAS_SPLIT, half = u % DAS_SPLIT;
      float* out = part + static_cast<int64_t>(u) * H * DAS_REC;
      if (mask[bt] == 0) {
        // frame absent: neutral element of the combine (m = -inf, l = 0, acc = 0)
        for (int i = threadIdx.x - 32; i < H * DAS_REC; i += C::CONSUMERS * 32)
          out[i] = (i % DAS_REC == 0) ? -INFINITY : 0.f;
        continue;
      }
      const int b = bt / T, t = bt % T;
      const int p_beg = half * p_half, p_end = min(P, p_beg + p_half);

      // per-unit constants, all as pairs over two neighbouring channels (one bf16x2 word of a K / V row): q0, q1,
      // q1 - pe, and the dot products of q0 / q1 with pe. Pairing over channels — not over (softmax, coda) — lets the
      // unpacked halves of a word, which sit in adjacent registers, feed the packed FMAs without any register moves.
      u64 q0p[8], q1p[8], q1mp[8];
      float c0 = 0.f, c1 = 0.f;
      const float* qh = qs + (static_cast<int64_t>(b) * H + head) * 128;
      const float* peh = pos_emb ? pos_emb + (static_cast<int64_t>(t) * H + head) * 64 : nullptr;
      {
        float q0v[16], q1v[16], q1m[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float a = qh[chan(e)], bq = qh[64 + chan(e)];
          const float pv = peh ? peh[chan(e)] : 0.f;
          q0v[e] = a;
          q1v[e] = bq;
          q1m[e] = bq - pv;
          c0 = fmaf(a, pv, c0);
          c1 = fmaf(bq, pv, c1);
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          q0p[e >> 1] = pack2(q0v[e], q0v[e + 1]);
          q1p[e >> 1] = pack2(q1v[e], q1v[e + 1]);
          q1mp[e >> 1] = pack2(q1m[e], q1m[e + 1]);
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          c0 += __shfl_xor_sync(0xffffffffu, c0, o);
          c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
      }
      u64 acc0p[8], acc1p[8];  // (acc[c], acc[c+1]) of the softmax / coda accumulators
#pragma unroll
      for (int w = 0; w < 8; ++w) acc0p[w] = acc1p[w] = 0ull;
      float m = -INFINITY, l = 0.f, a1sum = 0.f;  // m in the log2 domain
      const u64 neg1 = pack2(-1.f, -1.f);

      for (int p0 = p_beg; p0 < p_end; p0 += C::TOK) {
        mbar_wait(&full[slot], phase);
        const uint8_t* src = ring + slot * C::SLOT + hg * 512 + ((lane >> 2) & 3) * 128 + qd * 16;
        uint4 kraw[2][2], vraw[2][2];
        bool valid[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int tok = 2 * (ks + j * KS) + tsel;
          valid[j] = p0 + tok < p_end;
          const uint8_t* row = src + tok * C::ROW;
          if (valid[j]) {
            kraw[j][0] = *reinterpret_cast<const uint4*>(row + hsw * 64);
            kraw[j][1] = *reinterpret_cast<const uint4*>(row + 64 - hsw * 64);
            vraw[j][0] = *reinterpret_cast<const uint4*>(row + H * 128 + hsw * 64);
            vraw[j][1] = *reinterpret_cast<const uint4*>(row + H * 128 + 64 - hsw * 64);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);  // this warp's pieces of the slot are in registers
        if (++slot == DAS_SLOTS) { slot = 0; phase ^= 1; }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // a pair may hold one valid token only (odd token count): the other half-warp contributes nothing
          const bool ok = valid[j];
          u64 dp0 = 0ull, dp1 = 0ull;   // q0.k and q1.k, each as (even channels, odd channels) partial sums
          float l1a = 0.f, l1b = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const uint32_t word = ok ? ((w < 4) ? (&kraw[j][0].x)[w] : (&kraw[j][1].x)[w - 4]) : 0u;
            const u64 kk = pack2(__uint_as_float(word << 16), __uint_as_float(word & 0xffff0000u));
            dp0 = fma2(kk, q0p[w], dp0);
            dp1 = fma2(kk, q1p[w], dp1);
            float dlo, dhi;
            unpack2(fma2(kk, neg1, q1mp[w]), dlo, dhi);  // (q1 - pe) - k
            l1a += fabsf(dlo);
            l1b += fabsf(dhi);
          }
          float d0a, d0b, d1a, d1b;
          unpack2(dp0, d0a, d0b);
          unpack2(dp1, d1a, d1b);
          float d0s = d0a + d0b, d1s = d1a + d1b, l1s = l1a + l1b;
#pragma unroll
          for (int o = 1; o < 4; o <<= 1) {
            d0s += __shfl_xor_sync(0xffffffffu, d0s, o);
            d1s += __shfl_xor_sync(0xffffffffu, d1s, o);
            l1s += __shfl_xor_sync(0xffffffffu, l1s, o);
          }
          // softmax term (log2 domain): s0 = (q0.K + q0.pe)/8; an absent token keeps (m, l, acc) unchanged
          const float s0 = ok ? (d0s + c0) * (0.125f * kLog2e) : -INFINITY;
          const float mn = fmaxf(m, s0);
          const float resc = (mn == -INFINITY) ? 1.f : fast_exp2(m - mn);  // first key: exp2(-inf) = 0
          const float pr = ok ? fast_exp2(s0 - mn) : 0.f;
          m = mn;
          l = fmaf(l, resc, pr);
          // coda term: tanh((q1.K + q1.pe)/8) * 2 sigmoid(-y) = 2 / (1 + e^y), y = |q1 - pe - K|_1 / 8
          const float gate = __fdividef(2.f, 1.f + fast_exp2(l1s * (0.125f * kLog2e)));
          const float a1 = ok ? tanh_fast((d1s + c1) * 0.125f) * gate : 0.f;
          a1sum += a1;
          const u64 prp = pack2(pr, pr), a1p = pack2(a1, a1), rp = pack2(resc, resc);
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const uint32_t word = ok ? ((w < 4) ? (&vraw[j][0].x)[w] : (&vraw[j][1].x)[w - 4]) : 0u;
            const u64 vv = pack2(__uint_as_float(word << 16), __uint_as_float(word & 0xffff0000u));
            acc0p[w] = fma2(acc0p[w], rp, mul2(vv, prp));   // online softmax: rescale, then add p * v
            acc1p[w] = fma2(vv, a1p, acc1p[w]);             // coda: plain accumulation
          }
        }
      }
      // ---- fold the two tokens of the pair (lanes l and l^16 hold the same head and channels)
      float acc0[16], acc1[16];
      {
        const float m_o = __shfl_xor_sync(0xffffffffu, m, 16), l_o = __shfl_xor_sync(0xffffffffu, l, 16);
        const float mn = fmaxf(m, m_o);
        const float sa = (m == -INFINITY) ? 0.f : fast_exp2(m - mn);
        const float sb = (m_o == -INFINITY) ? 0.f : fast_exp2(m_o - mn);
        l = l * sa + l_o * sb;
        m = mn;
        a1sum += __shfl_xor_sync(0xffffffffu, a1sum, 16);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          float x0, x1, y0, y1;
          unpack2(acc0p[w], x0, x1);
          unpack2(acc1p[w], y0, y1);
          acc0[2 * w] = x0 * sa + __shfl_xor_sync(0xffffffffu, x0, 16) * sb;
          acc0[2 * w + 1] = x1 * sa + __shfl_xor_sync(0xffffffffu, x1, 16) * sb;
          acc1[2 * w] = y0 + __shfl_xor_sync(0xffffffffu, y0, 16);
          acc1[2 * w + 1] = y1 + __shfl_xor_sync(0xffffffffu, y1, 16);
        }
      }
      // V + pe: sum a (V + pe) = sum a V + (sum a) pe
      if (peh) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float pv = peh[chan(e)];
          acc0[e] = fmaf(l, pv, acc0[e]);
          acc1[e] = fmaf(a1sum, pv, acc1[e]);
        }
      }
      // ---- merge the KS key subsets of this unit (natural-log domain record, as the combine kernel expects)
      float* cb = comb + parity * (KS * H * DAS_REC);
      if (tsel == 0) {
        float* rec = cb + (static_cast<int64_t>(ks) * H + head) * DAS_REC;
        if (qd == 0) {
          rec[0] = m;
          rec[1] = l;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          rec[2 + chan(e)] = acc0[e];
          rec[66 + chan(e)] = acc1[e];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::CONSUMERS * 32) : "memory");
      for (int i = threadIdx.x - 32; i < H * 64; i += C::CONSUMERS * 32) {
        const int hh = i >> 6, d = i & 63;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < KS; ++w) M = fmaxf(M, cb[(w * H + hh) * DAS_REC]);
        float Ls = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < KS; ++w) {
          const float* rec = cb + (w * H + hh) * DAS_REC;
          const float sc = (rec[0] == -INFINITY) ? 0.f : fast_exp2(rec[0] - M);
          Ls = fmaf(rec[1], sc, Ls);
          a0 = fmaf(rec[2 + d], sc, a0);
          a1 += rec[66 + d];
        }
        float* o = out + hh * DAS_REC;
        if (d == 0) {
          o[0] = M * 0.6931471805599453f;  // back to the natural-log domain used by the cross-unit combine
          o[1] = Ls;
        }
        o[2 + d] = a0;
        o[66 + d] = a1;
      }
      parity ^= 1;
    }
  }
}

size_t dec_attn_stream_workspace_bytes(int B, int T, int H) {
  return static_cast<size_t>(B) * T * DAS_SPLIT * H * DAS_REC * sizeof(float);
}

template <int H, int KS>
static int launch_das(const dfd_ctx* ctx, const float* qs, const __nv_bfloat16* k, const __nv_bfloat16* v,
                      int64_t stride_b, int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask,
                      int B, int T, int P, float* part, cudaStream_t stream) {
  using C = DasCfg<H, KS>;
  static std::atomic<bool> configured[64] = {};  // per device; a repeated cudaFuncSetAttribute is harmless
  if (!configured[ctx->device & 63]) {
    DFD_CUDA_OK(cudaFuncSetAttribute(dec_attn_stream_kernel<H, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int num_units = B * T * DAS_SPLIT;
  const int grid = num_units < ctx->num_sms ? num_units : ctx->num_sms;
  dec_attn_stream_kernel<H, KS><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(qs, k, v, stride_b, stride_t, stride_p,
                                                                            pos_emb, mask, T, P, num_units, (v == k + H * 64) ? 1 : 0, part);
  DFD_CUDA_OK(cudaGetLastError());
  return 0;
}

// Partial records [B*T*DAS_SPLIT][H][130] into `part`; returns the number of records per clip through *recs_per_clip.
int decoder_attention_stream(const dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                             int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                             int T, int P, int H, float* part, int* recs_per_clip, cudaStream_t stream) {
  const __nv_bfloat16* kb = static_cast<const __nv_bfloat16*>(k);
  const __nv_bfloat16* vb = static_cast<const __nv_bfloat16*>(v);
  *recs_per_clip = T * DAS_SPLIT;
  switch (H) {
    case 4: return launch_das<4, 4>(ctx, qs, kb, vb, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, part, stream);
    case 8: return launch_das<8, 4>(ctx, qs, kb, vb, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, part, stream);
    case 12: return launch_das<12, 4>(ctx, qs, kb, vb, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, part, stream);
    case 16: return launch_das<16, 3>(ctx, qs, kb, vb, stride_b, stride_t, stride_p, pos_emb, mask, B, T, P, part, stream);
    default: return fail(DFD_ERR_INVALID, "decoder_attention: heads=%d unsupported (need 4, 8, 12 or 16)", H);
  }
}

}  // namespace dfd

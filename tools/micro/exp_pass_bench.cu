// Microbenchmark of the attention kernels' exp2 pass at their occupancy (4 warps per SM = one per sub-partition):
// per 32-element chunk: scale FFMA, 2^x, row sum, bf16 pair packing — with NUM of every DEN exponentials on the
// FMA-pipe polynomial instead of MUFU.EX2. Rolled chunk loop (small code), so it isolates pipe contention from
// instruction-cache effects. Prints cycles per chunk per warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.0551716648f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <int NUM, int DEN, bool PACK>
__global__ void exp_pass(const float* __restrict__ in, uint32_t* out, int chunks, long long* cycles) {
  __shared__ float s_in[4][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 32 * 33; i += 32) s_in[warp][i] = in[i] - 3.f;
  __syncthreads();
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t acc = 0;
  const float sc = 0.18f, mo = 0.5f;
  long long t0 = clock64();
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    float v[32], e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = s_in[warp][j * 33 + ((lane + c) & 31)];  // stands for the tcgen05.ld of a chunk
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = fmaf(v[j], sc, -mo);
      e[j] = ((j % DEN) < NUM) ? poly_exp2(x) : fast_exp2(x);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      sum0 += e[j];
      sum1 += e[j + 1];
      if (PACK) acc ^= pack_bf16(e[j], e[j + 1]);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(sum0 + sum1);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Hand-interleaved schedule pinned with volatile asm: behind every MUFU.EX2 of chunk c come the scale FFMA of an element
// four ahead and the row-sum FADD / bf16 packing of the same position of chunk c-1, so a single warp keeps the XU pipe
// (8 cycles per warp-wide ex2) continuously busy instead of issuing ex2 in clumps.
#define V_EX2(d, a) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a))
#define V_FMA(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define V_ADD(d, a) asm volatile("add.f32 %0, %0, %1;" : "+f"(d) : "f"(a))
#define V_PACK(d, lo, hi) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo))

__global__ void exp_pass_interleaved(const float* __restrict__ in, uint32_t* out, int chunks, long long* cycles) {
  __shared__ float s_in[4][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 32 * 33; i += 32) s_in[warp][i] = in[i] - 3.f;
  __syncthreads();
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t acc = 0;
  const float sc = 0.18f, nmo = -0.5f;
  float ep[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) ep[j] = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    float v[32], x[32], e[32];
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = s_in[warp][j * 33 + ((lane + c) & 31)];
#pragma unroll
    for (int j = 0; j < 4; ++j) V_FMA(x[j], v[j], sc, nmo);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      V_EX2(e[j], x[j]);
      if (j + 4 < 32) V_FMA(x[j + 4], v[j + 4], sc, nmo);
      if (j & 1) {
        V_ADD(sum1, ep[j]);
        V_PACK(pk[j >> 1], ep[j - 1], ep[j]);
      } else {
        V_ADD(sum0, ep[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= pk[j];
#pragma unroll
    for (int j = 0; j < 32; ++j) ep[j] = e[j];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(sum0 + sum1);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NUM, int DEN, bool PACK>
void run(const char* name, const float* in, uint32_t* out, long long* cyc, int warps) {
  const int chunks = 4096;
  long long h[148];
  exp_pass<NUM, DEN, PACK><<<148, warps * 32>>>(in, out, chunks, cyc);
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%d  cycles per 32-element chunk per warp = %.1f\n", name, warps, (double)h[0] / chunks);
}

int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 32 * 33 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  float hin[32 * 33];
  for (int i = 0; i < 32 * 33; ++i) hin[i] = (float)(i % 17) * 0.3f;
  cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
  {
    const int chunks = 4096;
    long long h[148];
    exp_pass_interleaved<<<148, 128>>>(in, out, chunks, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM=4  cycles per 32-element chunk per warp = %.1f\n", "mufu + pack, hand-interleaved",
           (double)h[0] / chunks);
  }
  for (int warps : {4}) {
    run<0, 1, false>("mufu only, no pack", in, out, cyc, warps);
    run<0, 1, true>("mufu only + pack", in, out, cyc, warps);
    run<1, 4, true>("1/4 poly + pack", in, out, cyc, warps);
    run<1, 3, true>("1/3 poly + pack", in, out, cyc, warps);
    run<1, 2, true>("1/2 poly + pack", in, out, cyc, warps);
    run<2, 3, true>("2/3 poly + pack", in, out, cyc, warps);
    run<1, 1, true>("all poly + pack", in, out, cyc, warps);
  }
  return 0;
}

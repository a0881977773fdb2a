// Microbenchmark: MUFU.EX2 and LDTM (tcgen05.ld) throughput per SM on B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void mufu_kernel(float* out, int iters, long long* cycles) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#define EX(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
    EX(a0) EX(a1) EX(a2) EX(a3) EX(a4) EX(a5) EX(a6) EX(a7)
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void ffma_mufu_kernel(float* out, int iters, long long* cycles) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, s = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    // softmax-like: ffma, ex2, add, per element; 4 independent elements
    float e0, e1, e2, e3;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(a0, 0.18f, -1.f)));
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(a1, 0.18f, -1.f)));
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(a2, 0.18f, -1.f)));
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(a3, 0.18f, -1.f)));
    s += e0 + e1; s += e2 + e3;
    a0 += 1e-3f; a1 += 1e-3f; a2 += 1e-3f; a3 += 1e-3f;
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// TMEM load throughput: each warp reads its 32 lanes x 32 columns repeatedly
__global__ void ldtm_kernel(float* out, int iters, long long* cycles) {
  __shared__ uint32_t tptr;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t base = tptr + (((threadIdx.x >> 5) & 3) * 32 << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(base + (i & 7) * 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc ^= r[0] ^ r[31];
  }
  long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    mufu_kernel<<<148, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8 * warps * 32;
    printf("mufu  warps/SM=%2d  cycles=%lld  ex2 per clk per SM = %.2f\n", warps, h[0], ops / h[0]);
  }
  for (int warps : {2, 4, 8, 16}) {
    ffma_mufu_kernel<<<148, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = (double)iters * 4 * warps * 32;
    printf("ffma+mufu+add warps/SM=%2d  cycles=%lld  ex2 per clk per SM = %.2f\n", warps, h[0], ops / h[0]);
  }
  for (int warps : {1, 4, 8, 16}) {
    ldtm_kernel<<<148, warps * 32>>>(out, iters, cyc);
    cudaError_t e = cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)iters * warps * 32 * 32 * 4;
    printf("ldtm x32 warps/SM=%2d  cycles=%lld  bytes per clk per SM = %.1f (%s)\n", warps, h[0], bytes / h[0], cudaGetErrorString(e));
  }
  return 0;
}

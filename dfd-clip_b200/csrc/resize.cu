// Input side of the path on the GPU: the data loader's frame transform
//   T.Resize(n_px, BICUBIC) -> T.CenterCrop(n_px)            (src/models.py:756-761, applied to stacked uint8 frames at
//                                                             src/datasets.py:672-676)
// for uint8 frames of any size, so that raw decoded frames can cross PCIe and enter the encoder directly (the remaining
// two steps, ConvertImageDtype + Normalize, are fused into the patch extraction kernel, rowwise.cu).
//
// torchvision's tensor path evaluates the resize as separable ANTIALIASED bicubic interpolation in fp32 (ATen
// upsample_bicubic2d_aa: Keys cubic a = -0.5, support 2 * max(scale, 1), weights normalised per output pixel, width
// pass first, then height), clamps to [0, 255], rounds half-to-even and casts back to uint8. The kernels below follow
// that order of operations: a table kernel (first tap, tap count, normalised fp32 weights per output column / row), a
// horizontal pass over the cropped columns into an fp32 intermediate, a vertical pass with clamp + rint -> uint8.
// HBM-bound byte work: one read of the source frames, one fp32 intermediate of H x n_px per channel, coalesced stores.
#include "common.cuh"
#include "host_common.h"

namespace dfd {

struct ResizePlan {
  int H, W, R, new_h, new_w, top, left, taps_x, taps_y;
  float scale_x, scale_y;
  size_t off_xmin, off_xsize, off_xw, off_ymin, off_ysize, off_yw, off_tmp, total;
};

// torchvision F.resize with an int size: the smaller edge becomes R, the other keeps the aspect ratio (truncated);
// F.center_crop: offsets int(round((size - R) / 2.0)) with Python's round-half-to-even.
static int make_resize_plan(int n_frames, int H, int W, int R, ResizePlan* p) {
  DFD_CHECK_ARG(H > 0 && W > 0 && R > 0, "resize: bad size %dx%d -> %d", H, W, R);
  p->H = H; p->W = W; p->R = R;
  if (W <= H) {
    p->new_w = R;
    p->new_h = static_cast<int>(static_cast<int64_t>(R) * H / W);
  } else {
    p->new_h = R;
    p->new_w = static_cast<int>(static_cast<int64_t>(R) * W / H);
  }
  auto half_even = [](int d) {  // round(d / 2.0), ties to even
    const int q = d / 2;
    return (d % 2 == 0) ? q : ((q % 2 == 0) ? q : q + 1);
  };
  p->top = half_even(p->new_h - R);
  p->left = half_even(p->new_w - R);
  p->scale_x = static_cast<float>(W) / static_cast<float>(p->new_w);
  p->scale_y = static_cast<float>(H) / static_cast<float>(p->new_h);
  auto taps = [](float scale) {
    const float support = scale >= 1.f ? 2.f * scale : 2.f;
    return static_cast<int>(ceilf(support)) * 2 + 1;
  };
  p->taps_x = taps(p->scale_x);
  p->taps_y = taps(p->scale_y);
  DFD_CHECK_ARG(p->taps_x <= 129 && p->taps_y <= 129, "resize: down-scaling by more than 32x is not supported");
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += up(b); return r; };
  p->off_xmin = take(sizeof(int) * R);
  p->off_xsize = take(sizeof(int) * R);
  p->off_xw = take(sizeof(float) * R * p->taps_x);
  p->off_ymin = take(sizeof(int) * R);
  p->off_ysize = take(sizeof(int) * R);
  p->off_yw = take(sizeof(float) * R * p->taps_y);
  p->off_tmp = take(sizeof(float) * static_cast<size_t>(n_frames) * 3 * H * R);
  p->total = o;
  return 0;
}

__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return ((a * x - 5.f * a) * x + 8.f * a) * x - 4.f * a;
  return 0.f;
}

// One thread per output index i (already offset by the crop): first source tap, tap count, normalised weights.
__global__ void resize_tables_kernel(int* __restrict__ tmin, int* __restrict__ tsize, float* __restrict__ tw, int n_out,
                                     int crop_offset, int in_size, float scale, int max_taps) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int i = o + crop_offset;
  const float support = scale >= 1.f ? 2.f * scale : 2.f;
  const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
  const float center = scale * (static_cast<float>(i) + 0.5f);
  int xmin = static_cast<int>(center - support + 0.5f);
  if (xmin < 0) xmin = 0;
  int xend = static_cast<int>(center + support + 0.5f);
  if (xend > in_size) xend = in_size;
  int xsize = xend - xmin;
  if (xsize > max_taps) xsize = max_taps;
  float* w = tw + static_cast<size_t>(o) * max_taps;
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float v = cubic_aa((static_cast<float>(j + xmin) - center + 0.5f) * invscale);
    w[j] = v;
    total += v;
  }
  if (total != 0.f)
    for (int j = 0; j < xsize; ++j) w[j] = w[j] / total;
  for (int j = xsize < 0 ? 0 : xsize; j < max_taps; ++j) w[j] = 0.f;
  tmin[o] = xmin;
  tsize[o] = xsize < 0 ? 0 : xsize;
}

// tmp[plane, y, x] = sum_j src[plane, y, xmin[x] + j] * w[x][j]     plane = frame * 3 + channel, x over the R crop columns
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, float* __restrict__ tmp, const int* __restrict__ xmin,
                const int* __restrict__ xsize, const float* __restrict__ xw, int taps, int H, int W, int R,
                int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % R);
  const int64_t row = idx / R;  // plane * H + y
  const uint8_t* s = src + row * W + xmin[x];
  const float* w = xw + static_cast<size_t>(x) * taps;
  const int n = xsize[x];
  float acc = n > 0 ? static_cast<float>(s[0]) * w[0] : 0.f;
  for (int j = 1; j < n; ++j) acc += static_cast<float>(s[j]) * w[j];
  tmp[idx] = acc;
}

// out[plane, y, x] = uint8(rint(clamp(sum_j tmp[plane, ymin[y] + j, x] * w[y][j], 0, 255)))
__global__ void __launch_bounds__(256)
resize_v_kernel(const float* __restrict__ tmp, uint8_t* __restrict__ out, const int* __restrict__ ymin,
                const int* __restrict__ ysize, const float* __restrict__ yw, int taps, int H, int R, int64_t total) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % R);
  const int y = static_cast<int>((idx / R) % R);
  const int64_t plane = idx / (static_cast<int64_t>(R) * R);
  const float* s = tmp + (plane * H + ymin[y]) * R + x;
  const float* w = yw + static_cast<size_t>(y) * taps;
  const int n = ysize[y];
  float acc = n > 0 ? s[0] * w[0] : 0.f;
  for (int j = 1; j < n; ++j) acc += s[static_cast<int64_t>(j) * R] * w[j];
  acc = fminf(fmaxf(acc, 0.f), 255.f);
  out[idx] = static_cast<uint8_t>(rintf(acc));
}

int resize_crop_u8(const dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, int R, uint8_t* out,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DFD_CHECK_ARG(n_frames >= 0, "resize_crop_u8: negative frame count");
  if (n_frames == 0) return 0;
  DFD_CHECK_ARG(frames && out, "resize_crop_u8: null pointer");
  ResizePlan p;
  DFD_TRY(make_resize_plan(n_frames, H, W, R, &p));
  if (!workspace || workspace_bytes < p.total)
    return fail(DFD_ERR_WORKSPACE, "resize_crop_u8: workspace %zu < %zu bytes", workspace_bytes, p.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* xmin = reinterpret_cast<int*>(ws + p.off_xmin);
  int* xsize = reinterpret_cast<int*>(ws + p.off_xsize);
  float* xw = reinterpret_cast<float*>(ws + p.off_xw);
  int* ymin = reinterpret_cast<int*>(ws + p.off_ymin);
  int* ysize = reinterpret_cast<int*>(ws + p.off_ysize);
  float* yw = reinterpret_cast<float*>(ws + p.off_yw);
  float* tmp = reinterpret_cast<float*>(ws + p.off_tmp);
  const int tb = (R + 127) / 128;
  resize_tables_kernel<<<tb, 128, 0, stream>>>(xmin, xsize, xw, R, p.left, W, p.scale_x, p.taps_x);
  resize_tables_kernel<<<tb, 128, 0, stream>>>(ymin, ysize, yw, R, p.top, H, p.scale_y, p.taps_y);
  DFD_CUDA_OK(cudaGetLastError());
  const int64_t planes = static_cast<int64_t>(n_frames) * 3;
  const int64_t total_h = planes * H * R;
  resize_h_kernel<<<static_cast<unsigned>((total_h + 255) / 256), 256, 0, stream>>>(frames, tmp, xmin, xsize, xw,
                                                                                    p.taps_x, H, W, R, total_h);
  DFD_CUDA_OK(cudaGetLastError());
  const int64_t total_v = planes * R * R;
  resize_v_kernel<<<static_cast<unsigned>((total_v + 255) / 256), 256, 0, stream>>>(tmp, out, ymin, ysize, yw, p.taps_y,
                                                                                    H, R, total_v);
  DFD_CUDA_OK(cudaGetLastError());
  (void)ctx;
  return 0;
}

}  // namespace dfd

extern "C" {

size_t dfd_resize_crop_u8_workspace_bytes(int n_frames, int H, int W, int R) {
  dfd::ResizePlan p;
  if (n_frames <= 0 || dfd::make_resize_plan(n_frames, H, W, R, &p) != 0) return 0;
  return p.total;
}

int dfd_resize_crop_u8(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, int R, uint8_t* out,
                       void* workspace, size_t workspace_bytes, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_resize_crop_u8: ctx is NULL");
  return dfd::resize_crop_u8(ctx, frames, n_frames, H, W, R, out, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Encoder self-attention (src/clip/model.py:188-195): per frame and head, softmax_k((q/8).k) v over the L = P+1
// tokens of ONE frame (no mask). Dispatch by tokens per frame to the three tcgen05 / TMEM kernels:
//   L <= 128          mha_fwd_tc_kernel   (attention_sm100.cu)     one 128-row query tile per (frame, head)
//   128 < L <= 208    mha_fwd_tc2_kernel  (attention_sm100_v2.cu)  ViT-B/16, B/32: two pipelined tiles, P in TMEM
//   208 < L <= 257    mha_fwd_tc3_kernel  (attention_sm100_v3.cu)  ViT-L/14: 256 keys on the tensor core, token 256 around it
// Longer sequences are rejected (a frame's keys stay on chip in all three kernels). The mma.sync kernel that served
// 257 < L <= 272 in round 1 is gone: no CLIP ViT has such a sequence length and nothing pre-Blackwell ships.
#include "common.cuh"
#include "host_common.h"

namespace dfd {

int mha_fwd_tc(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);
int mha_fwd_tc2(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);
int mha_fwd_tc3(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream);

int mha_fwd(const dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, cudaStream_t stream) {
  DFD_CHECK_ARG(n_frames >= 0 && L > 0 && H > 0, "mha_fwd: bad shape");
  DFD_CHECK_ARG(L <= 257, "mha_fwd: sequence length %d > 257 tokens per frame is not supported", L);
  if (n_frames == 0) return 0;
  DFD_CHECK_ARG(qkv && mix, "mha_fwd: null pointer");
  if (L <= 128) return mha_fwd_tc(ctx, qkv, mix, n_frames, L, H, stream);
  if (L <= 208) return mha_fwd_tc2(ctx, qkv, mix, n_frames, L, H, stream);
  return mha_fwd_tc3(ctx, qkv, mix, n_frames, L, H, stream);
}

}  // namespace dfd

extern "C" int dfd_mha_fwd(dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, void* stream) {
  dfd::clear_error();
  if (!ctx) return dfd::fail(DFD_ERR_INVALID, "dfd_mha_fwd: ctx is NULL");
  return dfd::mha_fwd(ctx, qkv, mix, n_frames, L, H, static_cast<cudaStream_t>(stream));
}

/*
 * dfdclip_b200 — C ABI of the B200-native DFD-CLIP hot path (libdfdclip_b200.so).
 *
 * The reference (ODD2/DFD-CLIP) is pure Python and has no FFI for this path; the boundary the library
 * sits under is the pair of Python surfaces
 *     src/clip/model.py:276-294   VisionTransformer.forward  (frame encoder, per-layer q/k/v taps)
 *     src/models.py:323-361       Decoder.forward            (temporal decoder / classification head)
 *     src/models.py:498-566       Detector.predict           (encoder -> taps -> decoder -> 5*l/|l|)
 * Every entry point below names the reference lines it replaces. All pointers are DEVICE pointers unless
 * stated otherwise; the caller (PyTorch) owns every buffer; all work is enqueued asynchronously on the
 * `stream` argument (a cudaStream_t passed as void*); nothing synchronises the device.
 *
 * Error convention: every function returns 0 on success or a DFD_ERR_* code; dfd_last_error() returns a
 * thread-local message. There is no CPU fallback: unsupported shapes/configs are errors.
 */
#ifndef DFDCLIP_B200_H_
#define DFDCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFD_ABI_VERSION 3

enum {
  DFD_OK = 0,
  DFD_ERR_INVALID = 1,     /* bad argument / unsupported shape or config */
  DFD_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
  DFD_ERR_UNSUPPORTED = 3, /* device is not sm_100 */
  DFD_ERR_WORKSPACE = 4    /* workspace too small */
};

/* GEMM epilogues (dfd_gemm_bf16). */
enum {
  DFD_EPI_STORE_BF16 = 0,       /* out_bf16 = acc + bias                     (QKV in_proj,  model.py:186)   */
  DFD_EPI_STORE_BF16_QGELU = 1, /* out_bf16 = quickgelu(acc + bias)          (c_fc + QuickGELU, :166,:209)  */
  DFD_EPI_STORE_F32 = 2,        /* out_f32  = acc + bias                     (conv1 as GEMM, :277)          */
  DFD_EPI_ADD_F32 = 3,          /* out_f32 += acc + bias  (in-place residual, out_proj/c_proj :222-223)     */
  DFD_EPI_ADD_BF16 = 4,         /* out_bf16 += acc + bias (in-place adapter residual, src/models.py:931)    */
  DFD_EPI_STORE_BF16_GELU = 5,  /* out_bf16 = gelu_erf(acc + bias)           (nn.GELU, src/models.py:893)   */
  /* LayerNorm folded into the GEMMs around it (dfd_gemm_bf16_ln only; model.py:221-223): */
  DFD_EPI_STORE_BF16_LNFOLD = 6,       /* out_bf16 = rstd*(acc - mu*colsum) + bias      (ln_1 + in_proj)    */
  DFD_EPI_STORE_BF16_QGELU_LNFOLD = 7, /* out_bf16 = quickgelu(rstd*(acc - mu*colsum) + bias)  (ln_2 + c_fc) */
  DFD_EPI_RESID_LN_F32 = 8             /* out_f32 += acc + bias, plus bf16(out) and per-row (sum, sum sq)   */
};

typedef struct dfd_ctx dfd_ctx;

int dfd_version(void);
const char* dfd_last_error(void);

/* One context per device. Fails with DFD_ERR_UNSUPPORTED unless the device is compute capability 10.x. */
int dfd_ctx_create(int device, dfd_ctx** out);
int dfd_ctx_destroy(dfd_ctx* ctx);

/* Optional per-kernel timing for bench.py's roofline: when enabled, every kernel launch of the encoder/decoder
 * calls is bracketed by a CUDA event pair on the launching stream. dfd_timing_read synchronises on those events,
 * returns per-tag total milliseconds and launch counts (arrays of dfd_timing_num_tags() entries) and clears them. */
int dfd_timing_enable(dfd_ctx* ctx, int on);
int dfd_timing_read(dfd_ctx* ctx, int max_tags, float* total_ms, int* counts);
int dfd_timing_num_tags(void);
const char* dfd_timing_tag_name(int tag);

/* ------------------------------------------------------------------------------------------------------
 * Unit kernels (exported so each can be parity-tested on its own).
 * ---------------------------------------------------------------------------------------------------- */

/* C[M,N] (epilogue) A[M,K] * W[N,K]^T. A, W: bf16, K contiguous, row pitches lda/ldw elements (multiples of 8).
 * bias: fp32 [N] or NULL. out: bf16 or fp32 per `epilogue`, row pitch ldo elements. N % 256 == 0, K % 8 == 0.
 * tcgen05 / TMEM / TMA kernel. Replaces F.linear / nn.Linear / nn.Conv2d of src/clip/model.py:186,197,209,211,277. */
int dfd_gemm_bf16(dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                  int64_t ldo, int M, int N, int K, int epilogue, void* stream);

/* LayerNorm folded into the GEMMs on either side of it. With x the fp32 residual rows, LN(x) W^T =
 * rstd * (x (gamma.W)^T - mu * colsum) + (b + W beta), colsum[n] = sum_k (gamma.W)[n,k]: the PRODUCER of x
 * (DFD_EPI_RESID_LN_F32: x += acc + bias, the out_proj / c_proj residual update of model.py:222-223) also writes
 * bf16(x) and, per row and per 128-column block, the partial (sum, sum of squares) of the updated values; the CONSUMER
 * (DFD_EPI_STORE_BF16[_QGELU]_LNFOLD: in_proj / c_fc of model.py:186, 209 on A = bf16(x), W = bf16(gamma.W),
 * bias = b + W beta) turns the partials into mu / rstd and finishes the normalisation in its epilogue. The
 * LayerNorm pass over x disappears. N % 256 == 0, slots <= 8 (rows of at most 1024 elements); for the consumer K
 * must be the row length the statistics were taken over. */
typedef struct {
  const float* stats_in; /* LNFOLD: fp32 [M, slots, 2] */
  const float* colsum;   /* LNFOLD: fp32 [N] */
  int slots;             /* LNFOLD: partial blocks per row (= 2 * row_length / 256 as written by RESID_LN) */
  float* stats_out;      /* RESID_LN: fp32 [M, 2*N/256, 2] */
  void* bf16_out;        /* RESID_LN: bf16 [M, N], row pitch ld_bf16 elements */
  int64_t ld_bf16;
} dfd_gemm_ln_args;
int dfd_gemm_bf16_ln(dfd_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                     void* out, int64_t ldo, int M, int N, int K, int epilogue, const dfd_gemm_ln_args* args,
                     void* stream);

/* Row LayerNorm, fp32 math, eps 1e-5, biased variance (src/clip/model.py:157-163).
 * x fp32 [rows, D]; pos (optional) fp32 [pos_period, D] is added to row r as pos[r % pos_period] before the
 * statistics (cls/positional embedding add of model.py:280-291 fused into ln_pre). Exactly one of
 * out_bf16 / out_f32 is non-NULL; out_f32 may alias x. D must be a multiple of 128, <= 2048. */
int dfd_layernorm(dfd_ctx* ctx, const float* x, const float* gamma, const float* beta, const float* pos,
                  int pos_period, void* out_bf16, float* out_f32, int64_t rows, int D, void* stream);

/* frames fp32 [n_frames,3,R,R] -> patch matrix bf16 [n_frames*(P+1), Kp]; row f*(P+1) is zero (cls slot),
 * row f*(P+1)+1+p holds patch p flattened (c, i, j) like conv1.weight[D,3,p,p] (model.py:277-279).
 * Columns [3*p*p, Kp) are zero. */
int dfd_patchify(dfd_ctx* ctx, const float* frames, void* out_bf16, int n_frames, int R, int patch, int Kp,
                 void* stream);

/* dfd_patchify for uint8 frames [n_frames,3,R,R]: every pixel becomes (x / 255 - mean[c]) / std[c] in fp32 —
 * T.ConvertImageDtype(torch.float32) + T.Normalize of the reference's CPU transform (src/models.py:762-768), same
 * operation order — before the bf16 cast. mean_std: HOST array {mean[3], std[3]}. frames 8-byte aligned. */
int dfd_patchify_u8(dfd_ctx* ctx, const uint8_t* frames, const float* mean_std, void* out_bf16, int n_frames, int R,
                    int patch, int Kp, void* stream);

/* Encoder self-attention over one packed QKV buffer (model.py:188-195): qkv bf16 [n_frames*L, 3*D]
 * (row = [q | k | v], each H x 64), no mask, softmax over keys of (q/8).k; mix bf16 [n_frames*L, D]. dh = 64. */
int dfd_mha_fwd(dfd_ctx* ctx, const void* qkv, void* mix, int n_frames, int L, int H, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Encoder: VisionTransformer.forward (src/clip/model.py:276-294, Transformer.forward :236-251).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int image_size; /* R   */
  int patch_size; /* p   */
  int width;      /* D   */
  int heads;      /* H, D == 64*H */
  int layers;
} dfd_vit_dims;

/* fp32 device pointers in the reference state_dict layout. Per-layer members are HOST
 * arrays of `layers` device pointers. */
typedef struct {
  const float* conv1_weight;         /* [D,3,p,p] */
  const float* class_embedding;      /* [D]       */
  const float* positional_embedding; /* [P+1,D]   */
  const float* ln_pre_weight;
  const float* ln_pre_bias;
  const float* const* ln_1_weight;
  const float* const* ln_1_bias;
  const float* const* in_proj_weight; /* [3D,D] */
  const float* const* in_proj_bias;   /* [3D]   */
  const float* const* out_proj_weight; /* [D,D] */
  const float* const* out_proj_bias;
  const float* const* ln_2_weight;
  const float* const* ln_2_bias;
  const float* const* c_fc_weight;   /* [4D,D] */
  const float* const* c_fc_bias;
  const float* const* c_proj_weight; /* [D,4D] */
  const float* const* c_proj_bias;
} dfd_vit_weights;

size_t dfd_encoder_packed_bytes(const dfd_vit_dims* dims);
size_t dfd_encoder_workspace_bytes(const dfd_vit_dims* dims, int n_frames);

/* Convert the fp32 parameters to the library's packed layout (bf16 GEMM operands, fp32 biases/LN/pos). */
int dfd_encoder_pack_weights(dfd_ctx* ctx, const dfd_vit_dims* dims, const dfd_vit_weights* w, void* packed,
                             void* stream);

/* frames fp32 [n_frames,3,R,R]. Runs layers [0, num_run_layers). If last_qkv_only != 0 the last of those
 * layers stops after its K/V projection (its attention/MLP output is not needed: SURVEY note D1; the Q block of that
 * layer's buffer is NOT written).
 * qkv_out: HOST array [layers] of bf16 [n_frames*L, 3D] device buffers (NULL entry: QKV of that layer is kept
 * only in scratch). x_out: NULL or HOST array [layers] of fp32 [n_frames*L, D] device buffers receiving the
 * residual stream after each layer (`with_out`, model.py:242). */
int dfd_encoder_forward(dfd_ctx* ctx, const dfd_vit_dims* dims, const void* packed, const float* frames,
                        int n_frames, int num_run_layers, int last_qkv_only, void* const* qkv_out,
                        float* const* x_out, void* workspace, size_t workspace_bytes, void* stream);

/* dfd_encoder_forward on uint8 frames: the float conversion and normalisation of the data loader's transform
 * (src/models.py:762-768) are fused into the patch extraction, so a clip crosses PCIe and HBM as 1 byte per
 * pixel instead of 4. mean_std: HOST array {mean[3], std[3]}. */
int dfd_encoder_forward_u8(dfd_ctx* ctx, const dfd_vit_dims* dims, const void* packed, const uint8_t* frames,
                           const float* mean_std, int n_frames, int num_run_layers, int last_qkv_only,
                           void* const* qkv_out, float* const* x_out, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Decoder: Decoder.forward / Transformer.forward / MultiheadAttention.forward (src/models.py:323-361,
 * 259-269, 136-146) in fp32, streaming the tapped bf16 K/V once per block.
 * ---------------------------------------------------------------------------------------------------- */
enum { DFD_ATTN_FRAME = 1, DFD_ATTN_TEMPORAL = 2 };

typedef struct {
  const float* class_embedding;      /* [D] */
  const float* positional_embedding; /* [T,1,H,64] or NULL (op_mode.temporal_position == 0) */
  const float* ln_pre_weight;
  const float* ln_pre_bias;
  const float* ln_post_weight;
  const float* ln_post_bias;
  /* HOST arrays of n_blocks device pointers */
  const float* const* ln_1_weight;
  const float* const* ln_1_bias;
  const float* const* in_proj_weight; /* [2D,D] */
  const float* const* in_proj_bias;   /* [2D]   */
  const float* const* out_proj_weight; /* [D,D] */
  const float* const* out_proj_bias;
  const float* const* ln_2_weight;
  const float* const* ln_2_bias;
  const float* const* c_fc_weight;
  const float* const* c_fc_bias;
  const float* const* c_proj_weight;
  const float* const* c_proj_bias;
  /* op_mode.aug_query (src/models.py:250-255, 265-267): HOST array of n_blocks - 1 device pointers to fp32 [D]
   * vectors added to the query after every block but the last, or NULL. */
  const float* const* augment_query;
  /* op_mode.attn_mode (src/models.py:107-115): 0 = softmax over all T*P keys (default), else a combination of
   * DFD_ATTN_FRAME (softmax over the patches of each frame) and DFD_ATTN_TEMPORAL (over the frames of each patch). */
  int attn_mode;
} dfd_decoder_weights;

/* K/V of tapped layer i: element (b,t,p,h,c) lives at k[i] + b*stride_b + t*stride_t + p*stride_p + h*64 + c
 * (bf16 elements). Views of the encoder's QKV buffers satisfy this without a copy. */
typedef struct {
  const void* const* k; /* HOST array [n_blocks] */
  const void* const* v;
  int64_t stride_b, stride_t, stride_p;
} dfd_kv_taps;

/* P and attn_mode size the score buffers of the non-default attention modes (0 for the default mode). */
size_t dfd_decoder_workspace_bytes(int B, int T, int P, int D, int n_blocks, int attn_mode);

/* mask: uint8 [B,T] (1 = frame present; src/models.py:324). block_out: fp32 [B, n_blocks, D] (x after every
 * block, the torch.cat of models.py:269). video_feature: fp32 [B, D] = ln_post(block_out[:, -1]) (:340-343). */
int dfd_decoder_forward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                        const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                        float* video_feature, void* workspace, size_t workspace_bytes, void* stream);

/* Detector.predict (src/models.py:498-566: encoder, taps, decoder) as ONE call = dfd_encoder_forward[_u8] +
 * dfd_decoder_forward with the same arguments, except that with overlap != 0 decoder block i is launched on a stream
 * owned by the context as soon as the K/V projection of encoder layer tap_layers[i] has been enqueued, so the
 * one-token decoder (launch/HBM-bound) runs beside the tensor-bound encoder layers that follow its tap. `stream` is
 * joined with that stream before the call returns (event fork/join: CUDA-graph capturable); results are bit-identical
 * to the two separate calls. frames: fp32 [n_frames,3,R,R], or uint8 with frames_are_u8 != 0 (then mean_std as in
 * dfd_encoder_forward_u8). tap_layers: HOST array [n_blocks], taps->k[i] / v[i] must point into qkv_out[tap_layers[i]].
 * n_frames must equal B*T. While per-kernel timing is enabled (dfd_timing_enable) the call runs without overlap. */
int dfd_predict_forward(dfd_ctx* ctx, const dfd_vit_dims* dims, const void* packed, const void* frames,
                        int frames_are_u8, const float* mean_std, int n_frames, int num_run_layers, int last_qkv_only,
                        void* const* qkv_out, void* enc_workspace, size_t enc_workspace_bytes, int D, int H,
                        int n_blocks, const dfd_decoder_weights* w, const dfd_kv_taps* taps, const int* tap_layers,
                        const uint8_t* mask, int B, int T, int P, float* block_out, float* video_feature,
                        void* dec_workspace, size_t dec_workspace_bytes, int overlap, void* stream);

/* logits[b,:] = scale * l / (||l||_2 + 1e-10), l = feature[b,:] @ proj[D,O] (models.py:359, 551-553).
 * scale <= 0 skips the normalisation. */
int dfd_project_logits(dfd_ctx* ctx, const float* feature, const float* proj, int B, int D, int O, float scale,
                       float* logits, void* stream);

/* The decoder's one-token-per-clip nn.Linear layers (src/models.py:69-73, 136, 145: in_proj, out_proj, c_fc, c_proj)
 * in exact fp32 FMA arithmetic, exported for unit tests: out[b,n] = act(sum_k x[b,k] W[n,k] + bias[n]) + residual[b,n],
 * act = QuickGELU (model.py:166-168) when quick_gelu != 0; bias / residual may be NULL, residual may alias out.
 * x fp32 [B,K], W fp32 [N,K] (nn.Linear layout), K % 4 == 0. Deterministic (fixed-order split-K reduction). */
size_t dfd_linear_f32_workspace_bytes(int B, int N);
int dfd_linear_f32(dfd_ctx* ctx, const float* x, const float* W, const float* bias, const float* residual, float* out,
                   int B, int N, int K, int quick_gelu, void* workspace, size_t workspace_bytes, void* stream);

/* op_mode.ema_frame (src/models.py:572-578): out[b] = EMA over the T frames of clip b, _x = _x*ratio + x[:, i]*(1-ratio)
 * starting from zero, evaluated in that order in fp32. x fp32 [B, T, frame_elems] -> out fp32 [B, frame_elems];
 * frame_elems % 4 == 0. */
int dfd_ema_frames(dfd_ctx* ctx, const float* x, float* out, int B, int T, int64_t frame_elems, float ratio,
                   void* stream);

/* One decoder attention call on its own (models.py:136-146 without in/out projections), for unit tests:
 * qs fp32 [B, H, 128] = per head [smax query(64) | coda query(64)]; mix fp32 [B, H*64].
 * workspace: at least 2*B*T*H*130*4 bytes (partial records of the streaming kernel). */
int dfd_decoder_attention(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                          int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                          int T, int P, int H, float* mix, void* workspace, size_t workspace_bytes, void* stream);

/* dfd_decoder_attention for op_mode.attn_mode != 0 (DFD_ATTN_FRAME | DFD_ATTN_TEMPORAL): scores, per-group
 * normalisation and the weighted sum of V as three passes. An all-masked softmax group yields NaN like the reference. */
size_t dfd_decoder_attention_modes_workspace_bytes(int B, int T, int P, int H);
int dfd_decoder_attention_modes(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                                int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                                int T, int P, int H, int attn_mode, float* mix, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * CompInvAdapter on the tapped K/V (src/models.py:783-940; called from Detector.predict :546-547): a bottleneck
 * per tapped layer and per {k, v}, applied IN PLACE to a bf16 tap (residual structs) before the decoder reads it.
 * ---------------------------------------------------------------------------------------------------- */
enum {
  DFD_ADAPTER_GELU_LN = 0, /* "768-x-768", "legacy-768-x-768": Linear, GELU, LayerNorm(x), Linear   (:797-818) */
  DFD_ADAPTER_LN_GELU = 1, /* "768-x-768-ln", "768-x-768-z0":  Linear, LayerNorm(x), GELU, Linear   (:835-861) */
  DFD_ADAPTER_NLN = 2,     /* "768-x-768-nln": Linear, LayerNorm((patches, x)), GELU, Linear        (:819-834) */
  DFD_ADAPTER_XXX = 3,     /* "768-xxx-768":   Linear, GELU, Linear, GELU, Linear                   (:881-899) */
  DFD_ADAPTER_LINEAR = 4,  /* "linear":        Linear(D, D), no residual                            (:900-917) */
  DFD_ADAPTER_BN = 5       /* "768-bn":        Linear(D, D), BatchNorm2d(num_frames) in eval mode   (:877-887):
                            * kv += scale[f] * (kv . W^T) + shift[f], f = frame of the row; ln_weight / ln_bias carry
                            * fp32 [rows / group_rows] per-frame scale = gamma / sqrt(running_var + eps) and
                            * shift = beta - running_mean * scale (the frame's position in its clip picks the channel) */
};

typedef struct {
  const void* w_down;     /* bf16 [inner, D]   Sequential[0].weight   ([D, D] for DFD_ADAPTER_LINEAR) */
  const void* w_mid;      /* bf16 [inner, inner], DFD_ADAPTER_XXX only, else NULL */
  const void* w_up;       /* bf16 [D, inner]   last Linear.weight */
  const float* ln_weight; /* fp32 [inner]; [group_rows - group_skip, inner] for DFD_ADAPTER_NLN */
  const float* ln_bias;
} dfd_adapter_weights;

size_t dfd_adapter_workspace_bytes(int type, int D, int inner, int64_t rows);

/* kv: bf16 [rows, D] with row pitch `ld` elements (a K or V column slice of a packed QKV buffer: ld = 3D), updated
 * in place: kv += up(act(down(kv))) (kv = down(kv) for DFD_ADAPTER_LINEAR). inner in {256, 512, 768, 1024}.
 * group_rows / group_skip (DFD_ADAPTER_NLN only): rows per frame and the leading rows of each frame that take no
 * part in the joint normalisation (L and 1 for the packed encoder layout with its CLS row; P and 0 otherwise). */
int dfd_adapter_apply(dfd_ctx* ctx, int type, int D, int inner, const dfd_adapter_weights* w, void* kv, int64_t ld,
                      int64_t rows, int group_rows, int group_skip, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Training step (config C5): the encoder is frozen, so only the decoder needs gradients. The decoder attention is
 * the one op of the head that streams the K/V taps; its forward-with-statistics and its backward are native, the
 * 1-token-per-clip linear / LayerNorm layers around it run under torch autograd (src/models.py:568-596 + autograd).
 * ---------------------------------------------------------------------------------------------------- */
size_t dfd_decoder_attention_workspace_bytes(int B, int T, int H);

/* dfd_decoder_attention that also saves, per (clip, head), stats fp32 [B, H, 66] = (softmax max M, softmax sum L,
 * o0[64] = softmax-weighted mean of V + pe) for the backward pass. */
int dfd_decoder_attention_train(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                                int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask, int B,
                                int T, int P, int H, float* mix, float* stats, void* workspace, size_t workspace_bytes,
                                void* stream);

/* Gradients of the decoder attention w.r.t. the queries (dqs fp32 [B, H, 128]) and the temporal position embedding
 * (dpos_emb fp32 [T, H, 64], NULL iff pos_emb is NULL) given dmix fp32 [B, H*64]. dk / dv: both NULL (frozen taps),
 * or contiguous fp32 [B, T, P, H, 64] buffers receiving the gradients w.r.t. K and V (trainable adapter on the
 * taps, src/models.py:546-547); keys of masked frames get zeros. */
int dfd_decoder_attention_backward(dfd_ctx* ctx, const float* qs, const void* k, const void* v, int64_t stride_b,
                                   int64_t stride_t, int64_t stride_p, const float* pos_emb, const uint8_t* mask,
                                   const float* stats, const float* dmix, int B, int T, int P, int H, float* dqs,
                                   float* dpos_emb, float* dk, float* dv, void* workspace, size_t workspace_bytes,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Input side: the data loader's T.Resize(n_px, BICUBIC) + T.CenterCrop(n_px) (src/models.py:756-761; applied to the
 * stacked uint8 frames at src/datasets.py:672-676) on the device. frames: uint8 [n_frames, 3, H, W] of any size;
 * out: uint8 [n_frames, 3, R, R]. torchvision's tensor path: the smaller edge is resized to R (the other keeps the
 * aspect ratio), antialiased separable bicubic interpolation in fp32 (Keys a = -0.5, support 2 * max(scale, 1),
 * normalised weights, width pass first), clamp to [0, 255], round half to even, then the centre crop — reproduced
 * within 1 LSB (fp32 summation order). The remaining transform steps (ConvertImageDtype + Normalize) are fused into
 * dfd_patchify_u8 / dfd_encoder_forward_u8. */
size_t dfd_resize_crop_u8_workspace_bytes(int n_frames, int H, int W, int R);
int dfd_resize_crop_u8(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, int R, uint8_t* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Training step of the decoder (BASELINE config C5). Replaces, inside the reference trainer's
 * forward(train=True) -> backward (src/trainer.py:147-165), the autograd graph of Decoder.forward's chain
 * (src/models.py:336-338, 259-269, 173-176, 136-146: ln_pre(class_embedding), then per block ln_1 -> in_proj ->
 * attention -> out_proj -> residual -> ln_2 -> c_fc -> QuickGELU -> c_proj -> residual, aug_query between blocks).
 * ln_post, the task projection(s), the logit normalisation and the loss stay with the caller.
 *
 * dfd_decoder_train_forward: block_out fp32 [B, n_blocks, D] as dfd_decoder_forward, activations kept in `saved`
 * (dfd_decoder_train_bytes; the same buffer must be handed to the backward untouched).
 * dfd_decoder_train_backward: given d_block_out fp32 [B, n_blocks, D], writes the gradient of every parameter of the
 * chain through `grads` — a dfd_decoder_weights whose pointers are OUTPUT buffers of the parameters' shapes
 * (class_embedding, positional_embedding [summed over the blocks], ln_pre_*, per block ln_1_*, in_proj_*, out_proj_*,
 * ln_2_*, c_fc_*, c_proj_*, augment_query; ln_post_* and attn_mode are ignored). dk / dv: NULL, or HOST arrays of
 * n_blocks device pointers to contiguous fp32 [B, T, P, H, 64] buffers for the gradients w.r.t. the tapped K / V (a
 * trainable adapter on the taps). Blocks block_hi .. block_lo (descending) are processed by one call: (n_blocks - 1, 0)
 * does everything; a caller that overlaps a gradient all-reduce with the backward walks down in several calls (the
 * gradient in flight lives in `saved`; class_embedding / ln_pre gradients come with the call that includes block 0,
 * positional_embedding's is complete after it). op_mode.attn_mode != 0 is rejected. fp32 throughout, deterministic. */
size_t dfd_decoder_train_bytes(int B, int T, int D, int n_blocks);
int dfd_decoder_train_forward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                              const dfd_kv_taps* taps, const uint8_t* mask, int B, int T, int P, float* block_out,
                              void* saved, size_t saved_bytes, void* stream);
int dfd_decoder_train_backward(dfd_ctx* ctx, int D, int H, int n_blocks, const dfd_decoder_weights* w,
                               const dfd_decoder_weights* grads, const dfd_kv_taps* taps, const uint8_t* mask, int B,
                               int T, int P, const float* d_block_out, float* const* dk, float* const* dv, void* saved,
                               size_t saved_bytes, int block_hi, int block_lo, void* stream);

/* The forward half of the trainer's step (src/trainer.py:147-156: frozen encoder under no_grad, then the decoder) as
 * ONE call = dfd_encoder_forward[_u8] + dfd_decoder_train_forward with the same arguments; with overlap != 0 the
 * activation-saving decoder block i runs on the context's stream right behind the K/V projection of encoder layer
 * tap_layers[i], beside the encoder layers that follow (same fork/join as dfd_predict_forward: graph capturable,
 * bit-identical to the two separate calls). The backward is dfd_decoder_train_backward on the same `saved`. */
int dfd_train_forward(dfd_ctx* ctx, const dfd_vit_dims* dims, const void* packed, const void* frames, int frames_are_u8,
                      const float* mean_std, int n_frames, int num_run_layers, int last_qkv_only, void* const* qkv_out,
                      void* enc_workspace, size_t enc_workspace_bytes, int D, int H, int n_blocks,
                      const dfd_decoder_weights* w, const dfd_kv_taps* taps, const int* tap_layers, const uint8_t* mask,
                      int B, int T, int P, float* block_out, void* saved, size_t saved_bytes, int overlap,
                      void* stream);

/* Backward of one of the decoder's nn.Linear layers (dfd_linear_f32), exported for unit tests:
 * dx[b,k] = (sum_n dy[b,n] W[n,k]) * quickgelu'(gelu_pre[b,k]) + dx_add[b,k];  dW[n,k] = sum_b dy[b,n] x[b,k];
 * db[n] = sum_b dy[b,n]. gelu_pre / dx_add / dx / dW / db may be NULL (db is only written together with dW). */
size_t dfd_linear_f32_backward_workspace_bytes(int B, int N, int K);
int dfd_linear_f32_backward(dfd_ctx* ctx, const float* x, const float* W, const float* dy, const float* gelu_pre,
                            const float* dx_add, float* dx, float* dW, float* db, int B, int N, int K, void* workspace,
                            size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFDCLIP_B200_H_ */

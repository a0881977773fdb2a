"""ctypes binding of libdfdclip_b200.so — the C ABI declared in include/dfdclip_b200.h.

There is no Python/CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfdclip_b200.so")

c_void_p, c_int, c_int64, c_size_t, c_float = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float)

EPI_STORE_BF16, EPI_STORE_BF16_QGELU, EPI_STORE_F32, EPI_ADD_F32, EPI_ADD_BF16, EPI_STORE_BF16_GELU = 0, 1, 2, 3, 4, 5
ADAPTER_GELU_LN, ADAPTER_LN_GELU, ADAPTER_NLN, ADAPTER_XXX, ADAPTER_LINEAR, ADAPTER_BN = 0, 1, 2, 3, 4, 5

# Every symbol include/dfdclip_b200.h declares (tests check the library exports each of them).
EXPORTS = (
    "dfd_version", "dfd_last_error", "dfd_ctx_create", "dfd_ctx_destroy",
    "dfd_gemm_bf16", "dfd_layernorm", "dfd_patchify", "dfd_mha_fwd",
    "dfd_encoder_packed_bytes", "dfd_encoder_workspace_bytes", "dfd_encoder_pack_weights", "dfd_encoder_forward",
    "dfd_decoder_workspace_bytes", "dfd_decoder_forward", "dfd_project_logits", "dfd_decoder_attention",
    "dfd_timing_enable", "dfd_timing_read", "dfd_timing_num_tags", "dfd_timing_tag_name",
    "dfd_decoder_attention_workspace_bytes", "dfd_decoder_attention_train", "dfd_decoder_attention_backward",
    "dfd_adapter_workspace_bytes", "dfd_adapter_apply", "dfd_ema_frames",
    "dfd_decoder_attention_modes_workspace_bytes", "dfd_decoder_attention_modes",
    "dfd_patchify_u8", "dfd_encoder_forward_u8", "dfd_gemm_bf16_ln", "dfd_predict_forward",
    "dfd_linear_f32_workspace_bytes", "dfd_linear_f32",
    "dfd_resize_crop_u8_workspace_bytes", "dfd_resize_crop_u8",
    "dfd_decoder_train_bytes", "dfd_decoder_train_forward", "dfd_decoder_train_backward", "dfd_train_forward",
    "dfd_linear_f32_backward_workspace_bytes", "dfd_linear_f32_backward",
)


class NativeError(RuntimeError):
    pass


class VitDims(ctypes.Structure):
    _fields_ = [("image_size", c_int), ("patch_size", c_int), ("width", c_int), ("heads", c_int), ("layers", c_int)]


_PP = ctypes.POINTER(c_void_p)


class VitWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "conv1_weight", "class_embedding", "positional_embedding", "ln_pre_weight", "ln_pre_bias")] + [
        (n, _PP) for n in (
            "ln_1_weight", "ln_1_bias", "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
            "ln_2_weight", "ln_2_bias", "c_fc_weight", "c_fc_bias", "c_proj_weight", "c_proj_bias")]


class DecoderWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "class_embedding", "positional_embedding", "ln_pre_weight", "ln_pre_bias", "ln_post_weight",
        "ln_post_bias")] + [
        (n, _PP) for n in (
            "ln_1_weight", "ln_1_bias", "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
            "ln_2_weight", "ln_2_bias", "c_fc_weight", "c_fc_bias", "c_proj_weight", "c_proj_bias",
            "augment_query")] + [("attn_mode", c_int)]


ATTN_FRAME, ATTN_TEMPORAL = 1, 2


class KvTaps(ctypes.Structure):
    _fields_ = [("k", _PP), ("v", _PP), ("stride_b", c_int64), ("stride_t", c_int64), ("stride_p", c_int64)]


class GemmLnArgs(ctypes.Structure):
    _fields_ = [("stats_in", c_void_p), ("colsum", c_void_p), ("slots", c_int), ("stats_out", c_void_p),
                ("bf16_out", c_void_p), ("ld_bf16", c_int64)]


EPI_STORE_BF16_LNFOLD, EPI_STORE_BF16_QGELU_LNFOLD, EPI_RESID_LN_F32 = 6, 7, 8


class AdapterWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_down", "w_mid", "w_up", "ln_weight", "ln_bias")]


_lib = None
_lib_lock = threading.Lock()
_ctxs = {}


def load_library():
    """dlopen the in-tree library; raise (never fall back) when it is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "libdfdclip_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python dfd-clip_b200/build.py`. There is no CPU/PyTorch fallback for this path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.dfd_last_error.restype = ctypes.c_char_p
        lib.dfd_version.restype = c_int
        lib.dfd_ctx_create.argtypes = [c_int, ctypes.POINTER(c_void_p)]
        lib.dfd_ctx_destroy.argtypes = [c_void_p]
        lib.dfd_gemm_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                      c_int, c_int, c_int, c_int, c_void_p]
        lib.dfd_layernorm.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int64, c_int, c_void_p]
        lib.dfd_patchify.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        lib.dfd_mha_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        lib.dfd_encoder_packed_bytes.argtypes = [ctypes.POINTER(VitDims)]
        lib.dfd_encoder_packed_bytes.restype = c_size_t
        lib.dfd_encoder_workspace_bytes.argtypes = [ctypes.POINTER(VitDims), c_int]
        lib.dfd_encoder_workspace_bytes.restype = c_size_t
        lib.dfd_encoder_pack_weights.argtypes = [c_void_p, ctypes.POINTER(VitDims), ctypes.POINTER(VitWeights),
                                                 c_void_p, c_void_p]
        lib.dfd_encoder_forward.argtypes = [c_void_p, ctypes.POINTER(VitDims), c_void_p, c_void_p, c_int, c_int,
                                            c_int, _PP, _PP, c_void_p, c_size_t, c_void_p]
        lib.dfd_gemm_bf16_ln.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                         c_int, c_int, c_int, c_int, ctypes.POINTER(GemmLnArgs), c_void_p]
        lib.dfd_patchify_u8.argtypes = [c_void_p, c_void_p, ctypes.POINTER(c_float), c_void_p, c_int, c_int, c_int,
                                        c_int, c_void_p]
        lib.dfd_encoder_forward_u8.argtypes = [c_void_p, ctypes.POINTER(VitDims), c_void_p, c_void_p,
                                               ctypes.POINTER(c_float), c_int, c_int, c_int, _PP, _PP, c_void_p,
                                               c_size_t, c_void_p]
        lib.dfd_decoder_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
        lib.dfd_decoder_workspace_bytes.restype = c_size_t
        lib.dfd_decoder_forward.argtypes = [c_void_p, c_int, c_int, c_int, ctypes.POINTER(DecoderWeights),
                                            ctypes.POINTER(KvTaps), c_void_p, c_int, c_int, c_int, c_void_p,
                                            c_void_p, c_void_p, c_size_t, c_void_p]
        lib.dfd_predict_forward.argtypes = [c_void_p, ctypes.POINTER(VitDims), c_void_p, c_void_p, c_int,
                                            ctypes.POINTER(c_float), c_int, c_int, c_int, _PP, c_void_p, c_size_t,
                                            c_int, c_int, c_int, ctypes.POINTER(DecoderWeights),
                                            ctypes.POINTER(KvTaps), ctypes.POINTER(c_int), c_void_p, c_int, c_int,
                                            c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]
        lib.dfd_train_forward.argtypes = [c_void_p, ctypes.POINTER(VitDims), c_void_p, c_void_p, c_int,
                                          ctypes.POINTER(c_float), c_int, c_int, c_int, _PP, c_void_p, c_size_t,
                                          c_int, c_int, c_int, ctypes.POINTER(DecoderWeights),
                                          ctypes.POINTER(KvTaps), ctypes.POINTER(c_int), c_void_p, c_int, c_int,
                                          c_int, c_void_p, c_void_p, c_size_t, c_int, c_void_p]
        lib.dfd_project_logits.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                           c_void_p]
        lib.dfd_linear_f32_workspace_bytes.argtypes = [c_int, c_int]
        lib.dfd_linear_f32_workspace_bytes.restype = c_size_t
        lib.dfd_linear_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_void_p, c_size_t, c_void_p]
        lib.dfd_decoder_attention.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                              c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                              c_size_t, c_void_p]
        lib.dfd_decoder_attention_workspace_bytes.argtypes = [c_int, c_int, c_int]
        lib.dfd_decoder_attention_workspace_bytes.restype = c_size_t
        lib.dfd_decoder_attention_train.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                                    c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                                    c_void_p, c_void_p, c_size_t, c_void_p]
        lib.dfd_decoder_attention_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                                       c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                       c_size_t, c_void_p]
        lib.dfd_resize_crop_u8_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        lib.dfd_resize_crop_u8_workspace_bytes.restype = c_size_t
        lib.dfd_resize_crop_u8.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                           c_void_p]
        lib.dfd_decoder_train_bytes.argtypes = [c_int, c_int, c_int, c_int]
        lib.dfd_decoder_train_bytes.restype = c_size_t
        lib.dfd_decoder_train_forward.argtypes = [c_void_p, c_int, c_int, c_int, ctypes.POINTER(DecoderWeights),
                                                  ctypes.POINTER(KvTaps), c_void_p, c_int, c_int, c_int, c_void_p,
                                                  c_void_p, c_size_t, c_void_p]
        lib.dfd_decoder_train_backward.argtypes = [c_void_p, c_int, c_int, c_int, ctypes.POINTER(DecoderWeights),
                                                   ctypes.POINTER(DecoderWeights), ctypes.POINTER(KvTaps), c_void_p,
                                                   c_int, c_int, c_int, c_void_p, _PP, _PP, c_void_p, c_size_t,
                                                   c_int, c_int, c_void_p]
        lib.dfd_linear_f32_backward_workspace_bytes.argtypes = [c_int, c_int, c_int]
        lib.dfd_linear_f32_backward_workspace_bytes.restype = c_size_t
        lib.dfd_linear_f32_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
        lib.dfd_adapter_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int64]
        lib.dfd_adapter_workspace_bytes.restype = c_size_t
        lib.dfd_adapter_apply.argtypes = [c_void_p, c_int, c_int, c_int, ctypes.POINTER(AdapterWeights), c_void_p,
                                          c_int64, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p]
        lib.dfd_decoder_attention_modes_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        lib.dfd_decoder_attention_modes_workspace_bytes.restype = c_size_t
        lib.dfd_decoder_attention_modes.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                                    c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                                    c_void_p, c_size_t, c_void_p]
        lib.dfd_ema_frames.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_void_p]
        lib.dfd_timing_enable.argtypes = [c_void_p, c_int]
        lib.dfd_timing_read.argtypes = [c_void_p, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_int)]
        lib.dfd_timing_tag_name.argtypes = [c_int]
        lib.dfd_timing_tag_name.restype = ctypes.c_char_p
        for name in EXPORTS:
            fn = getattr(lib, name)
            if fn.restype is c_int and name not in ("dfd_version",):
                fn.restype = c_int
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load_library().dfd_last_error()
        raise NativeError("libdfdclip_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ctx(device):
    """Per-device context handle (created on first use)."""
    lib = load_library()
    if not torch.cuda.is_available():
        raise NativeError("dfdclip_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    index = torch.device(device).index
    if index is None:
        index = torch.cuda.current_device()
    with _lib_lock:
        handle = _ctxs.get(index)
    if handle is None:
        out = c_void_p()
        check(lib.dfd_ctx_create(index, ctypes.byref(out)))
        handle = out
        with _lib_lock:
            _ctxs[index] = handle
    return handle


def timing_enable(device, on=True):
    """Bracket every kernel of the encoder/decoder calls with CUDA events (see dfd_timing_* in the header)."""
    check(load_library().dfd_timing_enable(ctx(device), 1 if on else 0))


def timing_read(device):
    """{kernel tag: (total_ms, launches)} accumulated since the last read; synchronises on the events."""
    lib = load_library()
    n = lib.dfd_timing_num_tags()
    ms = (c_float * n)()
    cnt = (c_int * n)()
    check(lib.dfd_timing_read(ctx(device), n, ms, cnt))
    return {lib.dfd_timing_tag_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}


class Workspace:
    """A grow-only scratch buffer of a module. A CUDA graph that captured a launch has the buffer's address baked in,
    so once a buffer has been handed out under stream capture it is never freed: a later, larger request allocates a
    new buffer and the old one is kept alive for the graphs that reference it (replaying after an eager call with a
    bigger batch would otherwise read and write memory already returned to the caching allocator)."""

    def __init__(self):
        self.buf = None
        self.captured = False
        self.retired = []

    def numel(self):
        return 0 if self.buf is None else self.buf.numel()

    def get(self, nbytes, device):
        nbytes = max(int(nbytes), 16)
        buf = self.buf
        capturing = torch.cuda.is_current_stream_capturing()
        if buf is None or buf.numel() < nbytes or buf.device != device:
            if buf is not None and self.captured:
                self.retired.append(buf)
            self.buf = None  # release before growing (unless a graph may still reference it)
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf = buf
            self.captured = False
        if capturing:
            self.captured = True
        return buf


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return c_void_p(0 if t is None else t.data_ptr())


def ptr_array(tensors):
    """Host array of device pointers (keeps no reference: caller must keep the tensors alive)."""
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


# ------------------------------------------------------------------------------------------ unit kernels
def gemm_bf16(a, w, bias, out, epilogue):
    """out (epilogue) a[M,K] @ w[N,K]^T on the tcgen05 kernel. a, w: bf16 2-D with unit inner stride."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    assert out.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    check(load_library().dfd_gemm_bf16(ctx(a.device), ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out),
                                       out.stride(0), m, n, k, epilogue, stream_ptr(a.device)))
    return out


def gemm_resid_ln(a, w, bias, x):
    """x (fp32, in place) += a @ w^T + bias on the SM-pair kernel; returns (bf16 copy of the new x, per-row partial
    statistics [M, 2*N/256, 2]) for a LayerNorm-folded consumer GEMM."""
    m, k = a.shape
    n = w.shape[0]
    xb = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    stats = torch.empty((m, 2 * n // 256, 2), dtype=torch.float32, device=a.device)
    args = GemmLnArgs(None, None, 0, stats.data_ptr(), xb.data_ptr(), xb.stride(0))
    check(load_library().dfd_gemm_bf16_ln(ctx(a.device), ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(x),
                                          x.stride(0), m, n, k, EPI_RESID_LN_F32, ctypes.byref(args),
                                          stream_ptr(a.device)))
    return xb, stats


def gemm_lnfold(xb, stats, w_folded, colsum, bias_folded, out, quickgelu=False):
    """out (bf16) = [quickgelu](rstd * (xb @ w_folded^T - mu * colsum) + bias_folded), mu / rstd from `stats`."""
    m, k = xb.shape
    n = w_folded.shape[0]
    args = GemmLnArgs(stats.data_ptr(), colsum.data_ptr(), stats.shape[1], None, None, 0)
    check(load_library().dfd_gemm_bf16_ln(
        ctx(xb.device), ptr(xb), xb.stride(0), ptr(w_folded), w_folded.stride(0), ptr(bias_folded), ptr(out),
        out.stride(0), m, n, k, EPI_STORE_BF16_QGELU_LNFOLD if quickgelu else EPI_STORE_BF16_LNFOLD,
        ctypes.byref(args), stream_ptr(xb.device)))
    return out


def layernorm(x, gamma, beta, pos=None, out_dtype=torch.bfloat16, out=None):
    """Row LayerNorm of fp32 x [rows, D] (eps 1e-5); optional periodic `pos` [period, D] added first."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    rows, d = x.shape
    if out is None:
        out = torch.empty((rows, d), dtype=out_dtype, device=x.device)
    bf = out.dtype == torch.bfloat16
    if rows == 0:
        return out
    check(load_library().dfd_layernorm(ctx(x.device), ptr(x), ptr(gamma), ptr(beta), ptr(pos),
                                       0 if pos is None else pos.shape[0], ptr(out) if bf else None,
                                       None if bf else ptr(out), rows, d, stream_ptr(x.device)))
    return out


def linear_f32(x, weight, bias=None, residual=None, quick_gelu=False, out=None):
    """fp32 ``act(x @ weight.T + bias) + residual`` with the decoder's small-batch linear kernel (x [B,K], weight [N,K])."""
    assert x.dtype == torch.float32 and weight.dtype == torch.float32 and x.is_contiguous() and weight.is_contiguous()
    b, k = x.shape
    n = weight.shape[0]
    if out is None:
        out = torch.empty((b, n), dtype=torch.float32, device=x.device)
    if b == 0:
        return out
    lib = load_library()
    nbytes = lib.dfd_linear_f32_workspace_bytes(b, n)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    check(lib.dfd_linear_f32(ctx(x.device), ptr(x), ptr(weight), ptr(bias), ptr(residual), ptr(out), b, n, k,
                             1 if quick_gelu else 0, ptr(ws), nbytes, stream_ptr(x.device)))
    return out


def mean_std_array(mean, std):
    """HOST float[6] = {mean[3], std[3]} for the uint8 entry points."""
    return (c_float * 6)(*[float(v) for v in mean], *[float(v) for v in std])


def patchify(frames, patch, kp=None, mean=None, std=None):
    """frames fp32 [F,3,R,R] (or uint8 with per-channel ``mean`` / ``std``: (x/255 - mean)/std is applied on the
    fly) -> bf16 [F*(P+1), Kp] patch matrix with zero cls rows."""
    assert frames.dtype in (torch.float32, torch.uint8) and frames.is_contiguous() and frames.dim() == 4
    f, _, r, _ = frames.shape
    k = 3 * patch * patch
    kp = kp or (k + 63) // 64 * 64
    g = r // patch
    out = torch.empty((f * (g * g + 1), kp), dtype=torch.bfloat16, device=frames.device)
    if frames.dtype == torch.uint8:
        check(load_library().dfd_patchify_u8(ctx(frames.device), ptr(frames), mean_std_array(mean, std), ptr(out), f,
                                             r, patch, kp, stream_ptr(frames.device)))
    else:
        check(load_library().dfd_patchify(ctx(frames.device), ptr(frames), ptr(out), f, r, patch, kp,
                                          stream_ptr(frames.device)))
    return out


def mha_fwd(qkv, n_frames, seq, heads):
    """Encoder self-attention on a packed bf16 QKV buffer [n_frames*seq, 3*heads*64] -> mix bf16 [.., heads*64]."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    mix = torch.empty((n_frames * seq, heads * 64), dtype=torch.bfloat16, device=qkv.device)
    check(load_library().dfd_mha_fwd(ctx(qkv.device), ptr(qkv), ptr(mix), n_frames, seq, heads,
                                     stream_ptr(qkv.device)))
    return mix


def decoder_attention(qs, k, v, pos_emb, mask):
    """qs fp32 [B,H,128]; k, v bf16 [B,T,P,H,64] (any b/t/p strides, unit inner strides); mask bool [B,T]."""
    b, t, p, h, dh = k.shape
    assert dh == 64 and k.stride(4) == 1 and k.stride(3) == 64 and k.stride() == v.stride()
    assert k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
    m8 = mask.to(torch.uint8).contiguous()
    mix = torch.empty((b, h * 64), dtype=torch.float32, device=k.device)
    ws = torch.empty((2 * b * t * h * 130,), dtype=torch.float32, device=k.device)
    check(load_library().dfd_decoder_attention(
        ctx(k.device), ptr(qs.contiguous()), ptr(k), ptr(v), k.stride(0), k.stride(1), k.stride(2),
        ptr(None if pos_emb is None else pos_emb.contiguous()), ptr(m8), b, t, p, h, ptr(mix), ptr(ws),
        ws.numel() * 4, stream_ptr(k.device)))
    return mix


def decoder_attention_modes(qs, k, v, pos_emb, mask, attn_mode):
    """decoder_attention for op_mode.attn_mode: attn_mode = ATTN_FRAME | ATTN_TEMPORAL bit set."""
    b, t, p, h = _check_kv(k, v)
    lib = load_library()
    m8 = mask.to(torch.uint8).contiguous()
    mix = torch.empty((b, h * 64), dtype=torch.float32, device=k.device)
    nbytes = lib.dfd_decoder_attention_modes_workspace_bytes(b, t, p, h)
    ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=k.device)
    check(lib.dfd_decoder_attention_modes(
        ctx(k.device), ptr(qs.contiguous()), ptr(k), ptr(v), k.stride(0), k.stride(1), k.stride(2),
        ptr(None if pos_emb is None else pos_emb.contiguous()), ptr(m8), b, t, p, h, attn_mode, ptr(mix), ptr(ws),
        nbytes, stream_ptr(k.device)))
    return mix


def project_logits(feature, proj, scale=5.0):
    b, d = feature.shape
    o = proj.shape[1]
    out = torch.empty((b, o), dtype=torch.float32, device=feature.device)
    check(load_library().dfd_project_logits(ctx(feature.device), ptr(feature.contiguous()), ptr(proj.contiguous()),
                                            b, d, o, scale, ptr(out), stream_ptr(feature.device)))
    return out


def _check_kv(k, v):
    b, t, p, h, dh = k.shape
    assert dh == 64 and k.stride(4) == 1 and k.stride(3) == 64 and k.stride() == v.stride()
    assert k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
    return b, t, p, h


def decoder_attention_train(qs, k, v, pos_emb, mask):
    """decoder_attention that also returns the per-(clip, head) statistics [B,H,66] the backward pass needs."""
    b, t, p, h = _check_kv(k, v)
    lib = load_library()
    m8 = mask.to(torch.uint8).contiguous()
    mix = torch.empty((b, h * 64), dtype=torch.float32, device=k.device)
    stats = torch.empty((b, h, 66), dtype=torch.float32, device=k.device)
    nbytes = lib.dfd_decoder_attention_workspace_bytes(b, t, h)
    ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=k.device)
    check(lib.dfd_decoder_attention_train(
        ctx(k.device), ptr(qs.contiguous()), ptr(k), ptr(v), k.stride(0), k.stride(1), k.stride(2),
        ptr(None if pos_emb is None else pos_emb.contiguous()), ptr(m8), b, t, p, h, ptr(mix), ptr(stats), ptr(ws),
        nbytes, stream_ptr(k.device)))
    return mix, stats


def decoder_attention_backward(qs, k, v, pos_emb, mask, stats, dmix, need_kv_grad=False):
    """(dqs [B,H,128], dpos_emb [T,H,64] or None) for dmix [B,H*64]; with ``need_kv_grad`` also (dk, dv) fp32
    [B,T,P,H,64] (trainable adapter on the taps), appended to the result."""
    b, t, p, h = _check_kv(k, v)
    lib = load_library()
    m8 = mask.to(torch.uint8).contiguous()
    dqs = torch.empty((b, h, 128), dtype=torch.float32, device=k.device)
    dpe = None if pos_emb is None else torch.empty((t, h, 64), dtype=torch.float32, device=k.device)
    dk = torch.empty((b, t, p, h, 64), dtype=torch.float32, device=k.device) if need_kv_grad else None
    dv = torch.empty_like(dk) if need_kv_grad else None
    nbytes = lib.dfd_decoder_attention_workspace_bytes(b, t, h)
    ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=k.device)
    check(lib.dfd_decoder_attention_backward(
        ctx(k.device), ptr(qs.contiguous()), ptr(k), ptr(v), k.stride(0), k.stride(1), k.stride(2),
        ptr(None if pos_emb is None else pos_emb.contiguous()), ptr(m8), ptr(stats.contiguous()),
        ptr(dmix.contiguous().float()), b, t, p, h, ptr(dqs), ptr(dpe), ptr(dk), ptr(dv), ptr(ws), nbytes,
        stream_ptr(k.device)))
    if need_kv_grad:
        return dqs, dpe, dk, dv
    return dqs, dpe


def linear_f32_backward(x, weight, dy, gelu_pre=None, dx_add=None, need_dx=True, need_dw=True):
    """Backward of ``linear_f32`` (x [B,K], weight [N,K], dy [B,N]): returns (dx [B,K], dW [N,K], db [N])."""
    b, k = x.shape
    n = weight.shape[0]
    dev = x.device
    dx = torch.empty((b, k), dtype=torch.float32, device=dev) if need_dx else None
    dw = torch.empty((n, k), dtype=torch.float32, device=dev) if need_dw else None
    db = torch.empty((n,), dtype=torch.float32, device=dev) if need_dw else None
    lib = load_library()
    nbytes = lib.dfd_linear_f32_backward_workspace_bytes(b, n, k)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    check(lib.dfd_linear_f32_backward(ctx(dev), ptr(x), ptr(weight), ptr(dy), ptr(gelu_pre), ptr(dx_add), ptr(dx),
                                      ptr(dw), ptr(db), b, n, k, ptr(ws), nbytes, stream_ptr(dev)))
    return dx, dw, db


def adapter_apply(kind, kv, rows, ld, width, inner, w_down, w_mid, w_up, ln_weight, ln_bias, group_rows, group_skip,
                  workspace=None):
    """In-place CompInvAdapter on one bf16 tap: ``kv`` is a tensor whose data pointer is the first element of a
    ``[rows, width]`` matrix with row pitch ``ld`` elements (a K or V column slice of a packed QKV buffer).
    Weights: bf16 matrices, fp32 LayerNorm parameters. Returns the ``Workspace`` (reusable)."""
    lib = load_library()
    dev = kv.device
    assert kv.dtype == torch.bfloat16
    if rows == 0:
        return workspace
    nbytes = lib.dfd_adapter_workspace_bytes(kind, width, inner, rows)
    if workspace is None:
        workspace = Workspace()
    buf = workspace.get(nbytes, dev)
    w = AdapterWeights(ptr(w_down).value, ptr(w_mid).value, ptr(w_up).value, ptr(ln_weight).value, ptr(ln_bias).value)
    check(lib.dfd_adapter_apply(ctx(dev), kind, width, inner, ctypes.byref(w), ptr(kv), ld, rows, group_rows,
                                group_skip, ptr(buf), buf.numel(), stream_ptr(dev)))
    return workspace


def resize_crop_u8(frames, size, workspace=None):
    """``T.Resize(size, BICUBIC)`` + ``T.CenterCrop(size)`` of the loader transform (reference src/models.py:756-761)
    on the device: uint8 frames ``[..., 3, H, W]`` -> uint8 ``[..., 3, size, size]`` (within 1 LSB of torchvision)."""
    if frames.dtype != torch.uint8 or frames.dim() < 3 or frames.shape[-3] != 3:
        raise ValueError("resize_crop_u8 expects uint8 frames [..., 3, H, W], got %s %s" % (frames.dtype, tuple(frames.shape)))
    if frames.device.type != "cuda":
        raise NativeError("dfdclip_b200 resize needs CUDA tensors (no CPU fallback)")
    lead = tuple(frames.shape[:-3])
    h, w = int(frames.shape[-2]), int(frames.shape[-1])
    flat = frames.reshape(-1, 3, h, w).contiguous()
    n = flat.shape[0]
    out = torch.empty((n, 3, size, size), dtype=torch.uint8, device=frames.device)
    if n == 0:
        return out.view(lead + (3, size, size))
    lib = load_library()
    nbytes = lib.dfd_resize_crop_u8_workspace_bytes(n, h, w, size)
    if nbytes == 0:
        raise NativeError("unsupported resize %dx%d -> %d" % (h, w, size))
    if workspace is None:
        workspace = Workspace()
    buf = workspace.get(nbytes, frames.device)
    check(lib.dfd_resize_crop_u8(ctx(frames.device), ptr(flat), n, h, w, size, ptr(out), ptr(buf), buf.numel(),
                                 stream_ptr(frames.device)))
    return out.view(lead + (3, size, size))


def ema_frames(x, ratio):
    """op_mode.ema_frame: x fp32 [B,T,...] -> [B,1,...], the reference's EMA recurrence over the frame axis."""
    assert x.dtype == torch.float32 and x.dim() >= 3
    x = x.contiguous()
    b, t = x.shape[:2]
    out = torch.empty((b, 1) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    elems = 1
    for n in x.shape[2:]:
        elems *= int(n)
    check(load_library().dfd_ema_frames(ctx(x.device), ptr(x), ptr(out), b, t, elems, float(ratio),
                                        stream_ptr(x.device)))
    return out

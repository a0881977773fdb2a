"""Drop-in for the part of the reference's ``src/clip/model.py`` that DFD-CLIP uses: the CLIP ViT visual tower
with per-layer q/k/v taps (reference: VisionTransformer src/clip/model.py:254-294, Transformer :229-251,
ResidualAttentionBlock :202-226, MultiheadAttention :171-199) and ``build_model`` (:453-496).

The modules below only HOLD parameters (same names, shapes and ``state_dict()`` keys as the reference, so the
decoder initialiser and ``*_weights.pt`` checkpoints work unchanged); the computation is one call into
libdfdclip_b200.so (``dfd_encoder_forward``): tcgen05 GEMMs, flash-style attention, vectorised LayerNorm.
There is no PyTorch or CPU fallback: calling ``forward`` without a B200 and the built library raises.
"""
import ctypes
from collections import OrderedDict

import torch
from torch import nn

from .. import _native


MAX_TOKENS_PER_FRAME = 257   # csrc/encoder.cu vit_shape: a frame's keys stay on chip in the attention kernels


class LayerNorm(nn.LayerNorm):
    """Parameter holder for a LayerNorm (eps 1e-5); also usable as a module (fp32 math like :157-163)."""

    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class MultiheadAttention(nn.Module):
    """Parameter holder: fused in_proj ([3D, D] weight, [3D] bias) and out_proj (reference :171-183)."""

    def __init__(self, embed_dim, n_head):
        super().__init__()
        # the reference leaves these uninitialised (torch.empty, :179-180); zeros keep a fresh module finite
        self.in_proj_weight = nn.Parameter(torch.zeros((3 * embed_dim, embed_dim)))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.n_head = n_head


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """CLIP ViT frame encoder. ``forward(x[N,3,R,R], with_out=False, with_q=False)`` returns, like the reference
    (:276-294, :236-251), one dict per layer with ``k``, ``v`` (and ``q`` / ``out`` on request), each k/v/q a
    ``[N, L, H, 64]`` tensor in fp32 like the reference's (:186-199; ``tap_dtype``), converted from that layer's packed
    bf16 ``[N*L, 3D]`` QKV buffer — the B200 path's compute dtype, which ``encode`` hands out without a copy and which
    ``Detector`` reads in place. ``out`` is the fp32 residual stream."""

    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.width = width
        self.layers = layers
        self.heads = heads
        self.patch_size = patch_size
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        # present in checkpoints, never applied on this path (reference :273-274 vs :294)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._packed = None
        self._packed_key = None
        self._workspace = _native.Workspace()
        self._resize_workspace = _native.Workspace()
        # dtype of the q/k/v tensors ``forward`` returns: fp32 as in the reference; torch.bfloat16 hands out views of
        # the packed buffers instead (no conversion pass)
        self.tap_dtype = torch.float32
        # normalisation applied on the fly to uint8 frames: the constants of the reference's CLIP transform
        # (src/models.py:762-768); Detector overrides them when its transform uses other values
        self.input_mean = (0.48145466, 0.4578275, 0.40821073)
        self.input_std = (0.26862954, 0.26130258, 0.27577711)

    # ---------------------------------------------------------------------------------------------- native
    @property
    def tokens_per_frame(self):
        return (self.input_resolution // self.patch_size) ** 2 + 1

    def _dims(self):
        return _native.VitDims(self.input_resolution, self.patch_size, self.width, self.heads, self.layers)

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in self.parameters())

    def invalidate_packed_weights(self):
        self._packed = None
        self._packed_key = None

    def _packed_weights(self):
        """bf16/fp32 packed copy of the parameters in the library's layout, rebuilt when a parameter changes."""
        key = self._weights_key()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise _native.NativeError(
                "dfdclip_b200 VisionTransformer has no CPU path: move the module to a B200 (`.to('cuda')`) first")
        lib = _native.load_library()
        dims = self._dims()
        nbytes = lib.dfd_encoder_packed_bytes(ctypes.byref(dims))
        if nbytes == 0:
            raise _native.NativeError("unsupported ViT shape for the B200 path: %s" % _native.load_library()
                                      .dfd_last_error().decode())
        params = []  # keep fp32 contiguous views alive until the pack kernels have run

        def f32(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            params.append(t)
            return t

        blocks = self.transformer.resblocks
        w = _native.VitWeights()
        w.conv1_weight = f32(self.conv1.weight).data_ptr()
        w.class_embedding = f32(self.class_embedding).data_ptr()
        w.positional_embedding = f32(self.positional_embedding).data_ptr()
        w.ln_pre_weight = f32(self.ln_pre.weight).data_ptr()
        w.ln_pre_bias = f32(self.ln_pre.bias).data_ptr()
        arrays = []

        def per_layer(getter):
            arr = _native.ptr_array([f32(getter(b)) for b in blocks])
            arrays.append(arr)
            return ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))

        w.ln_1_weight = per_layer(lambda b: b.ln_1.weight)
        w.ln_1_bias = per_layer(lambda b: b.ln_1.bias)
        w.in_proj_weight = per_layer(lambda b: b.attn.in_proj_weight)
        w.in_proj_bias = per_layer(lambda b: b.attn.in_proj_bias)
        w.out_proj_weight = per_layer(lambda b: b.attn.out_proj.weight)
        w.out_proj_bias = per_layer(lambda b: b.attn.out_proj.bias)
        w.ln_2_weight = per_layer(lambda b: b.ln_2.weight)
        w.ln_2_bias = per_layer(lambda b: b.ln_2.bias)
        w.c_fc_weight = per_layer(lambda b: b.mlp.c_fc.weight)
        w.c_fc_bias = per_layer(lambda b: b.mlp.c_fc.bias)
        w.c_proj_weight = per_layer(lambda b: b.mlp.c_proj.weight)
        w.c_proj_bias = per_layer(lambda b: b.mlp.c_proj.bias)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.dfd_encoder_pack_weights(_native.ctx(dev), ctypes.byref(dims), ctypes.byref(w),
                                                       _native.ptr(packed), _native.stream_ptr(dev)))
            torch.cuda.current_stream(dev).synchronize()  # params/arrays may be temporaries
        self._packed = packed
        self._packed_key = key
        return packed

    def _get_workspace(self, nbytes, dev):
        return self._workspace.get(nbytes, dev)

    def encode(self, x, keep_layers=None, need_out=False, last_qkv_only=None, qkv_into=None, frame_offset=0):
        """``qkv_into`` (optional): dict layer -> preallocated bf16 ``[N_total*L, 3D]`` buffer; the rows of these
        ``N`` frames are written at frame ``frame_offset`` (lets a caller encode a batch chunk by chunk into one set
        of tap buffers and run the decoder once).

        Run the encoder on frames x[N,3,R,R] (fp32, cuda). Returns ``(qkv, outs)``: ``qkv[l]`` is the packed
        bf16 ``[N*L, 3D]`` buffer of layer l for every l in ``keep_layers`` (default: all layers), ``outs[l]`` the
        fp32 residual stream after layer l when ``need_out``. Layers after the last kept one are not executed, and
        the last kept layer stops after its K/V projection unless ``need_out`` or ``last_qkv_only=False`` (dead-work
        skipping): the Q block of that layer's buffer is then left unwritten."""
        plan = self._encode_plan(x, keep_layers, need_out, last_qkv_only, qkv_into, frame_offset)
        if plan["n"] == 0:
            return plan["qkv"], plan["outs"]
        lib, dev = _native.load_library(), plan["dev"]
        with torch.cuda.device(dev):
            if plan["x"].dtype == torch.uint8:
                _native.check(lib.dfd_encoder_forward_u8(
                    _native.ctx(dev), ctypes.byref(plan["dims"]), _native.ptr(plan["packed"]), _native.ptr(plan["x"]),
                    plan["mean_std"], plan["n"], plan["run_layers"], plan["qkv_only"], plan["qkv_pp"], plan["out_pp"],
                    _native.ptr(plan["ws"]), plan["ws_bytes"], _native.stream_ptr(dev)))
            else:
                _native.check(lib.dfd_encoder_forward(
                    _native.ctx(dev), ctypes.byref(plan["dims"]), _native.ptr(plan["packed"]), _native.ptr(plan["x"]),
                    plan["n"], plan["run_layers"], plan["qkv_only"], plan["qkv_pp"], plan["out_pp"],
                    _native.ptr(plan["ws"]), plan["ws_bytes"], _native.stream_ptr(dev)))
        return plan["qkv"], plan["outs"]

    def _encode_plan(self, x, keep_layers=None, need_out=False, last_qkv_only=None, qkv_into=None, frame_offset=0):
        """Argument checking, output / workspace buffers and pointer arrays of one encoder pass (everything ``encode``
        does short of the launch; ``Detector.predict`` hands the same plan to ``dfd_predict_forward``)."""
        if x.dim() == 4 and x.dtype == torch.uint8 and x.shape[1] == 3 and x.device.type == "cuda" and \
                (x.shape[2] != self.input_resolution or x.shape[3] != self.input_resolution):
            # raw decoded frames of another size: the loader's Resize(BICUBIC) + CenterCrop on the device
            # (reference src/models.py:756-761), then the uint8 path below
            x = _native.resize_crop_u8(x, self.input_resolution, self._resize_workspace)
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.input_resolution or x.shape[3] != self.input_resolution:
            raise ValueError("expected frames of shape [N,3,%d,%d], got %s" %
                             (self.input_resolution, self.input_resolution, tuple(x.shape)))
        if x.device.type != "cuda":
            raise _native.NativeError("dfdclip_b200 encoder needs CUDA tensors (no CPU fallback)")
        lib = _native.load_library()
        dev = x.device
        x = x.detach()
        if x.dtype == torch.uint8:
            # raw pixels: ConvertImageDtype(float32) + Normalize of the data loader are fused into the patchify kernel
            x = x.contiguous()
        elif x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        n = x.shape[0]
        seq, d = self.tokens_per_frame, self.width
        keep = sorted(set(range(self.layers) if keep_layers is None else keep_layers))
        if keep and (keep[0] < 0 or keep[-1] >= self.layers):
            raise IndexError("layer index out of range: %s" % keep)
        if need_out:
            run_layers, qkv_only = self.layers, False
        else:
            run_layers = (keep[-1] + 1) if keep else 0
            qkv_only = True if last_qkv_only is None else bool(last_qkv_only)
        packed = self._packed_weights()
        dims = self._dims()
        if qkv_into is None:
            qkv = {l: torch.empty((n * seq, 3 * d), dtype=torch.bfloat16, device=dev) for l in keep}
        else:
            qkv = {}
            for l in keep:
                buf = qkv_into[l]
                if buf.dtype != torch.bfloat16 or buf.dim() != 2 or buf.shape[1] != 3 * d or not buf.is_contiguous() \
                        or buf.device != dev or buf.shape[0] < (frame_offset + n) * seq:
                    raise ValueError("qkv_into[%d] must be a contiguous bf16 [>=%d, %d] buffer on %s" %
                                     (l, (frame_offset + n) * seq, 3 * d, dev))
                qkv[l] = buf.narrow(0, frame_offset * seq, n * seq)
        outs = {l: torch.empty((n, seq, d), dtype=torch.float32, device=dev) for l in range(self.layers)} if need_out else {}
        plan = dict(n=n, dev=dev, x=x, qkv=qkv, outs=outs, run_layers=run_layers, qkv_only=1 if qkv_only else 0,
                    packed=packed, dims=dims)
        if n == 0:
            return plan
        ws_bytes = lib.dfd_encoder_workspace_bytes(ctypes.byref(dims), n)
        plan["ws_bytes"], plan["ws"] = ws_bytes, self._get_workspace(ws_bytes, dev)
        qkv_arr = _native.ptr_array([qkv.get(l) for l in range(self.layers)])
        out_arr = _native.ptr_array([outs.get(l) for l in range(self.layers)]) if need_out else None
        plan["keep_alive"] = (qkv_arr, out_arr)
        plan["qkv_pp"] = ctypes.cast(qkv_arr, ctypes.POINTER(ctypes.c_void_p))
        plan["out_pp"] = ctypes.cast(out_arr, ctypes.POINTER(ctypes.c_void_p)) if need_out else None
        plan["mean_std"] = _native.mean_std_array(self.input_mean, self.input_std) if x.dtype == torch.uint8 else None
        return plan

    def forward(self, x, with_out=False, with_q=False):
        n = x.shape[0]
        seq, h = self.tokens_per_frame, self.heads
        qkv, outs = self.encode(x, keep_layers=None, need_out=with_out, last_qkv_only=False)
        kvs = []
        for l in range(self.layers):
            view = qkv[l].view(n, seq, 3, h, 64)
            a = dict(k=view[:, :, 1].to(self.tap_dtype), v=view[:, :, 2].to(self.tap_dtype))
            if with_q:
                a["q"] = view[:, :, 0].to(self.tap_dtype)
            if with_out:
                a["out"] = outs[l]
            kvs.append(a)
        return kvs


class CLIP(nn.Module):
    """Holder returned by ``build_model``: only ``.visual`` is implemented (DFD-CLIP takes ``clip.load(..)[0].visual``,
    src/models.py:440). The text tower / ResNet towers of the reference are out of scope and raise."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("ModifiedResNet towers are not part of the DFD-CLIP hot path")
        vision_heads = vision_width // 64
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers, vision_heads,
                                        embed_dim)

    def encode_image(self, image):
        raise NotImplementedError("CLIP.encode_image (ln_post/proj head) is never used by DFD-CLIP")

    def encode_text(self, text):
        raise NotImplementedError("the CLIP text tower is out of scope of dfdclip_b200")

    def forward(self, image, text):
        raise NotImplementedError("CLIP.forward is out of scope of dfdclip_b200")


def _fp16_round_(module):
    """Same rounding ``convert_weights`` applies (reference :429-450): Conv/Linear weights and biases and ``proj``
    go through fp16; in_proj_* of the custom attention, LayerNorm and embeddings stay fp32."""
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)) and not isinstance(m, nn.LayerNorm):
                m.weight.copy_(m.weight.half().float())
                if m.bias is not None:
                    m.bias.copy_(m.bias.half().float())
            if isinstance(getattr(m, "proj", None), torch.Tensor):
                m.proj.copy_(m.proj.half().float())


def build_model(state_dict):
    """Build the (visual-only) CLIP model from a checkpoint state dict, inferring every size from tensor shapes
    like the reference (:453-496)."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("only ViT CLIP checkpoints are supported (no 'visual.proj' in the state dict)")
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    image_resolution = vision_patch_size * grid_size
    embed_dim = state_dict["visual.proj"].shape[1]
    if grid_size * grid_size + 1 > MAX_TOKENS_PER_FRAME:  # fail at load time, not at the first encode
        raise NotImplementedError("%d tokens per frame (resolution %d, patch %d): the B200 attention kernels support at "
                                  "most %d" % (grid_size * grid_size + 1, image_resolution, vision_patch_size,
                                               MAX_TOKENS_PER_FRAME))
    model = CLIP(embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size)
    visual = OrderedDict((k[len("visual."):], v.float()) for k, v in state_dict.items() if k.startswith("visual."))
    model.visual.load_state_dict(visual)
    _fp16_round_(model.visual)
    return model.eval()

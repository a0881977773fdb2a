"""Seeded synthetic weights and clips (there is no network for CLIP checkpoints or FF++ videos).

Every tensor is drawn from its own ``torch.Generator`` keyed by (seed, parameter name), on the CPU, so the same
values come out in this container, on the GPU box, in the golden-vector generator and in the bench, regardless
of creation order. Key names and shapes follow the reference's ``Detector.state_dict()`` and
CLIP checkpoint (``visual.*``) schemas; parameters the reference's ``build_model`` rounds to fp16
(conv/linear weights and biases, ``proj``; src/clip/model.py:429-450) are fp16-representable here too, exactly
like real CLIP checkpoints.
"""
import zlib
from collections import OrderedDict

import torch

# name -> (image_size, patch, width, heads, layers, output_dim)
VIT_CONFIGS = {
    "ViT-B/32": (224, 32, 768, 12, 12, 512),
    "ViT-B/16": (224, 16, 768, 12, 12, 512),
    "ViT-L/14": (224, 14, 1024, 16, 24, 768),
    # small shapes for fast parity tests (same code path, width still a multiple of 256)
    "tiny-256x4": (32, 16, 256, 4, 4, 64),
    "small-512x6": (64, 16, 512, 8, 6, 128),
}


def vit_dims(arch):
    if arch not in VIT_CONFIGS:
        raise KeyError("unknown synthetic architecture %r (have %s)" % (arch, sorted(VIT_CONFIGS)))
    keys = ("image_size", "patch_size", "width", "heads", "layers", "output_dim")
    return dict(zip(keys, VIT_CONFIGS[arch]))


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def _randn(seed, name, shape, std=1.0, mean=0.0, fp16=False):
    t = torch.randn(shape, generator=_gen(seed, name), dtype=torch.float32) * std + mean
    return t.half().float() if fp16 else t


def visual_state_dict(arch, seed=0):
    """fp32 state dict of the CLIP visual tower (keys as in ``VisionTransformer.state_dict()``)."""
    d = vit_dims(arch)
    r, p, w, layers, out = d["image_size"], d["patch_size"], d["width"], d["layers"], d["output_dim"]
    tokens = (r // p) ** 2 + 1
    sd = OrderedDict()
    scale = w ** -0.5
    sd["class_embedding"] = _randn(seed, "class_embedding", (w,), scale)
    sd["positional_embedding"] = _randn(seed, "positional_embedding", (tokens, w), scale)
    sd["proj"] = _randn(seed, "proj", (w, out), scale, fp16=True)
    sd["conv1.weight"] = _randn(seed, "conv1.weight", (w, 3, p, p), (3 * p * p) ** -0.5, fp16=True)
    for ln in ("ln_pre", "ln_post"):
        sd[ln + ".weight"] = _randn(seed, ln + ".weight", (w,), 0.1, 1.0)
        sd[ln + ".bias"] = _randn(seed, ln + ".bias", (w,), 0.1)
    attn_std = w ** -0.5
    proj_std = (w ** -0.5) * ((2 * layers) ** -0.5)
    fc_std = (2 * w) ** -0.5
    for i in range(layers):
        pre = "transformer.resblocks.%d." % i
        # in_proj_* are NOT fp16-rounded by the reference's convert_weights (custom attention module)
        sd[pre + "attn.in_proj_weight"] = _randn(seed, pre + "attn.in_proj_weight", (3 * w, w), attn_std)
        sd[pre + "attn.in_proj_bias"] = _randn(seed, pre + "attn.in_proj_bias", (3 * w,), 0.02)
        sd[pre + "attn.out_proj.weight"] = _randn(seed, pre + "attn.out_proj.weight", (w, w), proj_std, fp16=True)
        sd[pre + "attn.out_proj.bias"] = _randn(seed, pre + "attn.out_proj.bias", (w,), 0.02, fp16=True)
        for ln in ("ln_1", "ln_2"):
            sd[pre + ln + ".weight"] = _randn(seed, pre + ln + ".weight", (w,), 0.1, 1.0)
            sd[pre + ln + ".bias"] = _randn(seed, pre + ln + ".bias", (w,), 0.1)
        sd[pre + "mlp.c_fc.weight"] = _randn(seed, pre + "mlp.c_fc.weight", (4 * w, w), fc_std, fp16=True)
        sd[pre + "mlp.c_fc.bias"] = _randn(seed, pre + "mlp.c_fc.bias", (4 * w,), 0.02, fp16=True)
        sd[pre + "mlp.c_proj.weight"] = _randn(seed, pre + "mlp.c_proj.weight", (w, 4 * w), proj_std, fp16=True)
        sd[pre + "mlp.c_proj.bias"] = _randn(seed, pre + "mlp.c_proj.bias", (w,), 0.02, fp16=True)
    return sd


def clip_checkpoint_state_dict(arch, seed=0):
    """A CLIP checkpoint state dict: ``visual.*`` plus the smallest text tower ``build_model`` accepts
    (src/clip/model.py:453-496 infers every size from shapes; the text tower is never used by DFD-CLIP)."""
    sd = OrderedDict(("visual." + k, v) for k, v in visual_state_dict(arch, seed).items())
    embed = vit_dims(arch)["output_dim"]
    tw, ctx, vocab = 64, 8, 32
    sd["positional_embedding"] = _randn(seed, "t.positional_embedding", (ctx, tw), 0.01)
    sd["text_projection"] = _randn(seed, "t.text_projection", (tw, embed), tw ** -0.5, fp16=True)
    sd["logit_scale"] = torch.tensor(2.6593)
    sd["token_embedding.weight"] = _randn(seed, "t.token_embedding.weight", (vocab, tw), 0.02)
    sd["ln_final.weight"] = torch.ones(tw)
    sd["ln_final.bias"] = torch.zeros(tw)
    pre = "transformer.resblocks.0."
    sd[pre + "attn.in_proj_weight"] = _randn(seed, "t.in_proj_weight", (3 * tw, tw), tw ** -0.5, fp16=True)
    sd[pre + "attn.in_proj_bias"] = torch.zeros(3 * tw)
    sd[pre + "attn.out_proj.weight"] = _randn(seed, "t.out_proj.weight", (tw, tw), tw ** -0.5, fp16=True)
    sd[pre + "attn.out_proj.bias"] = torch.zeros(tw)
    for ln in ("ln_1", "ln_2"):
        sd[pre + ln + ".weight"] = torch.ones(tw)
        sd[pre + ln + ".bias"] = torch.zeros(tw)
    sd[pre + "mlp.c_fc.weight"] = _randn(seed, "t.c_fc.weight", (4 * tw, tw), tw ** -0.5, fp16=True)
    sd[pre + "mlp.c_fc.bias"] = torch.zeros(4 * tw)
    sd[pre + "mlp.c_proj.weight"] = _randn(seed, "t.c_proj.weight", (tw, 4 * tw), tw ** -0.5, fp16=True)
    sd[pre + "mlp.c_proj.bias"] = torch.zeros(tw)
    return sd


def layer_indices(arch, decode_mode="stride", decode_stride=2, decode_indices=()):
    """Tapped encoder layers (src/models.py:458-461)."""
    layers = vit_dims(arch)["layers"]
    if decode_mode == "stride":
        return list(range(0, layers, decode_stride))
    return list(decode_indices)


def decoder_state_dict(arch, num_frames, out_dims=(2,), taps=None, seed=0, visual=None, aug_query=False,
                       global_prediction=False, temporal_position=True):
    """fp32 state dict of the temporal decoder (keys as in ``Decoder.state_dict()``), initialised the way the
    reference does it: ln_1 / ln_2 / mlp of block i copied from encoder layer taps[i] (src/models.py:178-229),
    everything else random."""
    d = vit_dims(arch)
    w, heads = d["width"], d["heads"]
    taps = layer_indices(arch) if taps is None else list(taps)
    visual = visual_state_dict(arch, seed) if visual is None else visual
    scale = w ** -0.5
    sd = OrderedDict()
    sd["class_embedding"] = _randn(seed, "dec.class_embedding", (w,), scale)
    if temporal_position:
        sd["positional_embedding"] = _randn(seed, "dec.positional_embedding", (num_frames, 1, heads, w // heads), scale)
    for i, o in enumerate(out_dims):
        if global_prediction:  # op_mode.global_prediction: one projection per tapped layer (src/models.py:309-313)
            for layer in taps:
                name = "proj%dx%d_L%d" % (i, o, layer)
                sd[name] = _randn(seed, "dec." + name, (w, o), scale)
        else:
            sd["proj%dx%d" % (i, o)] = _randn(seed, "dec.proj%dx%d" % (i, o), (w, o), scale)
    if aug_query:  # op_mode.aug_query (src/models.py:250-255); zero-initialised there, random here
        for i in range(len(taps) - 1):
            sd["transformer.augment_query_%d" % i] = _randn(seed, "dec.augment_query_%d" % i, (w,), 0.5)
    for ln in ("ln_pre", "ln_post"):
        sd[ln + ".weight"] = _randn(seed, "dec." + ln + ".weight", (w,), 0.1, 1.0)
        sd[ln + ".bias"] = _randn(seed, "dec." + ln + ".bias", (w,), 0.1)
    for i, layer in enumerate(taps):
        pre = "transformer.resblocks.%d." % i
        src = "transformer.resblocks.%d." % layer
        sd[pre + "attn.in_proj.weight"] = _randn(seed, "dec." + pre + "in_proj.weight", (2 * w, w), 1.5 * w ** -0.5)
        sd[pre + "attn.in_proj.bias"] = _randn(seed, "dec." + pre + "in_proj.bias", (2 * w,), 0.05)
        sd[pre + "attn.out_proj.weight"] = _randn(seed, "dec." + pre + "out_proj.weight", (w, w), w ** -0.5)
        sd[pre + "attn.out_proj.bias"] = _randn(seed, "dec." + pre + "out_proj.bias", (w,), 0.02)
        for name in ("ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                     "mlp.c_proj.weight", "mlp.c_proj.bias"):
            sd[pre + name] = visual[src + name].clone()
    return sd


# adapter.struct.type -> Sequential indices of (first Linear, LayerNorm or None, middle Linear or None, last Linear)
ADAPTER_LAYOUT = {
    "768-x-768": (0, 2, None, 4),
    "legacy-768-x-768": (0, 2, None, 3),
    "768-x-768-nln": (0, 1, None, 4),
    "768-x-768-ln": (0, 1, None, 4),
    "768-x-768-z0": (0, 1, None, 4),
    "768-xxx-768": (0, None, 3, 6),
    "linear": (0, None, None, None),
}


def adapter_state_dict(arch, n_taps, struct_type, inner=256, seed=0, num_frames=None):
    """fp32 state dict of a ``CompInvAdapter`` (keys ``l{i}_{k|v}.{idx}.{weight|bias}``, src/models.py:783-919) with
    random, non-degenerate weights. The up-projection is scaled so that the adapter's contribution is about a
    quarter of the tap it is added to: the shipped structs start from the identity ("-z0", :856-858) and learn a
    correction, and a bf16 path cannot hold the north_star logit tolerance against fp32 when a random map as large
    as the tap itself is stacked on the taps (measured: the reference's own fp32 adapter applied to bf16-rounded
    taps already misses it at |delta| = 0.65 |tap|)."""
    d = vit_dims(arch)
    w = d["width"]
    patches = (d["image_size"] // d["patch_size"]) ** 2
    sd = OrderedDict()
    if struct_type == "768-bn":
        # Linear(w, w, bias=False) -> BatchNorm2d(num_frames) over [b, t, p, w] (channel = frame index, :877-887):
        # non-trivial affine and running statistics so that the eval-mode normalisation is exercised
        assert num_frames is not None, "the 768-bn struct needs num_frames"
        for i in range(n_taps):
            for j in ("k", "v"):
                pre = "l%d_%s." % (i, j)
                sd[pre + "0.weight"] = _randn(seed, "ad." + pre + "bn.lin", (w, w), 0.25 * w ** -0.5)
                sd[pre + "1.weight"] = _randn(seed, "ad." + pre + "bn.w", (num_frames,), 0.1, 1.0)
                sd[pre + "1.bias"] = _randn(seed, "ad." + pre + "bn.b", (num_frames,), 0.05)
                sd[pre + "1.running_mean"] = _randn(seed, "ad." + pre + "bn.rm", (num_frames,), 0.05)
                sd[pre + "1.running_var"] = _randn(seed, "ad." + pre + "bn.rv", (num_frames,), 0.1, 1.0).abs() + 0.1
                sd[pre + "1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        return sd
    first, ln, mid, last = ADAPTER_LAYOUT[struct_type]
    for i in range(n_taps):
        for j in ("k", "v"):
            pre = "l%d_%s." % (i, j)
            if struct_type == "linear":
                sd[pre + "0.weight"] = torch.eye(w) + _randn(seed, "ad." + pre + "0", (w, w), 0.2 * w ** -0.5)
                continue
            sd[pre + "%d.weight" % first] = _randn(seed, "ad." + pre + "down", (inner, w), w ** -0.5)
            if ln is not None:
                shape = (patches, inner) if struct_type == "768-x-768-nln" else (inner,)
                sd[pre + "%d.weight" % ln] = _randn(seed, "ad." + pre + "ln.w", shape, 0.1, 1.0)
                sd[pre + "%d.bias" % ln] = _randn(seed, "ad." + pre + "ln.b", shape, 0.1)
            if mid is not None:
                sd[pre + "%d.weight" % mid] = _randn(seed, "ad." + pre + "mid", (inner, inner), 1.5 * inner ** -0.5)
            sd[pre + "%d.weight" % last] = _randn(seed, "ad." + pre + "up", (w, inner), 0.4 * inner ** -0.5)
    return sd


def detector_state_dict(arch, num_frames, out_dims=(2,), taps=None, seed=0, adapter=None, adapter_inner=256,
                        ranking=False, **decoder_options):
    """Full ``Detector.state_dict()`` (encoder.* + decoder.* [+ adapter.*]), the on-disk format of ``*_weights.pt``
    (written at main.py:119-129, loaded strictly at inference.py:99). ``adapter`` = an
    ``adapter.struct.type`` string adds the CompInvAdapter parameters."""
    visual = visual_state_dict(arch, seed)
    sd = OrderedDict(("encoder." + k, v) for k, v in visual.items())
    for k, v in decoder_state_dict(arch, num_frames, out_dims, taps, seed, visual, **decoder_options).items():
        sd["decoder." + k] = v
    if ranking:  # train_mode.temporal == "ranking" (src/models.py:488-492)
        w = vit_dims(arch)["width"]
        sd["ranking_transform_param"] = _randn(seed, "ranking_transform_param", (w, 1), w ** -0.5)
    if adapter is not None:
        n_taps = len(layer_indices(arch) if taps is None else list(taps))
        for k, v in adapter_state_dict(arch, n_taps, adapter, adapter_inner, seed, num_frames=num_frames).items():
            sd["adapter." + k] = v
    return sd


def make_clips(batch, num_frames, image_size, seed=7, masked_tail=True):
    """x fp32 [B,T,3,R,R] standing for normalised frames, m bool [B,T]. With ``masked_tail`` clip 1 loses its last
    two frames and every 5th clip its last frame (padding of short clips, src/datasets.py:681-682)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, num_frames, 3, image_size, image_size, generator=g, dtype=torch.float32)
    m = torch.ones(batch, num_frames, dtype=torch.bool)
    if masked_tail and num_frames > 2:
        if batch > 1:
            m[1, -2:] = False
        m[4::5, -1] = False
    return x, m


def make_varied_clips(batch, num_frames, image_size, seed=7, video_ids=None, masked_tail=True):
    """Clips with content, unlike ``make_clips``' iid noise: every video has its own coarse colour layout, every clip
    its own mid-scale texture, contrast and brightness, every frame its own pixel noise — so the video features (and
    the class margins of a random head) differ from clip to clip the way they do on real videos. Built from exact
    operations only (block replication, elementwise fp32 multiply / add), so the same bits come out on every host.
    ``video_ids`` [batch] groups clips into videos (default: one video per clip). Returns x fp32 [B,T,3,R,R] and
    m bool [B,T] with the same trailing-frame padding pattern as ``make_clips``."""
    r = image_size
    coarse_n, mid_n = max(1, r // 32), max(1, r // 8)
    video_ids = list(range(batch)) if video_ids is None else [int(v) for v in video_ids]
    x = torch.empty(batch, num_frames, 3, r, r, dtype=torch.float32)
    for b in range(batch):
        gv = _gen(seed, "video%d" % video_ids[b])
        coarse = torch.randn(3, coarse_n, coarse_n, generator=gv)
        gc = _gen(seed, "clip%d" % b)
        mid = torch.randn(3, mid_n, mid_n, generator=gc)
        contrast = 0.5 + torch.rand((), generator=gc)
        brightness = 0.5 * torch.randn((), generator=gc)
        noise_std = 0.25 + 0.5 * torch.rand((), generator=gc)
        base = coarse.repeat_interleave(r // coarse_n, 1).repeat_interleave(r // coarse_n, 2) \
            + 0.5 * mid.repeat_interleave(r // mid_n, 1).repeat_interleave(r // mid_n, 2)
        noise = torch.randn(num_frames, 3, r, r, generator=gc)
        x[b] = (base[None] + noise * noise_std) * contrast + brightness
    m = torch.ones(batch, num_frames, dtype=torch.bool)
    if masked_tail and num_frames > 2:
        if batch > 1:
            m[1, -2:] = False
        m[4::5, -1] = False
    return x, m

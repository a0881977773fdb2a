"""ORACLE — test infrastructure only. CPU (torch fp32) restatement of the reference's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline; the product package never does.

The algorithm restated here is the reference's own PyTorch code (the reference is pure Python; the arithmetic is
torch's): ``/root/reference/src/clip/model.py`` and ``/root/reference/src/models.py``. Each function cites the
lines it follows. The restatement is *pinned*: ``oracle/gen_golden.py`` runs the UNMODIFIED reference in the build
container on seeded synthetic weights/clips and commits its outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this module against those vectors (fp32, ~1e-5). The reference itself ships no
tests, golden vectors or fixtures for this path.

Inputs are a ``Detector.state_dict()``-style dict of fp32 tensors so that the oracle and the
CUDA path share weights by value.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, sd, prefix):
    # LayerNorm.forward, src/clip/model.py:157-163 / src/models.py:58-68: fp32, eps = 1e-5 (torch default)
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _quick_gelu(x):
    # QuickGELU.forward, src/clip/model.py:166-168
    return x * torch.sigmoid(1.702 * x)


def infer_vit_dims(sd, prefix="encoder."):
    """Sizes from tensor shapes, as build_model does (src/clip/model.py:456-464)."""
    width = sd[prefix + "conv1.weight"].shape[0]
    patch = sd[prefix + "conv1.weight"].shape[-1]
    tokens = sd[prefix + "positional_embedding"].shape[0]
    grid = round((tokens - 1) ** 0.5)
    layers = len([k for k in sd if k.startswith(prefix) and k.endswith(".attn.in_proj_weight")])
    return dict(width=width, patch_size=patch, image_size=grid * patch, heads=width // 64, layers=layers,
                tokens=tokens)


def encoder_forward(sd, x, prefix="encoder.", with_out=False, with_q=False, num_layers=None):
    """VisionTransformer.forward (src/clip/model.py:276-294) + Transformer.forward (:236-251) +
    ResidualAttentionBlock.forward (:220-226) + MultiheadAttention.forward (:185-199).
    x: fp32 [N,3,R,R]. Returns a list (one entry per layer) of dicts with k, v [N,L,H,dh] (and q / out)."""
    dims = infer_vit_dims(sd, prefix)
    heads, width, patch = dims["heads"], dims["width"], dims["patch_size"]
    layers = dims["layers"] if num_layers is None else num_layers
    n = x.shape[0]
    # :277-279 conv1 (stride = kernel = patch, no bias) -> [N, grid^2, width]
    tok = F.conv2d(x.float(), sd[prefix + "conv1.weight"], None, stride=patch).flatten(2).transpose(1, 2)
    # :280-291 prepend the class embedding, add the positional embedding
    cls = sd[prefix + "class_embedding"].view(1, 1, width).expand(n, 1, width)
    h = torch.cat([cls, tok], dim=1) + sd[prefix + "positional_embedding"]
    h = _ln(h, sd, prefix + "ln_pre")  # :292
    kvs = []
    for i in range(layers):
        p = "%stransformer.resblocks.%d." % (prefix, i)
        u = _ln(h, sd, p + "ln_1")                                                      # :221
        qkv = F.linear(u, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])     # :186
        q, k, v = (t.reshape(n, -1, heads, width // heads) for t in qkv.chunk(3, dim=-1))  # :188-191
        aff = torch.einsum("nqhc,nkhc->nqkh", q / math.sqrt(q.shape[-1]), k)            # :193
        aff = aff.softmax(dim=-2)                                                       # :194 (over keys)
        mix = torch.einsum("nqlh,nlhc->nqhc", aff, v)                                   # :195
        h = h + F.linear(mix.flatten(-2), sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])  # :197, :222
        g = F.linear(_ln(h, sd, p + "ln_2"), sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])         # :209
        h = h + F.linear(_quick_gelu(g), sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])         # :211, :223
        a = dict(k=k, v=v)
        if with_q:
            a["q"] = q
        if with_out:
            a["out"] = h
        kvs.append(a)
    return kvs


def decoder_attention(qs, k, v, m, attn_mode=(), num_frames=None):
    """MultiheadAttention.forward of the decoder without its projections (src/models.py:138-144) with the two
    activations smax (:99-115; ``attn_mode`` = () default, or a subset of {"frame", "temporal"}) and coda (:117-125).
    qs: [B,1,H,2*dh] per head [smax query | coda query]; k, v: [B,S,H,dh]; m: bool [B,S]. Returns [B,1,H,dh]."""
    dh = k.shape[-1]
    q0, q1 = qs.split(dh, dim=-1)                       # :137 view(B,1,H,-1).split(dh,-1)
    mm = m.unsqueeze(1).unsqueeze(-1)                   # :138
    norm = dh ** 0.5
    smax = torch.einsum("nqhc,nkhc->nqkh", q0 / norm, k).masked_fill(~mm, float("-inf"))
    if len(attn_mode) == 0:
        smax = smax.softmax(dim=-2)                     # :105-106
    else:                                               # :107-115
        n, q, s, h = smax.shape
        aff = smax.view(n, q, num_frames, -1, h)
        parts = []
        if "frame" in attn_mode:
            parts.append(aff.softmax(dim=-2))
        if "temporal" in attn_mode:
            parts.append(aff.softmax(dim=-3))
        smax = sum(parts).view(n, q, s, h)
    coda_aff = torch.einsum("nqhc,nkhc->nqkh", q1 / norm, k).tanh()
    gate = -(q1 - k).abs().sum(-1).unsqueeze(1) / norm
    gate = 2 * gate.sigmoid().masked_fill(~mm, 0.0)
    aff = smax / 2 + (coda_aff * gate) / 2              # :140-142, n_act = 2
    return torch.einsum("nqlh,nlhc->nqhc", aff, v)      # :144


def decoder_forward(sd, kvs, m, out_dims, prefix="decoder.", layer_indices=None, attn_mode=()):
    """Decoder.forward (src/models.py:323-361) + Transformer.forward (:259-269) + ResidualAttentionBlock.forward
    (:173-176). kvs: list of {k, v: [B,T,P,H,dh]} (CLS already dropped); m: bool [B,T].
    op_mode is inferred from the state dict like the constructor creates parameters: ``augment_query_{i}`` present =
    aug_query (:250-255, 265-267); ``proj{i}x{o}_L{l}`` present = global_prediction (:309-313, 345-357; needs
    ``layer_indices``); no ``positional_embedding`` = temporal_position off.
    Returns (raw task logits, video feature [B,D] ([B,n_blocks,D] with global_prediction), block outputs)."""
    b, t, p, heads, dh = kvs[0]["k"].shape
    width = heads * dh
    mm = m.repeat_interleave(p, dim=-1)                                  # :324
    pe = sd.get(prefix + "positional_embedding")
    flat = []
    for kv in kvs:
        k, v = kv["k"].float(), kv["v"].float()
        if pe is not None:                                               # :326-329 (added to K and to V)
            k, v = k + pe, v + pe
        flat.append((k.flatten(1, 2), v.flatten(1, 2)))                  # :332-334
    x = sd[prefix + "class_embedding"].view(1, 1, -1).repeat(b, 1, 1)    # :336
    x = _ln(x, sd, prefix + "ln_pre")                                    # :337 (dropout p = 0)
    outs = []
    for i, (k, v) in enumerate(flat):
        bp = "%stransformer.resblocks.%d." % (prefix, i)
        y = _ln(x, sd, bp + "ln_1")
        qs = F.linear(y, sd[bp + "attn.in_proj.weight"], sd[bp + "attn.in_proj.bias"]).view(b, 1, heads, -1)  # :137
        mix = decoder_attention(qs, k, v, mm, attn_mode, t)
        x = x + F.linear(mix.flatten(-2), sd[bp + "attn.out_proj.weight"], sd[bp + "attn.out_proj.bias"])   # :146, :174
        g = F.linear(_ln(x, sd, bp + "ln_2"), sd[bp + "mlp.c_fc.weight"], sd[bp + "mlp.c_fc.bias"])
        x = x + F.linear(_quick_gelu(g), sd[bp + "mlp.c_proj.weight"], sd[bp + "mlp.c_proj.bias"])          # :175
        outs.append(x)
        aug = sd.get("%stransformer.augment_query_%d" % (prefix, i))
        if aug is not None and i != len(flat) - 1:                       # :265-267
            x = x + aug
    blocks = torch.cat(outs, dim=1)                                      # :269
    if any(key.startswith(prefix + "proj0x") and "_L" in key for key in sd):   # global_prediction
        feat = _ln(blocks, sd, prefix + "ln_post")                       # :342 on every block output
        n = len(layer_indices)
        logits = [sum((feat[:, j] @ sd["%sproj%dx%d_L%d" % (prefix, i, o, l)]) * (j + 1) / ((1 + n) * n / 2)
                      for j, l in enumerate(layer_indices)) for i, o in enumerate(out_dims)]   # :345-357
        return logits, feat, blocks
    feat = _ln(blocks[:, -1], sd, prefix + "ln_post")                    # :340-343
    logits = [feat @ sd["%sproj%dx%d" % (prefix, i, o)] for i, o in enumerate(out_dims)]  # :359
    return logits, feat, blocks


# adapter.struct.type -> Sequential indices of (first Linear, LayerNorm or None, middle Linear or None, last Linear)
_ADAPTER_LAYOUT = {
    "768-x-768": (0, 2, None, 4), "legacy-768-x-768": (0, 2, None, 3), "768-x-768-nln": (0, 1, None, 4),
    "768-x-768-ln": (0, 1, None, 4), "768-x-768-z0": (0, 1, None, 4), "768-xxx-768": (0, None, 3, 6),
    "linear": (0, None, None, None),
}


def adapter_forward(sd, kvs, struct_type, prefix="adapter."):
    """CompInvAdapter.forward (src/models.py:921-935) in eval mode (every Dropout is the identity) for the
    Sequential layouts of :797-917. kvs: list of {k, v: [B,T,P,H,dh]}; returns the adapted list (new tensors)."""
    b, t, p, h, d = kvs[0]["k"].shape
    if struct_type == "768-bn":
        # Linear(768, 768, bias=False) -> BatchNorm2d(num_frames) in eval mode (:877-887): dim 1 of [b, t, p, 768] is
        # the BatchNorm channel, i.e. one scalar affine per frame index from the running statistics (eps 1e-5)
        out = []
        for i, kv in enumerate(kvs):
            new = {}
            for name in ("k", "v"):
                pre = "%sl%d_%s." % (prefix, i, name)
                x = kv[name].float().reshape(b, t, p, h * d)
                y = F.linear(x, sd[pre + "0.weight"])
                scale = sd[pre + "1.weight"] / torch.sqrt(sd[pre + "1.running_var"] + 1e-5)
                shift = sd[pre + "1.bias"] - sd[pre + "1.running_mean"] * scale
                y = y * scale.view(1, t, 1, 1) + shift.view(1, t, 1, 1)
                new[name] = kv[name].float() + y.view(b, t, p, h, d)     # :930-931
            out.append(new)
        return out
    first, ln, mid, last = _ADAPTER_LAYOUT[struct_type]
    out = []
    for i, kv in enumerate(kvs):
        new = {}
        for name in ("k", "v"):
            pre = "%sl%d_%s." % (prefix, i, name)
            x = kv[name].float().reshape(b, t, p, h * d)                 # :926 view((b, t, p, -1))
            y = F.linear(x, sd[pre + "%d.weight" % first])
            if struct_type in ("768-x-768", "legacy-768-x-768"):          # Linear, GELU, LayerNorm(x), Linear
                y = F.layer_norm(F.gelu(y), (y.shape[-1],), sd[pre + "%d.weight" % ln], sd[pre + "%d.bias" % ln], 1e-5)
            elif struct_type in ("768-x-768-ln", "768-x-768-z0"):         # Linear, LayerNorm(x), GELU, Linear
                y = F.gelu(F.layer_norm(y, (y.shape[-1],), sd[pre + "%d.weight" % ln], sd[pre + "%d.bias" % ln], 1e-5))
            elif struct_type == "768-x-768-nln":                          # Linear, LayerNorm((P, x)), GELU, Linear
                y = F.gelu(F.layer_norm(y, tuple(y.shape[-2:]), sd[pre + "%d.weight" % ln], sd[pre + "%d.bias" % ln],
                                        1e-5))
            elif struct_type == "768-xxx-768":                            # Linear, GELU, Linear, GELU, Linear
                y = F.gelu(F.linear(F.gelu(y), sd[pre + "%d.weight" % mid]))
            if last is not None:
                y = F.linear(y, sd[pre + "%d.weight" % last])
            y = y.view(b, t, p, h, d)
            new[name] = kv[name].float() + y if struct_type != "linear" else y   # :930-933
        out.append(new)
    return out


def normalise_logits(task_logits):
    # Detector.predict, src/models.py:551-553
    return [5 * l / (torch.norm(l, dim=-1, keepdim=True) + 1e-10) for l in task_logits]


def ema_frames(x, m, ratio):
    """op_mode.ema_frame in Detector.forward (src/models.py:572-578)."""
    b, t, c, h, w = x.shape
    _x = torch.zeros((b, 1, c, h, w))
    for i in range(t):
        _x = _x * ratio + x[:, i].unsqueeze(1) * (1 - ratio)
    return _x, m[:, 0].unsqueeze(1)


def detector_predict(sd, x, m, layer_indices, out_dims, return_taps=False, adapter=None, attn_mode=(),
                     patch_indices=None):
    """Detector.predict (src/models.py:498-566; ``adapter`` = adapter.struct.type or None; ``patch_indices`` = the
    per-layer index arrays train_mode.patch_mask drew, :511-544).
    x: fp32 [B,T,3,R,R]; m: bool [B,T]. Returns (normalised task logits, video feature[, taps])."""
    b, t = x.shape[:2]
    run = max(layer_indices) + 1
    enc = encoder_forward(sd, x.flatten(0, 1), num_layers=run)                                   # :503
    kvs = [{n: enc[i][n][:, 1:].unflatten(0, (b, t)) for n in ("k", "v")} for i in layer_indices]  # :505-509
    if patch_indices is not None:
        kvs = [{n: kv[n][:, :, idx] for n in kv} for kv, idx in zip(kvs, patch_indices)]         # :543-544
    if adapter is not None:
        kvs = adapter_forward(sd, kvs, adapter)                                                  # :546-547
    logits, feat, _ = decoder_forward(sd, kvs, m, out_dims, layer_indices=layer_indices, attn_mode=attn_mode)  # :549
    logits = normalise_logits(logits)
    if return_taps:
        return logits, feat, kvs
    return logits, feat


def detector_eval_losses(task_logits, labels):
    """auc_roc driver = per-sample cross entropy (src/models.py:34-45) as used by Detector.forward (:590-593)."""
    return [F.cross_entropy(l, y, reduction="none") for l, y in zip(task_logits, labels)]


def video_scores(clip_logits, clips_per_video):
    """inference.py:121,140: softmax per clip, then mean over the clips of each video."""
    probs = clip_logits.softmax(dim=-1)
    out, start = [], 0
    for n in clips_per_video:
        out.append(probs[start:start + n].mean(0))
        start += n
    return torch.stack(out)

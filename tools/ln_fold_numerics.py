#!/usr/bin/env python
"""Precision of a LayerNorm folded into the following projection, emulated on the CPU in float64 with explicit bf16
roundings, against the separate fp32 LayerNorm followed by the same bf16 GEMM — both against the exact result:

  unfused   bf16(LN(x)) @ bf16(W)^T + b
  folded    rstd * (bf16(x) @ bf16(gamma . W)^T - mu * colsum(bf16(gamma . W))) + (b + W beta)   (csrc/gemm_sm100.cu LNFOLD)

x = the oracle's residual stream entering encoder layers 1 and 2 of the synthetic models (real inputs of ln_1), W = that
layer's in_proj. The folded form rounds x itself to bf16, so it loses precision only where |mean| >> std of a row
(printed). Output recorded in profiles/r2_lnfuse_ab.md.   python tools/ln_fold_numerics.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import dfd_oracle as oracle  # noqa: E402
from dfdclip_b200 import synthetic  # noqa: E402


def bf(t):
    return t.float().bfloat16().double()


for arch, frames, clips in (("tiny-256x4", 4, 3), ("small-512x6", 3, 2), ("ViT-B/16", 8, 1)):
    d = synthetic.vit_dims(arch)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), seed=0)
    x, m = synthetic.make_clips(clips, frames, d["image_size"], seed=7)
    with torch.no_grad():
        enc = oracle.encoder_forward(sd, x.flatten(0, 1), with_out=True, with_q=True)
    for layer in (1, 2):
        xin = enc[layer - 1]["out"].reshape(-1, d["width"]).double()
        pre = "encoder.transformer.resblocks.%d." % layer
        g, b = sd[pre + "ln_1.weight"].double(), sd[pre + "ln_1.bias"].double()
        w, bias = sd[pre + "attn.in_proj_weight"].double(), sd[pre + "attn.in_proj_bias"].double()
        mu = xin.mean(-1, keepdim=True)
        rstd = (xin.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
        ref = ((xin - mu) * rstd * g + b) @ w.t() + bias
        unfused = bf((xin - mu) * rstd * g + b) @ bf(w).t() + bias
        wf = bf(w * g)
        folded = rstd * (bf(xin) @ wf.t() - mu * wf.sum(-1)) + (bias + w @ b)
        rel = lambda y: ((y - ref).norm() / ref.norm()).item()  # noqa: E731
        print("%-12s layer %d  |mean|/std of the rows %.3f   relative error of [q|k|v]: unfused %.5f   folded %.5f" % (
            arch, layer, (mu.abs() * rstd).mean().item(), rel(unfused), rel(folded)))

#!/usr/bin/env python
"""Run one unit kernel of libdfdclip_b200 at the C2 shapes a few times (target for `ncu -k regex:...`) and print
its CUDA-event time. Usage: python tools/prof_kernel.py mha|gemm_qkv|gemm_fc|gemm_proj|gemm_out|ln [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dfdclip_b200 import _native as nat  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "mha"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
F, L, H, D = 512, 197, 12, 768
M = F * L
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, generator=g) * scale).to(dev, dtype)


if which == "mha_l":  # ViT-L/14 shapes (config C4): 512 frames x 16 heads, 257 tokens
    F, L, H, D = 512, 257, 16, 1024
    M = F * L
    qkv = rnd(M, 3 * D, scale=1.5)
    fn = lambda: nat.mha_fwd(qkv, F, L, H)
    flops = 4 * H * L * L * 64 * F
    nbytes = M * 4 * D * 2
elif which == "mha":
    qkv = rnd(M, 3 * D, scale=1.5)
    fn = lambda: nat.mha_fwd(qkv, F, L, H)
    flops = 4 * H * L * L * 64 * F
    nbytes = M * 4 * D * 2
elif which.startswith("gemm"):
    n, k, epi = {"gemm_qkv": (3 * D, D, 0), "gemm_fc": (4 * D, D, 1), "gemm_proj": (D, 4 * D, 3),
                 "gemm_out": (D, D, 3)}[which]
    a, w, bias = rnd(M, k), rnd(n, k, scale=0.03), rnd(n, dtype=torch.float32)
    out = torch.zeros(M, n, device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
    fn = lambda: nat.gemm_bf16(a, w, bias, out, epi)
    flops = 2 * M * n * k
    nbytes = M * k * 2 + M * n * (8 if epi == 3 else 2)
elif which == "dec_attn":
    B, T, P = 64, 8, 196
    qkv = rnd(B * T * L, 3 * D)
    view = qkv.view(B, T, L, 3, H, 64)
    k, v = view[:, :, 1:, 1], view[:, :, 1:, 2]
    qs = rnd(B, H, 128, dtype=torch.float32)
    pe = rnd(T, H, 64, scale=0.05, dtype=torch.float32)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)
    fn = lambda: nat.decoder_attention(qs, k, v, pe, mask)
    flops = 0
    nbytes = 2 * B * T * P * D * 2
elif which == "ln":
    x = rnd(M, D, dtype=torch.float32)
    gam, bet = rnd(D, dtype=torch.float32), rnd(D, dtype=torch.float32)
    fn = lambda: nat.layernorm(x, gam, bet)
    flops = 0
    nbytes = M * D * 6
else:
    raise SystemExit("unknown kernel " + which)

for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("%s: %.1f us  %.1f TFLOP/s  %.0f GB/s (algorithmic)" % (which, ms * 1e3, flops / ms / 1e9, nbytes / ms / 1e6))

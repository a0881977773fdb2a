#!/usr/bin/env python
"""Library bar for the encoder GEMM shapes: our tcgen05 kernel vs torch.matmul (cuBLAS) on the same operands."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dfdclip_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
M, D = 512 * 197, 768
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, generator=g) * scale).to(dev, dtype)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, n, k, epi in [("qkv", 3 * D, D, 0), ("fc", 4 * D, D, 1), ("proj", D, 4 * D, 3), ("out", D, D, 3)] + ([("proj_st32", D, 4 * D, 2), ("proj_bf16", D, 4 * D, 0), ("out_st32", D, D, 2), ("out_bf16", D, D, 0), ("fc_nogelu", 4 * D, D, 0)] if len(sys.argv) > 1 else []):
    a, w, bias = rnd(M, k), rnd(n, k, scale=0.03), rnd(n, dtype=torch.float32)
    out = torch.zeros(M, n, device=dev, dtype=torch.float32 if epi >= 2 else torch.bfloat16)
    ref = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    wt = w.t()
    t_ours = timeit(lambda: nat.gemm_bf16(a, w, bias, out, epi))
    t_lib = timeit(lambda: torch.matmul(a, wt, out=ref))
    fl = 2 * M * n * k
    print("%-5s M=%d N=%d K=%d  ours %.1f us (%.0f TF/s)   cublas(no epilogue) %.1f us (%.0f TF/s)" % (
        name, M, n, k, t_ours * 1e3, fl / t_ours / 1e9, t_lib * 1e3, fl / t_lib / 1e9))

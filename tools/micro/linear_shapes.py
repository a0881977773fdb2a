"""The four decoder linear shapes at B=64 through dfd_linear_f32 (for ncu / timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dfdclip_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = [(1536, 768, False), (768, 768, False), (3072, 768, True), (768, 3072, False)]
g = torch.Generator().manual_seed(0)
ops = []
for n, k, gelu in shapes:
    ops.append((torch.randn((b, k), generator=g).to(dev), (torch.randn((n, k), generator=g) * k ** -0.5).to(dev),
                torch.randn((n,), generator=g).to(dev), gelu))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    for x, w, bias, gelu in ops:
        flush.zero_()
        nat.linear_f32(x, w, bias, None, quick_gelu=gelu)
torch.cuda.synchronize()
# event timing, weights cold (L2 flushed) and warm
for x, w, bias, gelu in ops:
    for cold in (True, False):
        ts = []
        for _ in range(10):
            if cold:
                flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            out = torch.empty((b, w.shape[0]), device=dev)
            lib = nat.load_library()
            nbytes = lib.dfd_linear_f32_workspace_bytes(b, w.shape[0])
            ws = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            a.record()
            nat.check(lib.dfd_linear_f32(nat.ctx(dev), nat.ptr(x), nat.ptr(w), nat.ptr(bias), None, nat.ptr(out), b,
                                         w.shape[0], w.shape[1], 1 if gelu else 0, nat.ptr(ws), nbytes,
                                         nat.stream_ptr(dev)))
            c.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(c) * 1e3)
        ts.sort()
        print("N=%d K=%d %s: median %.1f us (memset + kernel)" % (w.shape[0], w.shape[1], "cold" if cold else "warm",
                                                                    ts[len(ts) // 2]))

#!/usr/bin/env python
"""Per-phase clock64 timeline of the ViT-B/16 attention kernel (CTA 0, first 64 work items). Needs a library built with
DFD_MHA_TRACE=1 (python dfd-clip_b200/build.py --force with that variable set writes the clocks in the kernel).
Slots per item: MMA thread 0 p_full[A] seen, 1 next load seen, 2 o_empty[A] seen, 3 p_full[B] seen, 4 o_empty[B] seen;
tile A warp 5 s_full seen, 6 MUFU turn granted, 7 P written, 8 o_full seen, 9 epilogue issued; tile B 10..14 alike."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dfdclip_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
F, L, H, D = 512, 197, 12, 768
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(F * L, 3 * D, generator=g) * 1.5).to(dev, torch.bfloat16)
for _ in range(3):
    nat.mha_fwd(qkv, F, L, H)
torch.cuda.synchronize()
lib = nat.load_library()
buf = (ctypes.c_longlong * (64 * 16))()
rc = (lib.dfd_debug_mha_trace4 if os.environ.get('DFD_MHA_VARIANT') == '4' else lib.dfd_debug_mha_trace)(buf)
assert rc == 0, rc
rows = [[buf[i * 16 + s] for s in range(16)] for i in range(64)]
items = [r for r in rows[4:40] if r[5] > 0]
names = {"A: s_full -> turn (max pass + wait)": (5, 6), "A: turn -> P written (exp pass)": (6, 7),
         "A: P written -> o_full (PV MMA)": (7, 8), "A: o_full -> epilogue issued": (8, 9),
         "B: s_full -> turn": (10, 11), "B: turn -> P written (exp pass)": (11, 12), "B: P -> o_full": (12, 13),
         "B: o_full -> epilogue issued": (13, 14)}
for name, (a, b) in names.items():
    d = [r[b] - r[a] for r in items]
    print("%-40s mean %7.0f  min %6d  max %6d clk" % (name, sum(d) / len(d), min(d), max(d)))
extra = {"P written (A, slot 7) -> MMA thread sees p_full[A] (slot 0)": (7, 0),
         "MMA thread sees p_full[A] (0) -> A sees o_full (8)": (0, 8),
         "A epilogue issued (9) -> MMA thread sees o_empty[A] (2)": (9, 2)}
for name, (a, b) in extra.items():
    d = [r[b] - r[a] for r in items]
    print("%-60s mean %7.0f  min %6d  max %6d clk" % (name, sum(d) / len(d), min(d), max(d)))
period = [items[i + 1][5] - items[i][5] for i in range(len(items) - 1)]
print("period per item (A s_full to next A s_full): mean %.0f clk" % (sum(period) / len(period)))
ab = [r[10] - r[5] for r in items]
print("B s_full - A s_full: mean %.0f" % (sum(ab) / len(ab)))
nexts = [items[i + 1][5] - items[i][9] for i in range(len(items) - 1)]
print("A: epilogue issued -> next s_full: mean %.0f" % (sum(nexts) / len(nexts)))
nextsb = [items[i + 1][10] - items[i][14] for i in range(len(items) - 1)]
print("B: epilogue issued -> next s_full: mean %.0f" % (sum(nextsb) / len(nextsb)))

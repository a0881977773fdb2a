"""Kernel-time table of the C5 training step (eager), from torch.profiler: where the decoder's forward/backward and
the optimizer spend GPU time next to the native encoder."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_detector  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
det, _ = build_detector("ViT-B/16", 8, dev)
det.train()
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 12
g = torch.Generator().manual_seed(5)
x = torch.randn((clips, 8, 3, 224, 224), generator=g).to(dev)
m = torch.ones((clips, 8), dtype=torch.bool, device=dev)
y = torch.randint(0, 2, (clips,), generator=g).to(dev)
opt = det.configure_optimizers(lr=1e-3)


def step():
    with torch.enable_grad():
        losses, _, _ = det(x, [y], m, train=True, single_task=0)
        losses[0].mean().backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
total = sum(r.device_time_total for r in rows)
print("total device time per step: %.1f us, %d kernels per step" % (total / n, sum(r.count for r in rows) / n))
for r in rows[:28]:
    print("%8.1f us/step  %5.1f%%  x%-4d %s" % (r.device_time_total / n, 100 * r.device_time_total / total, r.count / n,
                                               r.key[:110]))

#!/usr/bin/env python
"""The "library bar" of SURVEY 8(d): the reference's algorithm (the oracle port of Detector.predict: plain torch ops,
cuBLAS / cuDNN / ATen kernels) run ON THE B200 in fp32 (TF32 off), in fp32 with TF32, and under bf16 autocast, at
config C2 — what a user gets from the reference code on this GPU, next to the hand-written path. Not a bench arm and
not part of the product: a one-off measurement recorded in profiles/ and DESIGN.md."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import dfd_oracle  # noqa: E402
from dfdclip_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
arch, frames, clips = "ViT-B/16", 8, int(sys.argv[1]) if len(sys.argv) > 1 else 64
taps = synthetic.layer_indices(arch)
sd = {k: v.to(dev) for k, v in synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=taps, seed=0).items()}
x, m = synthetic.make_clips(clips, frames, 224, seed=7, masked_tail=False)
x, m = x.to(dev), m.to(dev)


def run(label, steps=5, autocast=False, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        for _ in range(2):
            dfd_oracle.detector_predict(sd, x, m, taps, (2,))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            dfd_oracle.detector_predict(sd, x, m, taps, (2,))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": label, "clips_per_s": clips / (ms * 1e-3), "ms_per_step": ms}


out = [run("torch fp32 (TF32 off)"), run("torch fp32 + TF32", tf32=True), run("torch bf16 autocast", autocast=True)]
print(json.dumps({"workload": "C2: %d clips x %d frames, %s, oracle port of Detector.predict on cuda:0" % (clips, frames, arch),
                  "results": out}))

"""BASELINE config C5 on N GPUs: frozen-encoder training step under DistributedDataParallel as the reference trains
(accelerate's DDP with find_unused_parameters=True, main.py:283): 12 clips x 8 frames per GPU, Detector.forward(train=True),
backward with the bucketed NCCL all-reduce of the 39 M decoder gradients, SGD step. Eager (the captured
GraphedTrainStep does not include the all-reduce). Rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5_ddp.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_detector  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    saved = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    clips, frames, steps = 12, 8, int(os.environ.get("C5_STEPS", "20"))
    det, _ = build_detector("ViT-B/16", frames, dev)
    det.train()
    model = DDP(det, device_ids=[local], find_unused_parameters=True)
    opt = det.configure_optimizers(lr=1e-3)
    g = torch.Generator().manual_seed(5 + rank)
    res = det.encoder.input_resolution
    x = torch.randn((clips, frames, 3, res, res), generator=g).to(dev)
    m = torch.ones((clips, frames), dtype=torch.bool, device=dev)
    y = torch.randint(0, 2, (clips,), generator=g).to(dev)

    def step():
        with torch.enable_grad():
            losses, _, _ = model(x, [y], m, train=True, single_task=0)
            losses[0].mean().backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(5):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # replicas must stay identical: compare a parameter checksum across ranks
    chk = torch.stack([p.detach().double().sum() for p in det.decoder.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item()
        line = {"config": "C5: frozen-encoder training step under DDP, %d clips x %d frames per GPU, ViT-B/16, SGD" % (
            clips, frames), "n_gpus": world, "ms_per_step": ms, "clips_per_s": world * clips / (ms * 1e-3),
            "replicas_identical": bool((lo == hi).item()), "steps": steps,
            "api": "torch DDP(find_unused_parameters=True) around dfdclip_b200.models.Detector, eager step"}
        os.write(saved, (json.dumps(line) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

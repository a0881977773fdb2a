"""BASELINE config C3 as stated: the FF++ cross-manipulation test shape of SURVEY 8(d) — 560 synthetic videos with
U{8..32} one-second clips each (seed 3), video-level averaging — scored data-parallel under torchrun: videos sharded
over the ranks by clip count (shard_videos), every rank scores its shard through score_videos_batched (packer +
HostClipStream, uint8 clips), ONE all_gather of per-video scores. Each rank materialises only the videos of its own
shard (the others are zero-stride placeholders that only carry their shape). Rank 0 prints one JSON line.
C3_PIN=1 page-locks each rank's videos first (a DataLoader with pin_memory=True): no staging pass in the driver.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c3_multi.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_detector  # noqa: E402
from dfdclip_b200.inference import (HostClipStream, pack_clip_batches, score_videos_batched,  # noqa: E402
                                    shard_videos)


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL's banner goes to stderr
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_videos, frames = int(os.environ.get("C3_VIDEOS", "560")), 8
    pin_sources = os.environ.get("C3_PIN", "0") == "1"
    det, _ = build_detector("ViT-B/16", frames, dev)
    res = det.encoder.input_resolution
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(8, 33, (n_videos,), generator=g).tolist()
    mine = set(shard_videos(counts, world)[rank])
    gr = torch.Generator().manual_seed(100 + rank)
    pool = torch.randint(0, 256, (64, frames, 3, res, res), generator=gr, dtype=torch.uint8)  # pixels to draw from
    videos, masks = [], []
    for i, n in enumerate(counts):
        if i in mine:
            idx = torch.randint(0, 64, (n,), generator=gr)
            v = pool[idx].clone()
            videos.append(v.pin_memory() if pin_sources else v)  # C3_PIN=1: what a pinning DataLoader delivers
        else:
            videos.append(torch.zeros((), dtype=torch.uint8).expand(n, frames, 3, res, res))
        mk = torch.ones((n, frames), dtype=torch.bool)
        masks.append(mk.pin_memory() if (pin_sources and i in mine) else mk)
    with torch.no_grad():
        warm = sorted(mine)[:4]  # warm-up without collectives: this rank's first videos through the same pipeline
        list(HostClipStream(det).run(pack_clip_batches([videos[i] for i in warm], [masks[i] for i in warm], 64)))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scores = score_videos_batched(det, videos, masks, batch_clips=64)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
        chk = scores.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).item())
    else:
        same = True
    if rank == 0:
        line = {"config": "C3: %d synthetic videos, %d clips (U{8..32} per video, seed 3), ViT-B/16, 8 frames, uint8 clips, "
                          "video-level mean of clip probabilities" % (n_videos, sum(counts)),
                "n_gpus": world, "page_locked_videos": pin_sources, "seconds": dt, "clips_per_s": sum(counts) / dt, "videos_per_s": n_videos / dt,
                "clips_on_rank0": sum(counts[i] for i in mine), "scores_finite": bool(torch.isfinite(scores).all().item()),
                "scores_identical_on_all_ranks": same,
                "api": "dfdclip_b200.inference.score_videos_batched (shard_videos + pack_clip_batches + HostClipStream, "
                       "one all_gather)"}
        os.write(saved, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Measurements of BASELINE.json's other configurations on one B200 (they are parity-test cases, not bench.py lines;
these numbers go to profiles/ next to the headline line):

  c3  video-level scoring (inference.py:107-156): one rank's share (1/8) of the FF++ test shape of SURVEY 8(d) —
      70 synthetic videos with U{8..32} one-second clips each — through score_videos_batched (packer + HostClipStream)
      and through the reference-style loop (one video at a time, chunks of 16 clips, blocking copies).
  c5  frozen-encoder training step (src/trainer.py:147 -> Detector.forward(train=True), backward, SGD step),
      12 clips x 8 frames per GPU (configs/deepfake/deepfake.yaml:97).

Prints one JSON line per configuration. Inputs are synthetic; weights are the seeded random weights of bench.py.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_detector  # noqa: E402


def run_c3(args, dev):
    from dfdclip_b200.inference import score_videos, score_videos_batched
    det, _ = build_detector(args.arch, args.frames, dev)
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(8, 33, (args.videos,), generator=g).tolist()
    res = det.encoder.input_resolution
    out = {"config": "C3 shard: %d videos, %d clips (U{8..32} per video), %s, %d frames" % (
        args.videos, sum(counts), args.arch, args.frames), "results": []}
    for dtype in (torch.uint8, torch.float32):
        if dtype == torch.uint8:
            videos = [torch.randint(0, 256, (n, args.frames, 3, res, res), generator=g, dtype=torch.uint8) for n in counts]
        else:
            videos = [torch.randn((n, args.frames, 3, res, res), generator=g) for n in counts]
        masks = [torch.ones((n, args.frames), dtype=torch.bool) for n in counts]
        with torch.no_grad():
            score_videos_batched(det, videos[:4], masks[:4], batch_clips=args.batch)  # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = score_videos_batched(det, videos, masks, batch_clips=args.batch)
            torch.cuda.synchronize()
            dt_b = time.perf_counter() - t0
            predict = lambda x, m: det.predict(x, m)[0][0]  # noqa: E731
            score_videos(predict, videos[:2], masks[:2], chunk_clips=16, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = score_videos(predict, videos, masks, chunk_clips=16, device=dev)
            torch.cuda.synchronize()
            dt_r = time.perf_counter() - t0
            # the same videos already page-locked (a DataLoader with pin_memory=True): no staging pass
            pv, pm = [v.pin_memory() for v in videos], [mm.pin_memory() for mm in masks]
            score_videos_batched(det, pv[:4], pm[:4], batch_clips=args.batch)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got_p = score_videos_batched(det, pv, pm, batch_clips=args.batch)
            torch.cuda.synchronize()
            dt_p = time.perf_counter() - t0
            del pv, pm
        out["results"].append({
            "pinned_sources_clips_per_s": sum(counts) / dt_p, "pinned_sources_s": dt_p,
            "pinned_equals_staged": bool(torch.equal(got_p, got)),
            "input": str(dtype).replace("torch.", ""),
            "batched_stream_clips_per_s": sum(counts) / dt_b, "batched_stream_s": dt_b,
            "per_video_chunk16_clips_per_s": sum(counts) / dt_r, "per_video_chunk16_s": dt_r,
            "max_abs_score_diff": (got - ref).abs().max().item()})
        del videos
    print(json.dumps(out))


def run_c5(args, dev):
    det, _ = build_detector(args.arch, args.frames, dev)
    det.train()
    clips = args.train_clips
    res = det.encoder.input_resolution
    g = torch.Generator().manual_seed(5)
    x = torch.randn((clips, args.frames, 3, res, res), generator=g).to(dev)
    m = torch.ones((clips, args.frames), dtype=torch.bool, device=dev)
    y = torch.randint(0, 2, (clips,), generator=g).to(dev)
    opt = det.configure_optimizers(lr=1e-3)

    def step():
        with torch.enable_grad():
            losses, _, _ = det(x, [y], m, train=True, single_task=0)
            loss = losses[0].mean()
            loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    ms = timed(step, args.steps)
    from dfdclip_b200.training import GraphedTrainStep
    graphed = GraphedTrainStep(det, opt, x, y, m)
    ms_graph = timed(lambda: graphed(x, y, m), args.steps)

    @torch.no_grad()
    def enc_only():
        det.encoder.encode(x.flatten(0, 1), keep_layers=det.layer_indices)

    ms_enc = timed(enc_only, args.steps)

    @torch.no_grad()
    def infer():
        det.predict(x, m)

    det.eval()
    ms_inf = timed(infer, args.steps)
    print(json.dumps({"config": "C5: frozen-encoder training step, %d clips x %d frames, %s, SGD" % (
        clips, args.frames, args.arch), "clips_per_s": clips / (ms_graph * 1e-3), "ms_per_step": ms_graph,
        "api": "dfdclip_b200.training.GraphedTrainStep (forward + backward + SGD step replayed from one CUDA graph)",
        "eager_clips_per_s": clips / (ms * 1e-3), "eager_ms_per_step": ms,
        "ms_encoder_forward_only": ms_enc, "ms_inference_predict_same_batch": ms_inf,
        "trainable_params": sum(p.numel() for p in det.parameters() if p.requires_grad)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c3", "c5", "both"])
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--videos", type=int, default=70)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--train-clips", type=int, default=12)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    if args.which in ("c5", "both"):
        run_c5(args, dev)
    if args.which in ("c3", "both"):
        run_c3(args, dev)


if __name__ == "__main__":
    main()

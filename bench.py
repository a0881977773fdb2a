#!/usr/bin/env python
"""Benchmark of the DFD-CLIP hot path on B200: clips/sec of Detector.predict (CLIP ViT frame encoder with K/V
taps + temporal decoder/head) on synthetic clips, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5]

One "step" = one pass of the hot path over one batch of synthetic clips per GPU (default workload C2: 64 clips x 8
frames x 224^2, ViT-B/16, taps [0,2,..,10], bf16 tensor-core math / fp32 accumulate). Under torchrun every rank runs
the same per-GPU batch (weak scaling, clips are independent units); the per-clip scores of all K steps are
all-gathered ONCE, after the last step and inside the timed region (the "final logit/score gather" of
inference.py:147 — never inside the clip loop). Prints ONE JSON line on rank 0.

Other BASELINE.json configurations: --workload c4 (ViT-L/14, 32 clips x 16 frames), c3 (the FF++ video-level job:
560 synthetic videos of U{8..32} clips, sharded over the ranks by clip count, per-video mean of clip probabilities,
one all_gather; strong scaling), c5 (frozen-encoder training step, 12 clips x 8 frames per GPU: native decoder
forward/backward, in-step gradient all-reduce, SGD).

Keys beyond the base contract: `value` is device-resident (inputs in HBM when the timed region starts); `e2e` is the
same metric through the public caller loop `dfdclip_b200.inference.predict_stream` with pinned HOST clips in and
host logits out (H2D / encoder / decoder + D2H of consecutive batches on three streams; every batch's copies inside
the timed region); `e2e_u8` the same on raw uint8 pixels; `e2e_single_call` one blocking host call per batch;
`roofline` for the dominant kernel family (tcgen05 GEMMs; per-launch CUDA events in a separate pass, peak from
MEASURED_PEAKS.json) plus `hbm_kernels`; `cpu_baseline` = the oracle port on the host cores (N=1 only);
`--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "clips/s"

# BASELINE.json configs reachable from this script (--workload); c2 is the configuration the metric is quoted on
WORKLOADS = {
    "c2": dict(arch="ViT-B/16", clips=64, frames=8, label="C2"),
    "c3": dict(arch="ViT-B/16", clips=64, frames=8, label="C3"),
    "c4": dict(arch="ViT-L/14", clips=32, frames=16, label="C4"),
    "c5": dict(arch="ViT-B/16", clips=12, frames=8, label="C5"),
}


def metric_name(args, res=224):
    what = {"c3": "enc+head, video-level job", "c5": "frozen enc + decoder fwd/bwd + SGD step"}.get(args.workload, "enc+head")
    return "clips/sec (%dx%d^2 frames, %s %s)" % (args.frames, res, args.arch, what)


# ----------------------------------------------------------------------------------------------- helpers
def flops_per_clip(dims, frames, taps, executed=True):
    """Algorithmic FLOPs (2*m*n*k) per clip, SURVEY 8(d). executed=True counts only what this implementation runs
    (layers after the last tap are skipped, the last tapped layer stops after its QKV projection)."""
    r, p, d, h, layers = dims["image_size"], dims["patch_size"], dims["width"], dims["heads"], dims["layers"]
    pp = (r // p) ** 2
    seq = pp + 1
    patch = 2 * pp * (3 * p * p) * d
    qkv = 2 * seq * d * 3 * d
    attn = 4 * h * seq * seq * 64
    out = 2 * seq * d * d
    mlp = 16 * seq * d * d
    full = qkv + attn + out + mlp
    if executed:
        last = max(taps)
        kv_only = qkv * 2 // 3  # the last tapped layer computes K and V only
        per_frame = patch + last * full + kv_only
        gemm = patch + last * (qkv + out + mlp) + kv_only
    else:
        per_frame = patch + layers * full
        gemm = patch + layers * (qkv + out + mlp)
    return per_frame * frames, gemm * frames


def launches_per_predict(layers_run_full, n_taps, n_tasks, adapter=False):
    """Kernels of this library per Detector.predict. Encoder: patchify, patch GEMM, ln_pre; per full layer QKV GEMM,
    attention, out-proj GEMM, ln_2, c_fc GEMM, c_proj GEMM (+ ln_1 unless it is folded into the QKV GEMM, the default;
    ln_2 is folded too with DFD_LN_FUSE=1); the last tapped layer only its K/V projection (+ ln_1). Decoder: ln_pre +
    broadcast, per block 2 LayerNorms, 4 linears of 2 kernels, attention stream + combine, block-output scatter; ln_post;
    one projection kernel per task."""
    mode = os.environ.get("DFD_LN_FUSE", "2")
    ln_per_layer = {"0": 2, "1": 0}.get(mode, 1)
    enc = 3 + (5 + ln_per_layer) * layers_run_full + 1 + (1 if mode == "0" else 0)
    dec = 2 + 13 * n_taps + 1 + n_tasks
    return enc + dec + (6 * n_taps if adapter else 0)  # adapter: down GEMM, norm/act kernel, up GEMM per k and v


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), tflops_burst=p.get("bf16_tflops"),
                    hbm_gbs=p.get("hbm_gbs"), source="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (NVML, nvidia-smi as a fallback)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _reason_names(self, mask):
        nv = self.nv
        table = [("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"),
                 ("sync_boost", "nvmlClocksThrottleReasonSyncBoost"),
                 ("app_clocks", "nvmlClocksThrottleReasonApplicationsClocksSetting")]
        out = []
        for name, attr in table:
            bit = getattr(nv, attr, None)
            if bit is not None and mask & bit:
                out.append(name)
        return out

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nv is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.reasons.update(self._reason_names(mask))
                else:
                    import subprocess
                    out = subprocess.run(
                        ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [t.strip() for t in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         f[2:]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def build_detector(arch, frames, device, adapter=None, taps=None):
    from dfdclip_b200 import synthetic
    from dfdclip_b200.config import CN
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    if taps:
        cfg.decode_mode = "index"
        cfg.decode_indices = list(taps)
    if adapter:
        cfg.adapter.type = "normal"
        cfg.adapter.frozen = 0
        cfg.adapter.struct = CN({"type": adapter, "x": 256})
    det = Detector(cfg, frames, None)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det.layer_indices, seed=0, adapter=adapter)
    det.load_state_dict(sd, strict=True)
    return det.to(device).eval(), sd


# ------------------------------------------------------------------------------------------------ CPU arm
class CpuArm:
    """The reference's implementation of the path on the host cores: the UNMODIFIED reference staged under
    baseline/_ref (oracle/stage_reference.py; kind "reference") when present, else the oracle port (kind "port").
    Test/measurement infrastructure: this is the one place bench.py touches oracle/."""

    def __init__(self, arch, frames, threads=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from dfdclip_b200 import synthetic
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.arch, self.frames = arch, frames
        self.dims = synthetic.vit_dims(arch)
        self.taps = synthetic.layer_indices(arch)
        self.det = None
        root = None
        try:
            import reference_runner
            root = reference_runner.reference_root()
            if root is not None:
                self.det = reference_runner.build_reference_detector(arch, frames, root=root)
        except Exception as exc:  # the port always exists
            print("reference not usable here (%s): timing the oracle port" % exc, file=sys.stderr)
            self.det = None
        if self.det is None:
            import dfd_oracle
            self.oracle = dfd_oracle
            self.sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=self.taps, seed=0)
        self.kind = "reference" if self.det is not None else "port"
        self.what = ("the unmodified reference's Detector (src/models.py imported from %s), torch fp32" % (
            os.path.relpath(root, ROOT) if root.startswith(ROOT) else root)
                     if self.det is not None else "Detector.predict restated in torch fp32 (oracle port)")

    def clips(self, n, seed=7):
        from dfdclip_b200 import synthetic
        return synthetic.make_clips(n, self.frames, self.dims["image_size"], seed=seed, masked_tail=False)

    def predict(self, x, m):
        with torch.no_grad():
            if self.det is not None:
                return self.det.predict(x, m)[0][0]
            return self.oracle.detector_predict(self.sd, x, m, self.taps, (2,))[0][0]

    def video_score(self, x, m, chunk=16):
        """One video the way inference.py:113-141 scores it: chunks of 16 clips, softmax, mean."""
        logits = torch.cat([self.predict(x[i:i + chunk], m[i:i + chunk]) for i in range(0, x.shape[0], chunk)])
        return logits.softmax(-1).mean(0)

    def train_step_fn(self, clips):
        """The reference trainer's step (src/trainer.py:147-178) on `clips` clips; needs the reference itself."""
        if self.det is None:
            return None
        det = self.det.train()
        opt = det.configure_optimizers(1e-3)
        x, m = self.clips(clips)
        y = torch.arange(clips) % 2

        def step():
            opt.zero_grad()
            losses, _, other = det(x, [y], m, comp=["raw"] * clips, speed=torch.ones(clips), train=True, single_task=0)
            (losses[0].mean() + sum(other.values())).backward()
            opt.step()
        return step


def time_cpu(fn, repeats):
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(args):
    """cpu_baseline leg of the b200 arm (rank 0, N=1): a bounded sample of the workload on the host cores."""
    arm = CpuArm(args.arch, args.frames)
    n = args.ref_clips
    if args.workload == "c5":
        step = arm.train_step_fn(n)
        if step is not None:
            times = time_cpu(step, 3)
            return {"value": n / min(times), "unit": UNIT, "cores": arm.threads, "kind": arm.kind,
                    "sample": "%d clips x %d frames, best of %d training steps (forward train + backward + SGD) of %s"
                              % (n, args.frames, len(times), arm.what)}
    x, m = arm.clips(n)
    times = time_cpu(lambda: arm.predict(x, m), 3)
    return {"value": n / min(times), "unit": UNIT, "cores": arm.threads, "kind": arm.kind,
            "sample": "%d clips x %d frames in one predict call, best of %d runs of %s" % (
                n, args.frames, len(times), arm.what)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the workload on this box's host cores (all of
    them), each step a bounded sample of the workload. Rank 0 alone runs; the other ranks exit without work."""
    if rank != 0:
        return
    arm = CpuArm(args.arch, args.frames)
    n = args.ref_clips
    if args.workload == "c5" and arm.train_step_fn(1) is not None:
        fn = arm.train_step_fn(n)
        sample = "%d clips x %d frames per step: forward(train) + backward + SGD step of %s" % (n, args.frames, arm.what)
    elif args.workload == "c3":
        x, m = arm.clips(n)
        fn = lambda: arm.video_score(x, m)   # noqa: E731
        sample = "one synthetic video of %d clips x %d frames per step, chunks of 16, softmax + mean (%s)" % (
            n, args.frames, arm.what)
    else:
        x, m = arm.clips(n)
        fn = lambda: arm.predict(x, m)       # noqa: E731
        sample = "%d synthetic clips x %d frames per step in one predict call (%s)" % (n, args.frames, arm.what)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    wl = WORKLOADS[args.workload]["label"]
    line = {
        "impl": "reference", "metric": metric_name(args, arm.dims["image_size"]), "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.workload == "c3" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s sample on the host CPU: %s" % (wl, sample), "arch": args.arch, "frames": args.frames,
                   "clips_per_step": n, "per_clip_rate": "clips/s is a per-clip rate: comparable with the GPU arm's "
                   "although the GPU arm's step holds more clips"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.threads, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------- b200 helpers
def init_device(world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    return dev, dist


def max_over_ranks(value, dev, dist):
    if dist is None:
        return value
    t = torch.tensor([value], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timed_kernels(dev, fn, passes):
    """Per-kernel durations: `fn` run `passes` times with one CUDA event pair around every launch (recorded by the
    library on the launching stream; 2-3 % overhead, so kept out of the headline region). {tag: (ms per pass, launches
    per pass)}."""
    from dfdclip_b200 import _native
    _native.timing_enable(dev, True)
    torch.cuda.synchronize()
    for _ in range(passes):
        fn()
    torch.cuda.synchronize()
    kernel_ms = _native.timing_read(dev)
    _native.timing_enable(dev, False)
    return {k: (v[0] / passes, v[1] / passes) for k, v in kernel_ms.items()}


def load_traffic():
    """DRAM bytes per launch of each GEMM instance at C2, from the `ncu --set full` capture of this round's binary
    (profiles/r2_traffic.json, written by profiles/summarize.py from the .ncu-rep); not measurable outside ncu."""
    for name in ("r2_traffic.json", "r1c_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as fh:
                return json.load(fh), name
    return None, None


def roofline_block(args, dims, taps, kernel_ms, clips_per_pass, step_ms, peaks):
    """roofline of the dominant kernel family (the tcgen05 GEMMs) + the HBM-bound kernels, from per-launch events."""
    total_flops, gemm_flops = flops_per_clip(dims, args.frames, taps, executed=True)
    gemm_tags = [k for k in kernel_ms if k.startswith("gemm_")]
    gemm_ms = sum(kernel_ms[k][0] for k in gemm_tags)
    gemm_launches = sum(kernel_ms[k][1] for k in gemm_tags)
    achieved = gemm_flops * clips_per_pass / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic = tname = None
    if args.workload == "c2" and args.arch == "ViT-B/16" and args.clips == 64 and args.frames == 8 and not args.adapter:
        tj, tname = load_traffic()
        if tj:
            tsum = sum(tj[k] * kernel_ms[k][1] for k in gemm_tags if k in tj)
            tcnt = sum(kernel_ms[k][1] for k in gemm_tags if k in tj)
            traffic = tsum / tcnt if tcnt else None
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_2sm_kernel (tcgen05 cta_group::2, all epilogues)", "achieved": achieved,
        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
        "peak_source": "%s (sustained bf16 GEMM)" % peaks["source"], "traffic": traffic,
        "traffic_note": "average DRAM bytes per GEMM launch from the ncu --set full capture profiles/%s" % tname,
        "launches_per_step": gemm_launches, "ms_per_step": gemm_ms,
        "share_of_step": gemm_ms / step_ms if step_ms else None,
        "whole_step_tflops": total_flops * clips_per_pass / (step_ms * 1e-3) / 1e12,
        "whole_step_frac": total_flops * clips_per_pass / (step_ms * 1e-3) / 1e12 / peaks["tflops"],
        "by_kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(kernel_ms.items())},
    }
    seq = (dims["image_size"] // dims["patch_size"]) ** 2 + 1
    hbm = {}
    if "layernorm" in kernel_ms and kernel_ms["layernorm"][1] > 0:
        ln_ms, ln_n = kernel_ms["layernorm"]
        ln_bytes = clips_per_pass * args.frames * seq * dims["width"] * (4 + 2) * ln_n  # fp32 row in, bf16 row out
        hbm["layernorm"] = {"bytes_per_step": ln_bytes, "launches_per_step": ln_n,
                            "achieved_gbs": ln_bytes / (ln_ms * 1e-3) / 1e9}
    if "dec_attn" in kernel_ms and kernel_ms["dec_attn"][1] > 0:
        da_ms, _ = kernel_ms["dec_attn"]
        # K and V of every patch token of every tapped layer, bf16, read once
        da_bytes = 2 * clips_per_pass * args.frames * (seq - 1) * dims["width"] * 2 * len(taps)
        hbm["dec_attn"] = {"bytes_per_step": da_bytes, "launches_per_step": kernel_ms["dec_attn"][1],
                           "achieved_gbs": da_bytes / (da_ms * 1e-3) / 1e9,
                           "note": "duration includes the cross-unit combine kernel"}
    for v in hbm.values():
        v["peak_gbs"] = peaks["hbm_gbs"]
        v["frac"] = v["achieved_gbs"] / peaks["hbm_gbs"] if peaks["hbm_gbs"] else None
    roofline["hbm_kernels"] = hbm
    return roofline


def base_line(args, world, value, step_ms, dims, clocks, scaling="weak"):
    return {
        "metric": metric_name(args, dims["image_size"]), "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "clocks": clocks.summary(),
    }


# ------------------------------------------------------------------------------- b200 arm: C2 / C4 (predict)
def run_predict(args, rank, world, local_rank):
    from dfdclip_b200 import synthetic
    dev, dist = init_device(world, local_rank)
    dims = synthetic.vit_dims(args.arch)
    det, _ = build_detector(args.arch, args.frames, dev, adapter=args.adapter,
                            taps=[int(t) for t in args.taps.split(",")] if args.taps else None)
    taps = det.layer_indices
    clips, frames, res = args.clips, args.frames, dims["image_size"]

    x_host, m_host = synthetic.make_clips(clips, frames, res, seed=7 + rank, masked_tail=False)
    x_host, m_host = x_host.pin_memory(), m_host.pin_memory()
    x, m = x_host.to(dev), m_host.to(dev)
    # every step's clip scores stay on the device; ONE all_gather after the last step assembles the job's scores
    scores = torch.empty((args.steps, clips, 2), device=dev)
    gathered = [torch.empty_like(scores) for _ in range(world)] if dist else None

    @torch.no_grad()  # as every inference caller of the reference does (inference.py:65, evaluator.py:50)
    def step(k=0):
        logits, _ = det.predict(x, m)
        scores[k].copy_(logits[0])

    for _ in range(max(args.warmup, 3)):
        step()
    if dist:
        dist.all_gather(gathered, scores)
    torch.cuda.synchronize()

    # ---- timed region: K steps + the final gather, device-resident inputs
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev.index if dev.index is not None else 0) as clocks:
        start.record()
        for k in range(args.steps):
            step(k)
        if dist:
            dist.all_gather(gathered, scores)
        stop.record()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    elapsed_ms = max_over_ranks(start.elapsed_time(stop), dev, dist)
    step_ms = elapsed_ms / args.steps
    value = world * clips * args.steps / (elapsed_ms * 1e-3)

    kernel_ms = timed_kernels(dev, step, max(3, min(args.steps, 10)))

    # ---- same metric end to end through the public API with HOST buffers (pinned H2D in, logits D2H out)
    e2e = e2e_u8 = e2e_single = None
    if not args.no_e2e:
        from dfdclip_b200.inference import HostClipStream, predict_from_host

        def timed_stream(xh, overlap):
            """K batches through the streaming caller loop (H2D of batch k+1, encoder of batch k and decoder + D2H of
            batch k-1 in flight together); every batch's H2D and D2H and, at N > 1, the final all_gather of the job's
            scores are inside the timed region."""
            pipe = HostClipStream(det, overlap_decoder=overlap)
            for _ in pipe.run((xh, m_host) for _ in range(3)):
                pass
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outs = list(pipe.run((xh, m_host) for _ in range(args.steps)))
            if dist:
                scores.copy_(torch.stack(outs), non_blocking=True)
                dist.all_gather(gathered, scores)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0, dev, dist)
            assert sum(o.shape[0] for o in outs) == clips * args.steps
            return world * clips * args.steps / dt, outs[-1]

        overlap = not args.no_overlap_decoder
        v, host_logits = timed_stream(x_host, overlap)
        e2e = {"value": v, "unit": UNIT,
               "h2d_bytes_per_step": x_host.numel() * x_host.element_size() + m_host.numel(),
               "d2h_bytes_per_step": host_logits.numel() * host_logits.element_size(),
               "api": "dfdclip_b200.inference.predict_stream (pinned fp32 clips in, host logits out; "
                      "copy / encoder / decoder streams pipelined across batches, overlap_decoder=%s)" % overlap}
        if args.ab_overlap:
            e2e["value_other_overlap_setting"] = timed_stream(x_host, not overlap)[0]
        # one blocking call per batch (latency mode): chunked encoder, nothing overlaps across batches
        for _ in range(2):
            predict_from_host(det, x_host, m_host)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            predict_from_host(det, x_host, m_host)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, dev, dist)
        e2e_single = {"value": world * clips * args.steps / dt, "unit": UNIT,
                      "api": "dfdclip_b200.inference.predict_from_host, one blocking call per batch"}
        # same streaming call on raw uint8 pixels (Detector.transform_uint8): float conversion + normalisation fused
        # into the patch extraction kernel, 1 byte per pixel over PCIe
        g8 = torch.Generator().manual_seed(70 + rank)
        x8_host = torch.randint(0, 256, tuple(x_host.shape), generator=g8, dtype=torch.uint8).pin_memory()
        v8, _ = timed_stream(x8_host, overlap)
        e2e_u8 = {"value": v8, "unit": UNIT, "h2d_bytes_per_step": x8_host.numel() + m_host.numel(),
                  "d2h_bytes_per_step": host_logits.numel() * host_logits.element_size(),
                  "input": "uint8 pixels, normalisation fused into patchify"}

    if rank != 0:
        return
    peaks = load_peaks()
    total_flops, _ = flops_per_clip(dims, frames, taps, executed=True)
    ref_flops, _ = flops_per_clip(dims, frames, taps, executed=False)
    line = base_line(args, world, value, step_ms, dims, clocks)
    line["config"] = {
        "workload": "%s: %s encoder + DFD head eval, %d synthetic clips x %d frames x %d^2 per GPU per step" % (
            WORKLOADS[args.workload]["label"] if (clips, frames, args.arch) == tuple(
                WORKLOADS[args.workload][k] for k in ("clips", "frames", "arch")) else "custom",
            args.arch, clips, frames, res) + (" + CompInvAdapter %s" % args.adapter if args.adapter else ""),
        "arch": args.arch, "clips_per_gpu": clips, "frames": frames, "taps": taps, "adapter": args.adapter,
        "parallelism": "dp%d" % world, "l2": "inputs_exceed_l2 (%.0f MB of fp32 frames per step)" % (x.numel() * 4 / 1e6),
        "collective": "one all_gather of all steps' clip scores after the last step, inside the timed region",
        "flops_per_clip_executed": total_flops, "flops_per_clip_reference": ref_flops,
        "kernel_timing": "separate pass with one CUDA event pair per launch"}
    line.update({
        "e2e": e2e, "e2e_u8": e2e_u8, "e2e_single_call": e2e_single,
        "gpu_launches": launches_per_predict(max(taps), len(taps), 1, bool(args.adapter)) * args.steps,
        "roofline": roofline_block(args, dims, taps, kernel_ms, clips, step_ms, peaks),
        "cpu_baseline": cpu_baseline(args) if world == 1 and not args.no_cpu_baseline else None,
    })
    emit(line)


# ---------------------------------------------------------------------------- b200 arm: C3 (video-level job)
def run_c3(args, rank, world, local_rank):
    """BASELINE config C3: 560 synthetic videos (FF++ test split: 280 real + 280 of one manipulation) of U{8..32}
    one-second clips (seed 3), uint8 frames in page-locked host memory (what a pinning DataLoader delivers), sharded
    over the ranks by clip count; each rank scores its shard through score_videos_batched (HostClipStream: H2D,
    encoder and decoder + D2H of consecutive 64-clip batches in flight together) and ONE all_gather assembles the
    per-video mean probabilities (inference.py:107-156). A step = the whole job; total work is fixed (strong scaling)."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import score_videos_batched, shard_videos
    dev, dist = init_device(world, local_rank)
    dims = synthetic.vit_dims(args.arch)
    det, _ = build_detector(args.arch, args.frames, dev)
    taps, frames, res = det.layer_indices, args.frames, dims["image_size"]
    counts = torch.randint(8, 33, (args.videos,), generator=torch.Generator().manual_seed(3)).tolist()
    mine = shard_videos(counts, world)[rank]
    my_clips = sum(counts[i] for i in mine)
    dtype = torch.float32 if args.c3_fp32 else torch.uint8
    gr = torch.Generator().manual_seed(100 + rank)
    # this rank's clips: one page-locked pool, each video a contiguous run of it (no two videos share pixels)
    if dtype == torch.uint8:
        pool = torch.randint(0, 256, (my_clips, frames, 3, res, res), generator=gr, dtype=torch.uint8).pin_memory()
    else:
        pool = torch.randn((my_clips, frames, 3, res, res), generator=gr).pin_memory()
    mask_pool = torch.ones((my_clips, frames), dtype=torch.bool).pin_memory()
    videos, masks, off = [], [], 0
    mine_set = set(mine)
    for i, n in enumerate(counts):
        if i in mine_set:
            videos.append(pool[off:off + n])
            masks.append(mask_pool[off:off + n])
            off += n
        else:  # other ranks' videos only carry their shape
            videos.append(torch.zeros((), dtype=dtype).expand(n, frames, 3, res, res))
            masks.append(torch.ones((), dtype=torch.bool).expand(n, frames))

    @torch.no_grad()
    def job():
        return score_videos_batched(det, videos, masks, batch_clips=args.clips)

    for _ in range(max(1, min(args.warmup, 2))):
        result = job()
    torch.cuda.synchronize()
    steps = args.steps
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index if dev.index is not None else 0) as clocks:
        start.record()
        for _ in range(steps):
            result = job()
        stop.record()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    elapsed_ms = max_over_ranks(start.elapsed_time(stop), dev, dist)
    step_ms = elapsed_ms / steps
    total_clips = sum(counts)
    value = total_clips * steps / (elapsed_ms * 1e-3)
    same = True
    if dist:
        chk = torch.nan_to_num(result).double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).item())
    # per-kernel durations: one device-resident batch of the job through the same predict (the whole job would need more
    # event pairs than the library keeps; per-clip kernel time does not depend on which batch is timed)
    xb, mb = pool[:args.clips].to(dev), mask_pool[:args.clips].to(dev)
    with torch.no_grad():
        kernel_ms = timed_kernels(dev, lambda: det.predict(xb, mb), 3)
    if rank != 0:
        return
    peaks = load_peaks()
    line = base_line(args, world, value, step_ms, dims, clocks, scaling="strong")
    total_flops, _ = flops_per_clip(dims, frames, taps, executed=True)
    n_batches = (my_clips + args.clips - 1) // args.clips
    bytes_in = pool.numel() * pool.element_size() + mask_pool.numel()
    line["config"] = {
        "workload": "C3: FF++ cross-manipulation test shape, %d synthetic videos / %d clips (U{8..32} per video, seed "
                    "3) x %d frames x %d^2, %s, video-level mean of clip probabilities, videos sharded over %d GPU(s) "
                    "by clip count" % (args.videos, total_clips, frames, res, args.arch, world),
        "arch": args.arch, "videos": args.videos, "clips_total": total_clips, "clips_on_rank0": my_clips,
        "frames": frames, "taps": taps, "input": "uint8 pixels" if dtype == torch.uint8 else "fp32 normalised frames",
        "host_memory": "page-locked (a pinning DataLoader)", "batch_clips": args.clips, "parallelism": "dp%d" % world,
        "collective": "one all_gather of per-video scores at the end of the job",
        "l2": "inputs_exceed_l2 (every batch is %.0f MB of fresh host pixels)" % (
            args.clips * frames * 3 * res * res * pool.element_size() / 1e6),
        "videos_per_s": args.videos * steps / (elapsed_ms * 1e-3), "scores_identical_on_all_ranks": same,
        "scores_finite": bool(torch.isfinite(result).all().item()), "flops_per_clip_executed": total_flops}
    # the job IS the end-to-end path: host pixels in, host-visible scores out, every copy inside the timed region
    line["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": bytes_in * world if world == 1 else None,
                   "h2d_bytes_per_step_rank0": bytes_in, "d2h_bytes_per_step": my_clips * 2 * 4,
                   "api": "dfdclip_b200.inference.score_videos_batched (shard_videos + HostClipStream + one all_gather); "
                          "`value` is this same measurement: the job has no device-resident variant"}
    line["gpu_launches"] = launches_per_predict(max(taps), len(taps), 1) * n_batches * steps
    # whole_step_*: the job's time apportioned to one batch of this rank
    line["roofline"] = roofline_block(args, dims, taps, kernel_ms, xb.shape[0], step_ms * xb.shape[0] / my_clips, peaks)
    line["cpu_baseline"] = cpu_baseline(args) if world == 1 and not args.no_cpu_baseline else None
    emit(line)


# ------------------------------------------------------------------------------ b200 arm: C5 (training step)
def run_c5(args, rank, world, local_rank):
    """BASELINE config C5: the reference trainer's step (src/trainer.py:147-178; DDP of main.py:283-287) — frozen
    encoder forward with K/V taps, decoder forward / backward, gradient all-reduce across the ranks, SGD — on 12 clips
    x 8 frames per GPU, through dfdclip_b200.training.TrainStep (one CUDA graph per step, the NCCL all-reduce inside)."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.training import TrainStep
    dev, dist = init_device(world, local_rank)
    dims = synthetic.vit_dims(args.arch)
    det, _ = build_detector(args.arch, args.frames, dev)
    det.train()
    taps, clips, frames, res = det.layer_indices, args.clips, args.frames, dims["image_size"]
    opt = det.configure_optimizers(lr=1e-3)
    g = torch.Generator().manual_seed(5 + rank)
    x_host = torch.randn((clips, frames, 3, res, res), generator=g).pin_memory()
    y_host = torch.randint(0, 2, (clips,), generator=g).pin_memory()
    m_host = torch.ones((clips, frames), dtype=torch.bool).pin_memory()
    x, y, m = x_host.to(dev), y_host.to(dev), m_host.to(dev)
    step = TrainStep(det, opt, x, y, m, group=(dist.group.WORLD if dist else None), pipeline=bool(args.c5_pipeline))

    for _ in range(max(args.warmup, 3)):
        step(x, y, m)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev.index if dev.index is not None else 0) as clocks:
        start.record()
        for _ in range(args.steps):
            loss, _ = step(x, y, m)
        stop.record()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    elapsed_ms = max_over_ranks(start.elapsed_time(stop), dev, dist)
    step_ms = elapsed_ms / args.steps
    value = world * clips * args.steps / (elapsed_ms * 1e-3)
    # end to end: every step's batch comes from pinned host memory and its loss goes back to the host
    for loss, _ in step.run_host((x_host, y_host, m_host) for _ in range(3)):   # untimed: copy stream, staging buffers
        loss.item()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stamps = []
    for loss, _ in step.run_host((x_host, y_host, m_host) for _ in range(args.steps)):
        loss_host = loss.item()
        stamps.append(time.perf_counter())
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, dev, dist)
    gaps = sorted((b - a) * 1e3 for a, b in zip([t0] + stamps, stamps))
    print("c5 e2e: per-step host time min %.2f / median %.2f / max %.2f ms" % (gaps[0], gaps[len(gaps) // 2], gaps[-1]),
          file=sys.stderr)
    same = True
    if dist:  # replicas must stay identical
        chk = torch.stack([p.detach().double().sum() for p in det.decoder.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).item())
    # roofline of the encoder GEMMs inside the step: an eager (un-graphed) pass of the same step with launch events
    kernel_ms = timed_kernels(dev, lambda: step.eager(x, y, m, serial=True), 3)
    description, launches = step.describe(), step.launches_per_step
    step.close()   # the graph holds captured NCCL plans: it must go before the process group does
    if rank != 0:
        return
    peaks = load_peaks()
    total_flops, _ = flops_per_clip(dims, frames, taps, executed=True)
    n_grad = sum(p.numel() for p in det.parameters() if p.requires_grad)
    line = base_line(args, world, value, step_ms, dims, clocks)
    line["config"] = {
        "workload": "C5: %s frozen-encoder training step (encoder forward with K/V taps + decoder forward/backward + "
                    "SGD), %d synthetic clips x %d frames x %d^2 per GPU per step" % (args.arch, clips, frames, res),
        "arch": args.arch, "clips_per_gpu": clips, "frames": frames, "taps": taps, "parallelism": "dp%d" % world,
        "trainable_parameters": n_grad, "optimizer": "SGD momentum 0.95 (Detector.configure_optimizers)",
        "collective": ("gradient all-reduce (%d fp32 = %.0f MB) inside the captured step" % (n_grad, n_grad * 4 / 1e6))
        if world > 1 else "none (1 GPU)", "replicas_identical": same, "final_loss": loss_host,
        "l2": "flushed_by_the_step (each step streams %.0f MB of taps and %.0f MB of parameters, gradients and "
              "momentum, far more than L2)" % (2 * clips * frames * 196 * 768 * 2 * len(taps) / 1e6, n_grad * 12 / 1e6),
        "flops_per_clip_executed": total_flops, "step": description}
    line["e2e"] = {"value": world * clips * args.steps / dt, "unit": UNIT,
                   "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8 + m_host.numel(),
                   "d2h_bytes_per_step": 4, "api": "dfdclip_b200.training.TrainStep.run_host: pinned host batches, the "
                   "H2D copy of batch k+1 overlaps step k, loss.item() per step"}
    line["gpu_launches"] = launches * args.steps
    line["roofline"] = roofline_block(args, dims, taps, kernel_ms, clips, step_ms, peaks)
    line["cpu_baseline"] = cpu_baseline(args) if world == 1 and not args.no_cpu_baseline else None
    emit(line)


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, written to the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line. Libraries write there too (NCCL prints its version banner to stdout under
    # NCCL_DEBUG=WARN, which the GPU boxes set), so file descriptor 1 is pointed at stderr for the whole run and the
    # result goes to a private duplicate of the original stdout.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration: c2 (default, the one the metric is quoted on), c3 video-level "
                         "job, c4 ViT-L/14 x 16 frames, c5 training step")
    ap.add_argument("--arch", default=None, help="default: the workload's architecture")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU per step (c3: clips per batch)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--videos", type=int, default=560, help="c3: number of synthetic videos")
    ap.add_argument("--c3-fp32", action="store_true", help="c3: fp32 normalised frames instead of uint8 pixels")
    ap.add_argument("--adapter", default=None, help="adapter.struct.type of a CompInvAdapter (x=256) on the taps, "
                    "e.g. 768-x-768-nln as in the shipped configs; default: none (BASELINE config C2)")
    ap.add_argument("--taps", default=None, help="comma-separated decode_indices (decode_mode=index), e.g. "
                    "6,7,8,9,10,11 as in the shipped configs; default: stride-2 taps")
    ap.add_argument("--ref-clips", type=int, default=None, help="clips per step of the CPU reference arm / "
                    "cpu_baseline sample (default: 16 = the reference's own chunk size, scripts/inference.sh; 2 for "
                    "ViT-L/14 x 16 frames, 4 for the training step)")
    ap.add_argument("--c5-pipeline", type=int, default=0, choices=[0, 1],
                    help="c5: TrainStep(pipeline=...): encode batch k+1 beside the decoder step of batch k")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap-decoder", action="store_true",
                    help="e2e stream: run the decoder on the encoder's stream instead of beside the next batch's encoder")
    ap.add_argument("--ab-overlap", action="store_true", help="e2e stream: also time the other overlap setting")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.arch = args.arch or wl["arch"]
    args.clips = args.clips or wl["clips"]
    args.frames = args.frames or wl["frames"]
    if args.ref_clips is None:
        args.ref_clips = {"c4": 2, "c5": 4}.get(args.workload, 16)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    done = False
    try:
        {"c3": run_c3, "c5": run_c5}.get(args.workload, run_predict)(args, rank, world, local_rank)
        done = True
    finally:
        if world > 1 and torch.distributed.is_initialized():
            # the result line is out; a communicator teardown that does not finish must not hold the box
            guard = threading.Timer(60.0, lambda: os._exit(0 if done else 1))
            guard.daemon = True
            guard.start()
            torch.cuda.synchronize()
            torch.distributed.destroy_process_group()
            guard.cancel()


if __name__ == "__main__":
    main()

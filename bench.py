#!/usr/bin/env python
"""Benchmark of the DFD-CLIP hot path on B200: clips/sec of Detector.predict (CLIP ViT frame encoder with K/V
taps + temporal decoder/head) on synthetic clips, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic clips per GPU (config C2: 64 clips x 8 frames
x 224^2, ViT-B/16, taps [0,2,..,10], bf16 tensor-core math / fp32 accumulate). Under torchrun every rank runs the
same per-GPU batch (weak scaling, clips are independent units) and the per-clip scores are all-gathered once
per step (the "final logit/score gather" of inference.py:147). Prints ONE JSON line on rank 0.

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

METRIC = "clips/sec (8x224^2 frames, ViT-B/16 enc+head)"
UNIT = "clips/s"


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
    enc = 4 + 7 * layers_run_full + 2
    dec = 2 + 9 * n_taps + 1 + n_tasks
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


def cpu_reference_clips_per_sec(arch, frames, clips, repeats, threads=None):
    """The reference's algorithm on the host cores: the oracle port (torch fp32), Detector.predict semantics."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfd_oracle
    from dfdclip_b200 import synthetic
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    dims = synthetic.vit_dims(arch)
    taps = synthetic.layer_indices(arch)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=taps, seed=0)
    x, m = synthetic.make_clips(clips, frames, dims["image_size"], seed=7, masked_tail=False)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            dfd_oracle.detector_predict(sd, x, m, taps, (2,))
            times.append(time.perf_counter() - t0)
    return clips / min(times), times, torch.get_num_threads()


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    clips = args.ref_clips  # a bounded sample of the C2 batch: enough rows to keep every host core busy
    torch.set_num_threads(os.cpu_count() or 1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfd_oracle
    from dfdclip_b200 import synthetic
    dims = synthetic.vit_dims(args.arch)
    taps = synthetic.layer_indices(args.arch)
    sd = synthetic.detector_state_dict(args.arch, args.frames, out_dims=(2,), taps=taps, seed=0)
    x, m = synthetic.make_clips(clips, args.frames, dims["image_size"], seed=7, masked_tail=False)
    with torch.no_grad():
        for _ in range(args.warmup):
            dfd_oracle.detector_predict(sd, x, m, taps, (2,))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dfd_oracle.detector_predict(sd, x, m, taps, (2,))
        dt = time.perf_counter() - t0
    value = clips * args.steps / dt
    cores = torch.get_num_threads()
    sample = "%d synthetic clip(s) x %d frames per step, Detector.predict restated in torch fp32 (oracle port)" % (
        clips, args.frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 sample: %s encoder+head, %d frames/clip, %d clips per step on host CPU" % (
            args.arch, args.frames, clips), "arch": args.arch, "frames": args.frames, "clips_per_step": clips},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- b200 arm
def run_b200(args, rank, world, local_rank):
    from dfdclip_b200 import _native, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    dims = synthetic.vit_dims(args.arch)
    det, _ = build_detector(args.arch, args.frames, dev, adapter=args.adapter,
                            taps=[int(t) for t in args.taps.split(",")] if args.taps else None)
    taps = det.layer_indices
    clips, frames, res = args.clips, args.frames, dims["image_size"]

    x_host, m_host = synthetic.make_clips(clips, frames, res, seed=7 + rank, masked_tail=False)
    x_host, m_host = x_host.pin_memory(), m_host.pin_memory()
    x, m = x_host.to(dev), m_host.to(dev)
    gathered = [torch.empty((clips, 2), device=dev) for _ in range(world)] if dist else None

    @torch.no_grad()  # as every inference caller of the reference does (inference.py:65, evaluator.py:50)
    def step():
        logits, _ = det.predict(x, m)
        if dist:
            dist.all_gather(gathered, logits[0])
        return logits[0]

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs, nothing but the hot path between the two events
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev.index if dev.index is not None else 0) as clocks:
        start.record()
        for _ in range(args.steps):
            out = step()
        stop.record()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    elapsed_ms = start.elapsed_time(stop)
    if dist:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = t.item()

    # ---- per-kernel durations for the roofline: the same steps again with a CUDA event pair around every launch
    # (recorded by the library on the launching stream). Kept out of the headline region: ~260 event records per
    # step cost 2-3 % of the step time.
    kt_steps = max(3, min(args.steps, 10))
    _native.timing_enable(dev, True)
    torch.cuda.synchronize()
    for _ in range(kt_steps):
        step()
    torch.cuda.synchronize()
    kernel_ms = _native.timing_read(dev)
    _native.timing_enable(dev, False)
    kernel_ms = {k: (v[0] * args.steps / kt_steps, v[1] * args.steps / kt_steps) for k, v in kernel_ms.items()}
    value = world * clips * args.steps / (elapsed_ms * 1e-3)

    # ---- same metric end to end through the public API with HOST buffers (pinned H2D in, logits D2H out)
    e2e = e2e_u8 = e2e_single = None
    if not args.no_e2e:
        from dfdclip_b200.inference import HostClipStream, predict_from_host

        def timed_stream(xh, overlap):
            """K batches through the streaming caller loop (H2D of batch k+1, encoder of batch k and decoder + D2H of
            batch k-1 in flight together); every batch's H2D and D2H is inside the timed region, the first copy and
            the last decoder are exposed."""
            pipe = HostClipStream(det, overlap_decoder=overlap)
            for _ in pipe.run((xh, m_host) for _ in range(3)):
                pass
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_out = 0
            for out in pipe.run((xh, m_host) for _ in range(args.steps)):
                n_out += out.shape[0]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert n_out == clips * args.steps
            if dist:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = t.item()
            return world * clips * args.steps / dt, out

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
        dt = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        e2e_single = {"value": world * clips * args.steps / dt, "unit": UNIT,
                      "api": "dfdclip_b200.inference.predict_from_host, one blocking call per batch"}
        # same streaming call on raw uint8 pixels (Detector.transform_uint8): float conversion + normalisation fused
        # into the patch extraction kernel, 1 byte per pixel over PCIe. Reported next to `e2e`, which keeps the
        # reference's fp32 clip format.
        g8 = torch.Generator().manual_seed(70 + rank)
        x8_host = torch.randint(0, 256, tuple(x_host.shape), generator=g8, dtype=torch.uint8).pin_memory()
        v8, _ = timed_stream(x8_host, overlap)
        e2e_u8 = {"value": v8, "unit": UNIT,
                  "h2d_bytes_per_step": x8_host.numel() + m_host.numel(),
                  "d2h_bytes_per_step": host_logits.numel() * host_logits.element_size(),
                  "input": "uint8 pixels, normalisation fused into patchify"}

    if rank != 0:
        return
    peaks = load_peaks()
    total_flops, gemm_flops = flops_per_clip(dims, frames, taps, executed=True)
    ref_flops, _ = flops_per_clip(dims, frames, taps, executed=False)
    gemm_tags = [k for k in kernel_ms if k.startswith("gemm_")]
    gemm_ms_per_step = sum(kernel_ms[k][0] for k in gemm_tags) / args.steps
    gemm_launches = sum(kernel_ms[k][1] for k in gemm_tags) / args.steps
    achieved = gemm_flops * clips / (gemm_ms_per_step * 1e-3) / 1e12 if gemm_ms_per_step > 0 else None
    # DRAM traffic of the GEMM family per launch, from the committed ncu capture of the same step (not measured live)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1c_traffic.json")
    if os.path.exists(tpath) and args.arch == "ViT-B/16" and clips == 64 and frames == 8:
        with open(tpath) as fh:
            tj = json.load(fh)
        tsum = sum(tj[k] * kernel_ms[k][1] for k in gemm_tags if k in tj)
        tcnt = sum(kernel_ms[k][1] for k in gemm_tags if k in tj)
        traffic = tsum / tcnt if tcnt else None
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_2sm_kernel (tcgen05 cta_group::2, all epilogues)", "achieved": achieved,
        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
        "peak_source": "%s (sustained bf16 GEMM)" % peaks["source"], "traffic": traffic,
        "traffic_note": "average DRAM bytes per GEMM launch (ncu capture in profiles/r1c_kernels.md)",
        "launches_per_step": gemm_launches, "ms_per_step": gemm_ms_per_step,
        "share_of_step": gemm_ms_per_step / (elapsed_ms / args.steps),
        "whole_step_tflops": total_flops * clips * world / (elapsed_ms / args.steps * 1e-3) / 1e12 / world,
        "whole_step_frac": total_flops * clips / (elapsed_ms / args.steps * 1e-3) / 1e12 / peaks["tflops"],
        "by_kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(kernel_ms.items())},
    }
    # HBM-bound sub-kernels: algorithmic bytes / measured duration against the measured copy bandwidth
    seq = (dims["image_size"] // dims["patch_size"]) ** 2 + 1
    m_rows = clips * frames * seq
    hbm = {}
    if "layernorm" in kernel_ms:
        ln_ms, ln_n = kernel_ms["layernorm"]
        ln_bytes = m_rows * dims["width"] * (4 + 2)          # fp32 row in, bf16 row out
        hbm["layernorm"] = {"bytes_per_launch": ln_bytes, "launches_per_step": ln_n / args.steps,
                            "achieved_gbs": ln_bytes * ln_n / (ln_ms * 1e-3) / 1e9}
    if "dec_attn" in kernel_ms:
        da_ms, da_n = kernel_ms["dec_attn"]
        da_bytes = 2 * clips * frames * (seq - 1) * dims["width"] * 2   # K and V of every patch token, bf16, once
        hbm["dec_attn"] = {"bytes_per_launch": da_bytes, "launches_per_step": da_n / args.steps,
                           "achieved_gbs": da_bytes * da_n / (da_ms * 1e-3) / 1e9,
                           "note": "duration includes the cross-unit combine kernel"}
    for v in hbm.values():
        v["peak_gbs"] = peaks["hbm_gbs"]
        v["frac"] = v["achieved_gbs"] / peaks["hbm_gbs"] if peaks["hbm_gbs"] else None
    roofline["hbm_kernels"] = hbm
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, times, cores = cpu_reference_clips_per_sec(args.arch, frames, args.ref_clips, repeats=4)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d clips x %d frames, best of %d runs of the oracle port (torch fp32) of Detector.predict" % (
                   args.ref_clips, frames, len(times))}
    n_full = max(taps)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2: %s encoder + DFD head eval, %d synthetic clips x %d frames x %d^2 per GPU per step" % (
            args.arch, clips, frames, res) + (" + CompInvAdapter %s" % args.adapter if args.adapter else ""),
            "arch": args.arch, "clips_per_gpu": clips, "frames": frames, "taps": taps, "adapter": args.adapter,
            "parallelism": "dp%d" % world, "l2": "inputs_exceed_l2 (%.0f MB of fp32 frames per step)" % (
                x.numel() * 4 / 1e6), "flops_per_clip_executed": total_flops, "flops_per_clip_reference": ref_flops,
            "kernel_timing": "separate pass of %d steps with one CUDA event pair per launch" % kt_steps},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "e2e_u8": e2e_u8,
        "e2e_single_call": e2e_single,
        "gpu_launches": launches_per_predict(n_full, len(taps), 1, bool(args.adapter)) * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
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
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--adapter", default=None, help="adapter.struct.type of a CompInvAdapter (x=256) on the taps, "
                    "e.g. 768-x-768-nln as in the shipped configs; default: none (BASELINE config C2)")
    ap.add_argument("--taps", default=None, help="comma-separated decode_indices (decode_mode=index), e.g. "
                    "6,7,8,9,10,11 as in the shipped configs; default: stride-2 taps")
    ap.add_argument("--ref-clips", type=int, default=1, help="clips per step of the CPU reference arm / cpu_baseline "
                    "(1 clip = 8 frames measured fastest per clip on the host: 5.4 vs 4.3 clips/s at 4 clips)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap-decoder", action="store_true",
                    help="e2e stream: run the decoder on the encoder's stream instead of beside the next batch's encoder")
    ap.add_argument("--ab-overlap", action="store_true", help="e2e stream: also time the other overlap setting")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1 and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

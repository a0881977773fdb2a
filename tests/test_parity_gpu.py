"""GPU parity tests proper: the drop-in Detector / VisionTransformer / Decoder (CUDA path through the C ABI)
against (1) golden vectors of the unmodified reference and (2) the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's north_star: per-layer features cosine >= 0.999, clip logits within 2e-2 absolute,
identical predicted labels (the CUDA path computes the encoder in bf16 with fp32 accumulation, the reference fp32).
"""
import numpy as np
import pytest
import torch

from helpers import (TOL_FEATURE_COSINE, TOL_LOGIT_ABS, cosine, golden_inputs, golden_tensor, load_golden,
                     load_oracle)

pytestmark = pytest.mark.gpu


def build_detector(arch, num_frames, layer_indices, device, sd=None, decode_indices=None):
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    if decode_indices is not None:
        cfg.decode_mode = "index"
        cfg.decode_indices = list(decode_indices)
    det = Detector(cfg, num_frames, None)
    if sd is None:
        sd = synthetic.detector_state_dict(arch, num_frames, out_dims=(2,), taps=det.layer_indices, seed=0)
    det.load_state_dict(sd, strict=True)
    assert det.layer_indices == list(layer_indices)
    return det.to(device).eval(), sd


@pytest.mark.parametrize("case", ["tiny", "small", "vitb16", "vitl14"])
def test_detector_predict_matches_reference_golden(cuda_device, case):
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    det, _ = build_detector(g["arch"], g["num_frames"], g["layer_indices"], cuda_device, sd)
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
    torch.cuda.synchronize()
    got = logits[0].cpu().numpy()
    assert got.shape == g["logits"].shape
    err = np.abs(got - g["logits"]).max()
    assert err <= TOL_LOGIT_ABS, f"{case}: max |dlogit| {err:.4f}"
    assert np.array_equal(got.argmax(-1), g["pred_labels"])
    assert np.allclose(np.linalg.norm(got, axis=-1), 5.0, atol=1e-3)
    assert cosine(feats["video"].cpu(), torch.from_numpy(g["video_feature"])) >= TOL_FEATURE_COSINE
    # eval forward: per-sample cross entropy on the normalised logits (src/models.py:590-596)
    labels = torch.from_numpy(g["labels"]).to(cuda_device)
    losses, logits2 = det(x.to(cuda_device), [labels], m.to(cuda_device), single_task=0)
    assert np.abs(losses[0].cpu().numpy() - g["losses"]).max() <= 2 * TOL_LOGIT_ABS
    assert torch.equal(logits2[0], logits[0])


@pytest.mark.parametrize("case", ["tiny", "small", "vitb16", "vitl14"])
def test_encoder_taps_match_reference_golden(cuda_device, case):
    """VisionTransformer.forward(x, with_out=True, with_q=True): every layer's q/k/v/out (model.py:236-251)."""
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    det, _ = build_detector(g["arch"], g["num_frames"], g["layer_indices"], cuda_device, sd)
    kvs = det.encoder(x.flatten(0, 1).to(cuda_device), with_out=True, with_q=True)
    torch.cuda.synchronize()
    n = x.shape[0] * x.shape[1]
    assert len(kvs) == det.encoder.layers
    for layer, a in enumerate(kvs):
        assert tuple(a["k"].shape) == (n, det.encoder.tokens_per_frame, det.encoder.heads, 64)
        for key in ("q", "k", "v", "out"):
            ref, got = golden_tensor(g, key, layer, a[key])
            c = cosine(got, ref)
            assert c >= TOL_FEATURE_COSINE, f"{case} layer {layer} {key}: cosine {c:.5f}"
            rel = ((got - ref).norm() / ref.norm()).item()
            assert rel < 3e-2, f"{case} layer {layer} {key}: rel err {rel:.4f}"


def test_predict_against_oracle_with_masks_and_index_taps(cuda_device):
    """Same seeded inputs through the CPU oracle and the CUDA path; taps chosen by index (decode_mode='index')."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    arch, t, b = "small-512x6", 5, 7
    taps = [1, 2, 5]
    det, sd = build_detector(arch, t, taps, cuda_device, decode_indices=taps)
    x, m = synthetic.make_clips(b, t, synthetic.vit_dims(arch)["image_size"], seed=11)
    m[2, 1:] = False  # a clip with a single valid frame
    with torch.no_grad():
        ref_logits, ref_feat, ref_taps = oracle.detector_predict(sd, x, m, taps, (2,), return_taps=True)
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
    torch.cuda.synchronize()
    assert (logits[0].cpu() - ref_logits[0]).abs().max().item() <= TOL_LOGIT_ABS
    assert torch.equal(logits[0].cpu().argmax(-1), ref_logits[0].argmax(-1))
    assert cosine(feats["video"].cpu(), ref_feat) >= TOL_FEATURE_COSINE
    # per-layer tapped features
    qkv, _ = det.encoder.encode(x.flatten(0, 1).to(cuda_device), keep_layers=taps)
    seq, h = det.encoder.tokens_per_frame, det.encoder.heads
    for i, layer in enumerate(taps):
        view = qkv[layer].view(b, t, seq, 3, h, 64)
        for j, key in ((1, "k"), (2, "v")):
            c = cosine(view[:, :, 1:, j].float().cpu(), ref_taps[i][key])
            assert c >= TOL_FEATURE_COSINE, (layer, key, c)


def test_masked_frames_do_not_matter_cuda(cuda_device):
    from dfdclip_b200 import synthetic
    arch, t, b = "tiny-256x4", 4, 4
    det, _ = build_detector(arch, t, [0, 2], cuda_device)
    x, m = synthetic.make_clips(b, t, 32, seed=3)
    assert not m.all()
    x2 = x.clone()
    x2[~m] = 100.0
    a, _ = det.predict(x.to(cuda_device), m.to(cuda_device))
    c, _ = det.predict(x2.to(cuda_device), m.to(cuda_device))
    torch.cuda.synchronize()
    assert torch.equal(a[0], c[0])


def test_batch_independence_and_determinism(cuda_device):
    """Clips are independent units: a clip's logits do not depend on its batch mates; reruns are bit-identical."""
    from dfdclip_b200 import synthetic
    arch, t, b = "tiny-256x4", 4, 6
    det, _ = build_detector(arch, t, [0, 2], cuda_device)
    x, m = synthetic.make_clips(b, t, 32, seed=5)
    x, m = x.to(cuda_device), m.to(cuda_device)
    full, _ = det.predict(x, m)
    again, _ = det.predict(x, m)
    part, _ = det.predict(x[2:5], m[2:5])
    torch.cuda.synchronize()
    assert torch.equal(full[0], again[0])
    assert torch.equal(full[0][2:5], part[0])


def test_empty_batch(cuda_device):
    det, _ = build_detector("tiny-256x4", 4, [0, 2], cuda_device)
    x = torch.zeros(0, 4, 3, 32, 32, device=cuda_device)
    m = torch.zeros(0, 4, dtype=torch.bool, device=cuda_device)
    logits, _ = det.predict(x, m)
    assert tuple(logits[0].shape) == (0, 2)


def test_unsupported_configs_raise(cuda_device):
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:tiny-256x4"
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    cfg.train_mode.compression = "sync"
    with pytest.raises(NotImplementedError):
        Detector(cfg, 4, None)
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:tiny-256x4"
    cfg.adapter.type = "lora"
    with pytest.raises(NotImplementedError):
        Detector(cfg, 4, None)


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors, never fall back to PyTorch/CPU."""
    import dfdclip_b200._native as nat
    det_cfg = None
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:tiny-256x4"
    cfg.out_dim = [2]
    det = Detector(cfg, 4, None).eval()
    with pytest.raises(nat.NativeError):
        det.predict(torch.zeros(1, 4, 3, 32, 32), torch.ones(1, 4, dtype=torch.bool))


@pytest.mark.gpu
def test_predict_from_host_matches_predict(cuda_device):
    """The chunked-encoder / batched-decoder host pipeline gives bit-identical logits to one predict() call."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import HostClipPipeline
    from dfdclip_b200.models import Detector
    arch, frames, clips = "small-512x6", 3, 13
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    det = Detector(cfg, frames, None)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det.layer_indices, seed=0)
    det.load_state_dict(sd, strict=True)
    det = det.to(cuda_device).eval()
    x, m = synthetic.make_clips(clips, frames, synthetic.vit_dims(arch)["image_size"], seed=11)
    ref = det.predict(x.to(cuda_device), m.to(cuda_device))[0][0].cpu()
    pipe = HostClipPipeline(det, chunk_clips=4)
    pipe.MAX_BATCH = 9  # force two decoder passes as well
    got = pipe(x.pin_memory(), m.pin_memory())
    assert torch.equal(got, ref)
    again = pipe(x.pin_memory(), m.pin_memory())
    assert torch.equal(again, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("u8", [False, True])
def test_predict_stream_matches_predict(cuda_device, overlap, u8):
    """The cross-batch pipeline (copy stream / encoder / decoder stream, two slots) yields, batch by batch and in
    order, logits bit-identical to predict() on the same clips: ragged batch sizes (buffers shrink and grow), an empty
    batch in the middle, more batches than slots, fp32 and raw uint8 clips."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import HostClipStream
    arch, frames = "small-512x6", 3
    det, _ = build_detector(arch, frames, [0, 2, 4], cuda_device)
    res = synthetic.vit_dims(arch)["image_size"]
    sizes = [5, 2, 0, 7, 7, 1, 6]
    batches = []
    for i, n in enumerate(sizes):
        x, m = synthetic.make_clips(max(n, 1), frames, res, seed=40 + i)
        if u8:
            x = torch.randint(0, 256, tuple(x.shape), generator=torch.Generator().manual_seed(90 + i), dtype=torch.uint8)
        batches.append((x[:n].pin_memory(), m[:n].pin_memory()))
    want = []
    for x, m in batches:
        if x.shape[0] == 0:
            want.append(torch.empty((0, 2)))
        else:
            want.append(det.predict(x.to(cuda_device), m.to(cuda_device))[0][0].cpu())
    pipe = HostClipStream(det, overlap_decoder=overlap)
    for _ in range(2):  # second pass reuses the slot buffers
        got = list(pipe.run(iter(batches)))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.shape == w.shape and torch.equal(g, w)


@pytest.mark.gpu
@pytest.mark.parametrize("taps", [[0, 2, 4], [3, 4, 5], [4, 0, 2], [1, 1, 5]])
def test_single_call_predict_is_bit_identical_to_the_two_call_path(cuda_device, monkeypatch, taps):
    """dfd_predict_forward (decoder blocks on the context's side stream, each released by an event behind its tap's
    K/V projection) against dfd_encoder_forward followed by dfd_decoder_forward on one stream: same kernels, same
    arguments, so logits and video features must be bit-identical — also for taps that are not ascending or repeat a
    layer, for uint8 frames, for masked frames, and when called again and again without a host sync in between."""
    from dfdclip_b200 import synthetic
    arch, frames, clips = "small-512x6", 3, 9
    det, _ = build_detector(arch, frames, taps, cuda_device, decode_indices=taps)
    res = synthetic.vit_dims(arch)["image_size"]
    x, m = synthetic.make_clips(clips, frames, res, seed=5)
    x8 = torch.randint(0, 256, tuple(x.shape), generator=torch.Generator().manual_seed(6), dtype=torch.uint8)
    assert not m.all()
    for inp in (x.to(cuda_device), x8.to(cuda_device)):
        md = m.to(cuda_device)
        with torch.no_grad():
            monkeypatch.setenv("DFD_OVERLAP", "0")
            assert not det._single_call_ok(False)
            ref_logits, ref_feats = det.predict(inp, md, with_video_features=True)
            monkeypatch.setenv("DFD_OVERLAP", "1")
            assert det._single_call_ok(False)
            runs = [det.predict(inp, md, with_video_features=True) for _ in range(4)]
        torch.cuda.synchronize()
        for logits, feats in runs:
            assert torch.equal(logits[0], ref_logits[0])
            assert torch.equal(feats["video"], ref_feats["video"])
    # with gradients enabled and trainable decoder parameters the call keeps the autograd path
    with torch.enable_grad():
        assert det._single_call_ok(False) == (not any(p.requires_grad for p in det.decoder.parameters()))


@pytest.mark.gpu
def test_full_size_c2_properties(cuda_device):
    """BASELINE config C2 at full size (ViT-B/16, 64 clips x 8 frames): size-independent properties instead of an
    oracle run — clips are independent units (a sub-batch reproduces its rows bit-exactly), reruns are bit-identical,
    masked frames do not matter, logits have norm 5, and the video-level mean of probabilities (inference.py:121,140)
    equals the per-clip softmax averaged by hand."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import video_mean_probs
    det, _ = build_detector("ViT-B/16", 8, [0, 2, 4, 6, 8, 10], cuda_device)
    x, m = synthetic.make_clips(64, 8, 224, seed=7)
    x, m = x.to(cuda_device), m.to(cuda_device)
    assert not m.all()
    full, _ = det.predict(x, m)
    again, _ = det.predict(x, m)
    part, _ = det.predict(x[20:29], m[20:29])
    x2 = x.clone()
    x2[~m] = -7.0
    masked, _ = det.predict(x2, m)
    torch.cuda.synchronize()
    assert torch.isfinite(full[0]).all()
    assert torch.equal(full[0], again[0])
    assert torch.equal(full[0][20:29], part[0])
    assert torch.equal(full[0], masked[0])
    assert torch.allclose(full[0].norm(dim=-1), torch.full((64,), 5.0, device=cuda_device), atol=1e-3)
    counts = [10, 30, 24]
    scores = video_mean_probs(full[0], counts)
    probs = full[0].softmax(-1)
    assert torch.allclose(scores[1], probs[10:40].mean(0), atol=1e-6)
    assert torch.allclose(scores.sum(-1), torch.ones(3, device=cuda_device), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["0", "1"])
def test_parity_with_the_other_layernorm_modes(cuda_device, mode):
    """DFD_LN_FUSE selects how the encoder's LayerNorms run: 2 (default: ln_1 folded into the QKV GEMM's epilogue, its
    producer c_proj emitting bf16(x) and row statistics), 1 (ln_2 folded as well), 0 (separate LayerNorm kernels). The
    other two modes must meet the same golden-vector tolerances, at the small AND the BASELINE sizes. The switch is read
    once per process, so the golden-parity tests are re-run in a child process with it set."""
    import os
    import subprocess
    import sys
    if os.environ.get("DFD_LN_FUSE") is not None:
        pytest.skip("already running with an explicit LayerNorm mode")
    env = dict(os.environ, DFD_LN_FUSE=mode)
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_parity_gpu.py"),
                          os.path.join(here, "test_fullsize_gpu.py"), "-q", "-m", "gpu", "-x", "-k",
                          "golden or from_host or full_size or c2_64 or c4_vitl14 or c5_training"], env=env,
                         capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


@pytest.mark.gpu
def test_video_level_scoring_with_the_real_detector(cuda_device):
    """BASELINE config C3 in miniature: videos of different lengths, clips chunked per predict call, softmax per clip
    and mean over each video's clips (inference.py:107-156) — the driver loop on the CUDA path against the oracle."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import score_videos
    oracle = load_oracle()
    arch, t = "small-512x6", 3
    det, sd = build_detector(arch, t, [0, 2, 4], cuda_device)
    res = synthetic.vit_dims(arch)["image_size"]
    counts = [5, 1, 9, 0, 4]
    x, m = synthetic.make_clips(sum(counts), t, res, seed=23)
    videos, masks, s = [], [], 0
    for n in counts:
        videos.append(x[s:s + n])
        masks.append(m[s:s + n])
        s += n
    got = score_videos(lambda xx, mm: det.predict(xx, mm)[0][0], videos, masks, chunk_clips=4, device=cuda_device)
    with torch.no_grad():
        ref_logits, _ = oracle.detector_predict(sd, x, m, det.layer_indices, (2,))
    ref = oracle.video_scores(ref_logits[0], [c for c in counts if c > 0])
    from dfdclip_b200.inference import score_videos_batched
    batched = score_videos_batched(det, videos, masks, batch_clips=7)  # batches cut across video boundaries
    assert torch.equal(torch.nan_to_num(batched), torch.nan_to_num(got))
    # already-pinned videos take the no-staging path (clip ranges copied straight into the device batch buffer)
    pinned = score_videos_batched(det, [v.pin_memory() for v in videos], [mm.pin_memory() for mm in masks],
                                  batch_clips=7)
    assert torch.equal(torch.nan_to_num(pinned), torch.nan_to_num(got))
    got = got.cpu()
    assert torch.isnan(got[3]).all()  # a video without clips is skipped (inference.py:109-111)
    keep = [i for i, c in enumerate(counts) if c > 0]
    assert (got[keep] - ref).abs().max().item() < 5e-3
    assert torch.equal(got[keep].argmax(-1), ref.argmax(-1))


@pytest.mark.gpu
def test_long_clips_t20(cuda_device):
    """Clips of 20 frames (the reference's T=20 configurations): S = T*P keys per clip through the streaming decoder
    attention and the attn_mode kernels, against the oracle."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    arch, t, b = "tiny-256x4", 20, 3
    det, sd = build_detector(arch, t, [0, 2], cuda_device)
    x, m = synthetic.make_clips(b, t, 32, seed=31)
    m[2, 7:] = False
    with torch.no_grad():
        ref_logits, ref_feat = oracle.detector_predict(sd, x, m, [0, 2], (2,))
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
    torch.cuda.synchronize()
    assert (logits[0].cpu() - ref_logits[0]).abs().max().item() <= TOL_LOGIT_ABS
    assert cosine(feats["video"].cpu(), ref_feat) >= TOL_FEATURE_COSINE


@pytest.mark.gpu
def test_predict_is_cuda_graph_capturable(cuda_device):
    """Every launch of the path goes to the caller's stream and nothing synchronises the device, so a whole
    Detector.predict can be captured in a CUDA graph; replays are bit-identical to the eager call and follow the
    contents of the captured input buffers."""
    from dfdclip_b200 import synthetic
    arch, t, b = "small-512x6", 3, 6
    det, _ = build_detector(arch, t, [0, 2, 4], cuda_device)
    res = synthetic.vit_dims(arch)["image_size"]
    x, m = synthetic.make_clips(b, t, res, seed=3)
    x2, _ = synthetic.make_clips(b, t, res, seed=4)
    xs, m = x.to(cuda_device).clone(), m.to(cuda_device)
    ref1 = det.predict(xs, m)[0][0].clone()          # also warms up (weight packing, function attributes)
    ref2 = det.predict(x2.to(cuda_device), m)[0][0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        det.predict(xs, m)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = det.predict(xs, m)[0][0]
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref1)
    xs.copy_(x2.to(cuda_device))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref2)

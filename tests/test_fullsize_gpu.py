"""GPU parity at the BASELINE.json configuration sizes against golden vectors of the UNMODIFIED reference
(oracle/gen_golden_full.py): C2 (ViT-B/16, 64 clips x 8 frames), the C4 shape (ViT-L/14, 16 frames per clip), the C5
training step (12 clips, loss + every trainable gradient) and a C3 miniature (5 videos of 8..32 clips, per-video mean
of the clip probabilities). The fixtures' task head is centred so that the class margins straddle zero (both classes,
near ties): "identical predicted labels" is tested where it is hard.

Tolerances are BASELINE.json's north_star: per-layer features cosine >= 0.999, clip logits within 2e-2 absolute,
identical predicted labels.
"""
import numpy as np
import pytest
import torch

from helpers import TOL_FEATURE_COSINE, TOL_LOGIT_ABS, cosine, fullsize_inputs, load_golden

pytestmark = pytest.mark.gpu


def build_detector(g, sd, device):
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + g["arch"]
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    det = Detector(cfg, g["num_frames"], None)
    det.load_state_dict(sd, strict=True)
    assert det.layer_indices == g["layer_indices"]
    return det.to(device)


def check_logits(name, got, g):
    """|dlogit| <= 2e-2 on every clip and identical labels; a flipped label is reported with its reference margin."""
    ref = g["logits"]
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    assert err <= TOL_LOGIT_ABS, f"{name}: max |dlogit| {err:.4f}"
    flipped = np.nonzero(got.argmax(-1) != g["pred_labels"])[0]
    report = [(int(i), float(g["margin"][i]), float(got[i, 0] - got[i, 1])) for i in flipped]
    assert len(flipped) == 0, f"{name}: flipped labels (clip, reference margin, measured margin): {report}"
    assert np.allclose(np.linalg.norm(got, axis=-1), 5.0, atol=1e-3)
    return err


def check_taps(name, det, g, x, device, chunk):
    """Sampled K/V of every tapped layer (global flat indices over [N, L, H, 64]) and the tensors' norms."""
    b, t = x.shape[:2]
    seq, h = det.encoder.tokens_per_frame, det.encoder.heads
    per_frame = seq * h * 64
    got = {(layer, key): torch.empty(len(g["idx_%s_%d" % (key, layer)])) for layer in det.layer_indices
           for key in ("k", "v")}
    sq = {k: 0.0 for k in got}
    for i in range(0, b, chunk):
        n = min(b, i + chunk) - i
        qkv, _ = det.encoder.encode(x[i:i + n].flatten(0, 1).to(device), keep_layers=det.layer_indices)
        lo, hi = i * t * per_frame, (i + n) * t * per_frame
        for layer in det.layer_indices:
            view = qkv[layer][:n * t * seq].view(n * t, seq, 3, h, 64)
            for j, key in ((1, "k"), (2, "v")):
                flat = view[:, :, j].float().contiguous().flatten()
                idx = torch.from_numpy(g["idx_%s_%d" % (key, layer)])
                sel = (idx >= lo) & (idx < hi)
                got[(layer, key)][sel] = flat[(idx[sel] - lo).to(device)].cpu()
                sq[(layer, key)] += flat.double().pow(2).sum().item()
        del qkv
    for (layer, key), val in got.items():
        ref = torch.from_numpy(g["val_%s_%d" % (key, layer)])
        c = cosine(val, ref)
        assert c >= TOL_FEATURE_COSINE, f"{name} layer {layer} {key}: cosine {c:.5f}"
        rel = ((val - ref).norm() / ref.norm()).item()
        assert rel < 3e-2, f"{name} layer {layer} {key}: rel err {rel:.4f}"
        assert abs(sq[(layer, key)] ** 0.5 / float(g["norm_%s_%d" % (key, layer)]) - 1) < 5e-3


def test_c2_64_clips_match_the_reference(cuda_device):
    """BASELINE config C2 at full size in ONE predict call (the bench's step) against the reference's chunked run."""
    g = load_golden("vitb16_c2")
    sd, x, m = fullsize_inputs("vitb16_c2", g)
    assert x.shape[0] == 64 and (g["pred_labels"] == 0).any() and (g["pred_labels"] == 1).any()
    assert (np.abs(g["margin"]) < 0.05).any(), "the fixture must contain a near tie"
    det = build_detector(g, sd, cuda_device).eval()
    with torch.no_grad():
        logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
        torch.cuda.synchronize()
        got = logits[0].cpu().numpy()
        check_logits("c2", got, g)
        for b in range(64):   # per clip: a wrong clip must not hide in the batch cosine
            c = cosine(feats["video"][b].cpu(), torch.from_numpy(g["video_feature"][b]))
            assert c >= TOL_FEATURE_COSINE, f"clip {b}: video feature cosine {c:.5f}"
        labels = torch.from_numpy(g["labels"]).to(cuda_device)
        losses, _ = det(x.to(cuda_device), [labels], m.to(cuda_device), single_task=0)
        assert np.abs(losses[0].cpu().numpy() - g["losses"]).max() <= 2 * TOL_LOGIT_ABS
        check_taps("c2", det, g, x, cuda_device, chunk=16)
        # the e2e caller loop (pinned host clips in, host logits out) gives the same bits as predict
        from dfdclip_b200.inference import predict_from_host
        host = predict_from_host(det, x.pin_memory(), m.pin_memory())
        assert torch.equal(host.cpu(), logits[0].cpu())


def test_c4_vitl14_16_frames_match_the_reference(cuda_device):
    g = load_golden("vitl14_c4")
    sd, x, m = fullsize_inputs("vitl14_c4", g)
    assert g["num_frames"] == 16 and len(g["layer_indices"]) == 12
    det = build_detector(g, sd, cuda_device).eval()
    with torch.no_grad():
        logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
        torch.cuda.synchronize()
        check_logits("c4", logits[0].cpu().numpy(), g)
        assert cosine(feats["video"].cpu(), torch.from_numpy(g["video_feature"])) >= TOL_FEATURE_COSINE
        check_taps("c4", det, g, x, cuda_device, chunk=4)


def test_c5_training_step_matches_the_reference(cuda_device):
    """Detector.forward(train=True) + backward of the mean task loss (src/trainer.py:147-165) on 12 clips x 8 frames:
    loss, logits and the gradient of every trainable parameter against the reference's autograd."""
    g = load_golden("vitb16_c5")
    sd, x, m = fullsize_inputs("vitb16_c5", g)
    det = build_detector(g, sd, cuda_device).train()
    labels = torch.from_numpy(g["labels"]).to(cuda_device)
    with torch.enable_grad():
        losses, logits, other = det(x.to(cuda_device), [labels], m.to(cuda_device), comp=["raw"] * x.shape[0],
                                    speed=torch.ones(x.shape[0], device=cuda_device), train=True, single_task=0)
        assert other == {}
        loss = losses[0].mean()
        loss.backward()
    torch.cuda.synchronize()
    check_logits("c5", logits[0].detach().cpu().numpy(), g)
    assert abs(loss.item() - float(g["loss"])) <= TOL_LOGIT_ABS
    params = dict(det.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(n for n, p in params.items() if p.requires_grad)
    worst = (1.0, None)
    for name in names:
        grad = params[name].grad
        assert grad is not None, name
        ref_norm = float(g["gnorm_" + name])
        got_norm = grad.double().norm().item()
        idx = torch.from_numpy(g["gidx_" + name]).to(cuda_device)
        c = cosine(grad.flatten()[idx].cpu(), torch.from_numpy(g["gval_" + name]))
        worst = min(worst, (c, name))
        assert c >= 0.995, f"{name}: gradient cosine {c:.5f}"
        assert abs(got_norm / ref_norm - 1) < 5e-2, f"{name}: gradient norm {got_norm:.4e} vs {ref_norm:.4e}"
    print("worst gradient cosine", worst)


def test_c3_video_level_scores_match_the_reference(cuda_device):
    """5 synthetic videos of U{8..32} clips: every clip through the batched driver, softmax per clip, mean per video
    (inference.py:107-141) against the reference's one-video-at-a-time loop."""
    from dfdclip_b200.inference import score_videos_batched
    g = load_golden("vitb16_c3")
    sd, x, m = fullsize_inputs("vitb16_c3", g)
    counts = [int(c) for c in g["counts"]]
    det = build_detector(g, sd, cuda_device).eval()
    videos, masks, s = [], [], 0
    for n in counts:
        videos.append(x[s:s + n])
        masks.append(m[s:s + n])
        s += n
    with torch.no_grad():
        scores = score_videos_batched(det, videos, masks, batch_clips=64).cpu().numpy()
        logits = det.predict(x[:counts[0]].to(cuda_device), m[:counts[0]].to(cuda_device))[0][0].cpu().numpy()
    assert np.abs(scores - g["video_scores"]).max() < 5e-3
    assert np.array_equal(scores.argmax(-1), g["video_labels"])
    assert np.abs(logits - g["logits"][:counts[0]]).max() <= TOL_LOGIT_ABS


@pytest.mark.parametrize("m,n,k,epi", [
    # ViT-B/16 at C2 (M = 64 * 8 * 197 = 788 * 128): QKV, out-proj, c_fc, c_proj, patch embedding
    (100864, 2304, 768, 0), (100864, 768, 768, 3), (100864, 3072, 768, 1), (100864, 768, 3072, 3),
    (100864, 768, 768, 2),
    # ViT-L/14 at C4 (M = 32 * 16 * 257 = 1028 * 128): QKV, out-proj, c_fc, c_proj, patch embedding (K = 588 -> 640)
    (131584, 3072, 1024, 0), (131584, 1024, 1024, 3), (131584, 4096, 1024, 1), (131584, 1024, 4096, 3),
    (131584, 1024, 640, 2)])
def test_gemm_at_the_baseline_shapes(cuda_device, m, n, k, epi):
    """The tcgen05 GEMM at the full C2 / C4 problem sizes against torch fp32, checked in row
    chunks so that the fp32 reference stays small."""
    import dfdclip_b200._native as nat
    g = torch.Generator(device="cpu").manual_seed(n * 7 + k + epi)
    a = torch.randn(4096, k, generator=g).to(cuda_device, torch.bfloat16).repeat(m // 4096 + 1, 1)[:m].contiguous()
    a[::3] *= 0.5          # rows differ between the repeats of the random block
    a[1::7] *= -1.25
    w = (torch.randn(n, k, generator=g) * (k ** -0.5)).to(cuda_device, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda_device)
    if epi in (nat.EPI_STORE_BF16, nat.EPI_STORE_BF16_QGELU):
        out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=cuda_device)
        tol, res = 6e-3, None
    elif epi == nat.EPI_STORE_F32:
        out = torch.full((m, n), float("nan"), dtype=torch.float32, device=cuda_device)
        tol, res = 2e-5, None
    else:
        res = torch.randn(4096, n, generator=g).to(cuda_device).repeat(m // 4096 + 1, 1)[:m].contiguous()
        out = res.clone()
        tol = 2e-5
    nat.gemm_bf16(a, w, bias, out, epi)
    torch.cuda.synchronize()
    wf = w.float().t().contiguous()
    step = 16384
    for r0 in range(0, m, step):
        ref = a[r0:r0 + step].float() @ wf + bias
        if epi == nat.EPI_STORE_BF16_QGELU:
            ref = ref * torch.sigmoid(1.702 * ref)
        if res is not None:
            ref = ref + res[r0:r0 + step]
        got = out[r0:r0 + step].float()
        assert torch.isfinite(got).all(), f"rows {r0}..: unwritten or non-finite outputs"
        err = ((got - ref).norm() / ref.norm()).item()
        assert err < tol, f"rows {r0}..: rel err {err:.3e}"

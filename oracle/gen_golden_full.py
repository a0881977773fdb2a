"""Golden vectors of the UNMODIFIED reference at the BASELINE.json configuration sizes (run in this container, CPU fp32).

ORACLE / test infrastructure; companion of ``gen_golden.py`` (same stand-ins for yacs / ftfy, same TorchScript
parameter-holder route through the reference's own ``clip.load`` + ``build_model``). Cases:

  vitb16_c2   C2: ViT-B/16, 64 clips x 8 frames through ``Detector.predict`` in chunks of 16 clips, the way the
              reference's caller does it (inference.py:113-118 with scripts/inference.sh's --batch_size 16): logits,
              labels, video features, 4096-sample taps of every tapped layer, eval losses.
  vitl14_c4   C4 shape: ViT-L/14, 16 frames per clip, 4 clips, all 12 taps.
  vitb16_c5   C5: ``Detector.forward(train=True)`` on 12 clips x 8 frames (configs/deepfake/deepfake.yaml:97) +
              ``losses.mean().backward()`` (src/trainer.py:147-165): loss and every trainable parameter's gradient
              (norm + 2048 samples).
  vitb16_c3   C3 miniature: 5 videos of U{8..32} clips (seed 3), chunks of 16, softmax per clip, mean per video
              (inference.py:107-141).

Inputs are ``synthetic.make_varied_clips`` (clips with content, so that the features differ between clips) and the
task head ``decoder.proj0x2`` is re-drawn so that the class margins of the 64 C2 clips straddle zero (both classes,
near ties included): the head's difference direction is made orthogonal to the mean video feature the reference
itself produces. That head is stored in the fixture (``proj0x2``) and loaded by the tests; everything else is the
seeded ``synthetic`` state dict. ``python oracle/gen_golden_full.py [case ...]``.
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as gg  # noqa: E402  (stubs, holder archive, sample indices)
from gen_golden import GOLDEN_DIR, REFERENCE, ROOT, synthetic  # noqa: E402
from reference_runner import build_reference_detector  # noqa: E402

CHUNK = 16          # scripts/inference.sh: --batch_size 16
N_GRAD_SAMPLES = 2048


def predict_chunked(det, x, m, sample_taps=False, seeds=None):
    """model.predict over chunks of CHUNK clips (inference.py:113-118). With sample_taps the encoder is run once more
    per chunk with with_out / with_q to pick the sampled tap values (global flat indices over [N, L, H, dh])."""
    logits, feats = [], []
    b, t = x.shape[:2]
    taps = None
    with torch.no_grad():
        for i in range(0, b, CHUNK):
            lg, ft = det.predict(x[i:i + CHUNK], m[i:i + CHUNK], with_video_features=True)
            logits.append(lg[0])
            feats.append(ft["video"])
            if sample_taps:
                kvs = det.encoder(x[i:i + CHUNK].flatten(0, 1))
                if taps is None:
                    taps = {}
                    for layer in det.layer_indices:
                        for key in ("k", "v"):
                            numel = b * t * kvs[layer][key][0].numel()
                            idx = gg.sample_indices(numel, seed=seeds[(layer, key)])
                            taps[(layer, key)] = (idx, torch.empty(idx.numel()), torch.zeros((), dtype=torch.float64))
                per_chunk = (min(b, i + CHUNK) - i) * t * kvs[0]["k"][0].numel()
                lo = i * t * kvs[0]["k"][0].numel()
                for layer in det.layer_indices:
                    for key in ("k", "v"):
                        idx, val, sq = taps[(layer, key)]
                        flat = kvs[layer][key].contiguous().flatten()
                        sel = (idx >= lo) & (idx < lo + per_chunk)
                        val[sel] = flat[idx[sel] - lo]
                        sq += flat.double().pow(2).sum()
                del kvs
    return torch.cat(logits), torch.cat(feats), taps


def centred_head(features, width, seed=0):
    """A [D, 2] task head whose class margin straddles zero on `features`: columns c + d and c - d with
      * the common direction c along the mean feature mu (c.f ~ 1 for every clip: the raw logits keep a healthy norm,
        as a trained head's do. With a random c the common part changes sign from clip to clip and 5 l / |l|
        (src/models.py:551-553) divides by almost nothing there: the normalised logits of such clips are
        ill-conditioned for ANY finite-precision evaluation, the reference's own included), and
      * the difference direction d orthogonal to mu, scaled so that the margins of the normalised logits have a
        standard deviation of about 1: both classes occur and near ties (|margin| < 0.05) are likely among 64 clips."""
    g = torch.Generator().manual_seed(1000 + seed)
    f = features.double()
    mu = f.mean(0)
    c = mu / (mu @ mu)
    d = torch.randn(width, generator=g, dtype=torch.float64) * width ** -0.5
    d = d - (d @ mu) / (mu @ mu) * mu
    common = (f @ c).abs().mean()
    spread = (f @ d).std()
    d = d * (common / spread) * (1.0 / 5.0) * 2 ** -0.5     # margin ~ 5 sqrt(2) d.f / c.f ~ N(0, 1)
    return torch.stack([c + d, c - d], dim=1).float()


def common_fields(det, arch, num_frames, x, m, labels, logits, feats):
    return {
        "arch": np.array(arch), "num_frames": np.array(num_frames), "batch": np.array(x.shape[0]),
        "layer_indices": np.array(det.layer_indices), "mask": m.numpy(), "labels": labels.numpy(),
        "logits": logits.numpy(), "video_feature": feats.numpy(), "pred_labels": logits.argmax(-1).numpy(),
        "margin": (logits[:, 0] - logits[:, 1]).numpy(), "proj0x2": det.decoder.proj0x2.detach().numpy().copy(),
        "clips": np.array("make_varied_clips"), "chunk": np.array(CHUNK),
    }


def save(name, out):
    path = os.path.join(GOLDEN_DIR, "reference_%s.npz" % name)
    np.savez_compressed(path, **out)
    mg = np.abs(out["margin"]) if "margin" in out else np.zeros(1)
    print("%-10s labels %s, |margin| min %.4f, <0.05: %d, <0.2: %d -> %s (%.1f KiB)" % (
        name, np.bincount(out["pred_labels"], minlength=2).tolist(), mg.min(), (mg < 0.05).sum(), (mg < 0.2).sum(),
        os.path.relpath(path, ROOT), os.path.getsize(path) / 1024), flush=True)


def tap_fields(out, taps):
    for (layer, key), (idx, val, sq) in taps.items():
        out["idx_%s_%d" % (key, layer)] = idx.numpy()
        out["val_%s_%d" % (key, layer)] = val.numpy()
        out["norm_%s_%d" % (key, layer)] = np.array(float(sq.sqrt()), dtype=np.float64)


def case_c2():
    arch, t, b = "ViT-B/16", 8, 64
    x, m = synthetic.make_varied_clips(b, t, 224, seed=7)
    det = build_reference_detector(arch, t)
    t0 = time.time()
    prev = os.path.join(GOLDEN_DIR, "reference_vitb16_c2.npz")
    if os.path.exists(prev) and os.environ.get("DFD_GOLDEN_REUSE_FEATURES", "1") == "1":
        feats = torch.from_numpy(np.load(prev)["video_feature"])  # the features do not depend on the head
    else:
        _, feats, _ = predict_chunked(det, x, m)
    print("c2 pass 1 (features for the head) %.0f s" % (time.time() - t0), flush=True)
    proj = centred_head(feats, 768)
    det.decoder.proj0x2.data.copy_(proj)
    seeds = {(layer, key): 7000 + 10 * layer + (key == "v") for layer in det.layer_indices for key in ("k", "v")}
    logits, feats, taps = predict_chunked(det, x, m, sample_taps=True, seeds=seeds)
    labels = torch.arange(b) % 2
    out = common_fields(det, arch, t, x, m, labels, logits, feats)
    # the eval forward's task loss is the per-sample cross entropy of these logits (src/models.py:38-44, 590-593)
    out["losses"] = torch.nn.functional.cross_entropy(logits, labels, reduction="none").numpy()
    tap_fields(out, taps)
    save("vitb16_c2", out)
    return proj


def load_c2_head():
    path = os.path.join(GOLDEN_DIR, "reference_vitb16_c2.npz")
    return torch.from_numpy(np.load(path)["proj0x2"])


def case_c4():
    arch, t, b = "ViT-L/14", 16, 4
    x, m = synthetic.make_varied_clips(b, t, 224, seed=7)
    det = build_reference_detector(arch, t)
    # the head is centred on a wider sample of clips than the 4 stored ones (features only)
    xs, ms = synthetic.make_varied_clips(12, t, 224, seed=8, masked_tail=False)
    _, feats, _ = predict_chunked(det, xs, ms)
    del xs
    proj = centred_head(feats, 1024, seed=4)
    det.decoder.proj0x2.data.copy_(proj)
    seeds = {(layer, key): 9000 + 10 * layer + (key == "v") for layer in det.layer_indices for key in ("k", "v")}
    logits, feats, taps = predict_chunked(det, x, m, sample_taps=True, seeds=seeds)
    labels = torch.arange(b) % 2
    out = common_fields(det, arch, t, x, m, labels, logits, feats)
    out["losses"] = torch.nn.functional.cross_entropy(logits, labels, reduction="none").numpy()
    tap_fields(out, taps)
    save("vitl14_c4", out)


def case_c5():
    arch, t, b = "ViT-B/16", 8, 12
    x, m = synthetic.make_varied_clips(b, t, 224, seed=17)
    det = build_reference_detector(arch, t, proj=load_c2_head())
    det.train()
    labels = torch.arange(b) % 2
    comp = ["raw"] * b
    speed = torch.ones(b)
    losses, logits, other = det(x, [labels], m, comp=comp, speed=speed, train=True, single_task=0)
    assert other == {}
    loss = losses[0].mean()
    loss.backward()       # src/trainer.py:162 (accelerator.backward of the mean task loss)
    out = common_fields(det, arch, t, x, m, labels, logits[0].detach(), torch.zeros(0))
    out["losses"] = losses[0].detach().numpy()
    out["loss"] = np.array(loss.item())
    names = []
    for name, p in det.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        names.append(name)
        gflat = p.grad.flatten()
        out["gnorm_" + name] = np.array(gflat.double().norm().item())
        idx = torch.randint(0, gflat.numel(), (min(N_GRAD_SAMPLES, gflat.numel()),),
                            generator=torch.Generator().manual_seed(len(names)))
        out["gidx_" + name] = idx.numpy()
        out["gval_" + name] = gflat[idx].numpy()
    out["grad_names"] = np.array(names)
    save("vitb16_c5", out)


def case_c3():
    arch, t = "ViT-B/16", 8
    counts = torch.randint(8, 33, (5,), generator=torch.Generator().manual_seed(3)).tolist()
    video_ids = [v for v, n in enumerate(counts) for _ in range(n)]
    x, m = synthetic.make_varied_clips(len(video_ids), t, 224, seed=27, video_ids=video_ids)
    det = build_reference_detector(arch, t, proj=load_c2_head())
    logits, scores, s = [], [], 0
    with torch.no_grad():
        for n in counts:                                            # inference.py:107-141, one video at a time
            lg = torch.cat([det.predict(x[s + i:s + min(n, i + CHUNK)], m[s + i:s + min(n, i + CHUNK)])[0][0]
                            for i in range(0, n, CHUNK)])
            logits.append(lg)
            scores.append(lg.softmax(-1).mean(0))
            s += n
    logits = torch.cat(logits)
    labels = torch.tensor([v % 2 for v in video_ids])
    out = common_fields(det, arch, t, x, m, labels, logits, torch.zeros(0))
    out["counts"] = np.array(counts)
    out["video_ids"] = np.array(video_ids)
    out["video_scores"] = torch.stack(scores).numpy()
    out["video_labels"] = torch.stack(scores).argmax(-1).numpy()
    save("vitb16_c3", out)


CASES = {"vitb16_c2": case_c2, "vitl14_c4": case_c4, "vitb16_c5": case_c5, "vitb16_c3": case_c3}


def main(argv):
    warnings.filterwarnings("ignore")
    gg.install_stubs()
    sys.path.insert(0, REFERENCE)
    torch.set_num_threads(os.cpu_count() or 1)
    for name, fn in CASES.items():
        if len(argv) > 1 and name not in argv[1:]:
            continue
        t0 = time.time()
        fn()
        print("%s done in %.0f s" % (name, time.time() - t0), flush=True)


if __name__ == "__main__":
    main(sys.argv)

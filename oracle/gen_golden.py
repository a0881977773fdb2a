"""Generate tests/golden/*.npz by running the UNMODIFIED reference (``/root/reference``) in this container.

ORACLE / test infrastructure. The reference is pure Python and cannot travel to the GPU box, so its outputs on
seeded synthetic inputs are committed as small fixtures, together with this script (``python oracle/gen_golden.py``).

How the reference is run without modification:
  * ``yacs`` and ``ftfy`` are not installed: tiny stand-ins are injected into ``sys.modules`` (a dict-with-attributes
    ``CfgNode``; ``ftfy.fix_text`` = identity — the tokenizer is imported by ``src/clip`` but never used);
  * ``Detector.__init__`` calls ``clip.load(config.architecture)``: a TorchScript archive of a parameter-holder
    module whose ``state_dict()`` has CLIP's keys is written to a temp dir, so the reference's own loader and
    ``build_model`` (fp16 rounding included) build the encoder;
  * the full ``Detector`` state dict from ``dfdclip_b200.synthetic`` is then loaded with the reference's strict
    ``load_state_dict`` — the route ``inference.py:99`` takes with ``*_weights.pt``.
"""
import contextlib
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("DFD_REFERENCE", "/root/reference")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from dfdclip_b200 import synthetic  # noqa: E402  (weights/clips factory shared with tests and bench)

# name -> (arch, T, B, full K/V stored?[, adapter.struct.type])
CASES = {
    "tiny": ("tiny-256x4", 4, 3, True),
    "small": ("small-512x6", 3, 2, False),
    "vitb16": ("ViT-B/16", 8, 2, False),
    "vitl14": ("ViT-L/14", 4, 1, False),
    # CompInvAdapter (src/models.py:783-940), one case per Sequential layout
    "tiny_ad_x": ("tiny-256x4", 4, 3, True, "768-x-768"),
    "tiny_ad_legacy": ("tiny-256x4", 4, 3, True, "legacy-768-x-768"),
    "tiny_ad_nln": ("tiny-256x4", 4, 3, True, "768-x-768-nln"),
    "tiny_ad_ln": ("tiny-256x4", 4, 3, True, "768-x-768-ln"),
    "tiny_ad_z0": ("tiny-256x4", 4, 3, True, "768-x-768-z0"),
    "tiny_ad_xxx": ("tiny-256x4", 4, 3, True, "768-xxx-768"),
    "tiny_ad_linear": ("tiny-256x4", 4, 3, True, "linear"),
    "vitb16_ad_nln": ("ViT-B/16", 8, 2, False, "768-x-768-nln"),
    "vitb16_ad_z0": ("ViT-B/16", 8, 2, False, "768-x-768-z0"),
    "vitb16_ad_bn": ("ViT-B/16", 8, 2, False, "768-bn"),  # the reference hard-codes Linear(768, 768): width 768 only
    # non-default decoder modes (src/models.py:107-115, 250-267, 345-357, 511-544, 572-578)
    "tiny_aug_query": ("tiny-256x4", 4, 3, True, None, {"aug_query": 1}),
    "tiny_global_pred": ("tiny-256x4", 4, 3, True, None, {"global_prediction": 1}),
    "small_gp_aq": ("small-512x6", 3, 2, False, None, {"global_prediction": 1, "aug_query": 1}),
    "tiny_ema": ("tiny-256x4", 4, 3, True, None, {"ema_frame": 0.3, "temporal_position": 0}),
    "tiny_no_tpos": ("tiny-256x4", 4, 3, True, None, {"temporal_position": 0}),
    "tiny_attn_frame": ("tiny-256x4", 4, 3, True, None, {"attn_mode": "frame"}),
    "tiny_attn_tf": ("tiny-256x4", 4, 3, True, None, {"attn_mode": "temporal+frame"}),
    "small_attn_temporal": ("small-512x6", 3, 2, False, None, {"attn_mode": "temporal"}),
    # trainer-side auxiliary losses of Detector.forward(train=True) (src/models.py:598-738)
    "small_tm_ranking": ("small-512x6", 3, 4, False, None, None, None, {"temporal": "ranking"}),
    "small_tm_triplet": ("small-512x6", 3, 4, False, None, None, None, {"temporal": "triplet"}),
    "small_pm_batch": ("small-512x6", 3, 2, False, None, None, {"type": "batch", "ratio": 0.5}),
    "small_pm_sample": ("small-512x6", 3, 2, False, None, None, {"type": "sample", "ratio": 0.5}),
    "tiny_pm_adapter": ("tiny-256x4", 4, 3, True, "768-x-768-z0", None, {"type": "batch", "ratio": 0.5}),
}
N_SAMPLES = 4096


def install_stubs():
    class CfgNode(dict):
        def __init__(self, init_dict=None, key_list=None, new_allowed=False):
            super().__init__(init_dict or {})

        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name)

        def __setattr__(self, name, value):
            self[name] = value

    yacs = types.ModuleType("yacs")
    yacs_config = types.ModuleType("yacs.config")
    yacs_config.CfgNode = CfgNode
    yacs.config = yacs_config
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_config)
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda text: text
    sys.modules.setdefault("ftfy", ftfy)


class FakeAccelerator:
    device = torch.device("cpu")

    @contextlib.contextmanager
    def main_process_first(self):
        yield


def write_jit_holder(state_dict, path):
    """TorchScript archive of empty modules carrying `state_dict` as buffers under the same dotted names."""
    root = torch.nn.Module()
    for key, tensor in state_dict.items():
        mod = root
        parts = key.split(".")
        for part in parts[:-1]:
            if not hasattr(mod, part):
                mod.add_module(part, torch.nn.Module())
            mod = getattr(mod, part)
        mod.register_buffer(parts[-1], tensor.clone())
    torch.jit.script(root).save(path)


def sample_indices(numel, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(N_SAMPLES, numel),), generator=g)


def run_case(name, arch, num_frames, batch, full, adapter=None, op_mode=None, patch_mask=None, train_mode=None):
    from src.models import Detector  # the reference's own class

    dims = synthetic.vit_dims(arch)
    tmp = tempfile.mkdtemp(prefix="dfd_golden_")
    ckpt = os.path.join(tmp, "clip_%s.pt" % name)
    write_jit_holder(synthetic.clip_checkpoint_state_dict(arch, seed=0), ckpt)

    cfg = Detector.get_default_config()
    cfg.architecture = ckpt
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    from yacs.config import CfgNode
    for key, val in (op_mode or {}).items():
        cfg.op_mode[key] = val
    if patch_mask is not None:
        cfg.train_mode["patch_mask"] = CfgNode(dict(patch_mask))
    for key, val in (train_mode or {}).items():
        cfg.train_mode[key] = val
    if adapter is not None:
        cfg.adapter.type = "normal"
        cfg.adapter.frozen = 0
        cfg.adapter.struct = CfgNode({"type": adapter, "x": 256})
    torch.manual_seed(1)
    det = Detector(cfg, num_frames, FakeAccelerator())
    os.remove(ckpt)

    op = dict(op_mode or {})
    sd = synthetic.detector_state_dict(arch, num_frames, out_dims=(2,), taps=det.layer_indices, seed=0,
                                       adapter=adapter, adapter_inner=256, aug_query=bool(op.get("aug_query")),
                                       global_prediction=bool(op.get("global_prediction")),
                                       temporal_position=bool(op.get("temporal_position", 1)),
                                       ranking=(train_mode or {}).get("temporal") == "ranking")
    # the encoder built by the reference's clip.load/build_model must equal the synthetic fp32 values exactly
    # (they are fp16-representable where build_model rounds)
    for k, v in det.encoder.state_dict().items():
        assert torch.equal(v, sd["encoder." + k]), "encoder weight mismatch after the reference loader: " + k
    missing = det.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    det.eval()

    x, m = synthetic.make_clips(batch, num_frames, dims["image_size"], seed=7)
    labels = torch.arange(batch) % 2
    extra = {}
    with torch.no_grad():
        taps = det.encoder(x.flatten(0, 1), with_out=True, with_q=True)
        if patch_mask is not None:
            # train_mode.patch_mask only acts with train=True (:511); the module stays in eval mode (dropout is 0)
            np.random.seed(1234)
            logits, feats = det.predict(x, m, with_video_features=True, with_adapt_features=adapter is not None,
                                        train=True)
            np.random.seed(1234)
            losses, logits2, _ = det(x, [labels], m, single_task=0, train=True)
            np.random.seed(1234)   # the index draws, for the oracle (same numpy calls as :511-541)
            num_patch = (dims["image_size"] // dims["patch_size"]) ** 2
            num_select = int(num_patch * patch_mask["ratio"])
            drawn, idx = [], None
            for _ in det.layer_indices:
                if patch_mask["type"] == "sample" or idx is None:
                    idx = np.random.choice(range(num_patch), num_select, replace=False)
                drawn.append(idx)
            extra["patch_indices"] = np.stack(drawn)
            extra["patch_mask_type"] = np.array(patch_mask["type"])
            extra["patch_mask_ratio"] = np.array(patch_mask["ratio"])
        elif train_mode:
            # Detector.forward(train=True): task losses + the auxiliary losses; the module stays in eval mode
            # (dropout is 0) and the sample pairs alternate raw / c23 like the compression sampler produces them
            import random
            comp = ["raw" if i % 2 == 0 else "c23" for i in range(batch)]
            speed = torch.linspace(0.7, 1.3, batch)[torch.randperm(batch, generator=torch.Generator().manual_seed(5))]
            random.seed(11)
            losses, logits, other = det(x, [labels], m, comp=comp, speed=speed, train=True, single_task=0)
            logits2, feats = logits, {"video": np.zeros((0,), dtype=np.float32)}
            extra["speed"] = speed.numpy()
            extra["comp_raw_first"] = np.array(True)
            extra["train_mode"] = np.array(repr(sorted(train_mode.items())))
            for key, val in other.items():
                extra["other_" + key.replace("/", "_")] = np.array(float(val))
        elif op.get("ema_frame"):
            # ema_frame acts in Detector.forward, not in predict (:572-578)
            losses, logits = det(x, [labels], m, single_task=0)
            logits2, feats = logits, {"video": np.zeros((0,), dtype=np.float32)}
        else:
            logits, feats = det.predict(x, m, with_video_features=True, with_adapt_features=adapter is not None)
            losses, logits2 = det(x, [labels], m, single_task=0)
    assert torch.allclose(logits[0], logits2[0], rtol=0, atol=0, equal_nan=True)

    out = {
        "arch": np.array(arch), "num_frames": np.array(num_frames), "batch": np.array(batch),
        "layer_indices": np.array(det.layer_indices), "mask": m.numpy(), "labels": labels.numpy(),
        "logits": logits[0].numpy(), "video_feature": np.asarray(feats["video"]), "losses": losses[0].numpy(),
        "pred_labels": logits[0].argmax(-1).numpy(),
    }
    out.update(extra)
    if op_mode:
        out["op_mode"] = np.array(repr(sorted(op.items())))
    if adapter is not None:
        out["adapter"] = np.array(adapter)
    if adapter is not None and not train_mode:
        # adapted taps as the decoder received them: predict's kvs after the adapter AND after Decoder.forward's
        # in-place list edits (positional embedding added, (t, p) flattened — src/models.py:326-334)
        for i, kv in enumerate(feats["adapt"]):
            for key in ("k", "v"):
                t = kv[key].contiguous().float()
                if det.decoder.positional_embedding is not None:
                    t = (t.view(batch, num_frames, -1, *t.shape[-2:]) - det.decoder.positional_embedding).contiguous()
                out["norm_adapt_%s_%d" % (key, i)] = np.array(t.norm().item(), dtype=np.float64)
                if full:
                    out["adapt_%s_%d" % (key, i)] = t.detach().numpy()
                else:
                    idx = sample_indices(t.numel(), seed=5000 + 10 * i + (key == "v"))
                    out["idx_adapt_%s_%d" % (key, i)] = idx.numpy()
                    out["val_adapt_%s_%d" % (key, i)] = t.flatten()[idx].detach().numpy()
    for i, a in enumerate(taps):
        if adapter is not None or op_mode or patch_mask or train_mode:
            break  # the raw taps are pinned by the default-mode cases
        for key in ("q", "k", "v", "out"):
            t = a[key].contiguous().float()
            out["norm_%s_%d" % (key, i)] = np.array(t.norm().item(), dtype=np.float64)
            if full:
                out["%s_%d" % (key, i)] = t.numpy()
            else:
                idx = sample_indices(t.numel(), seed=1000 * i + {"q": 0, "k": 1, "v": 2, "out": 3}[key])
                out["idx_%s_%d" % (key, i)] = idx.numpy()
                out["val_%s_%d" % (key, i)] = t.flatten()[idx].numpy()
    path = os.path.join(GOLDEN_DIR, "reference_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-8s %-12s logits=%s labels=%s -> %s (%.1f KiB)" % (
        name, arch, np.round(out["logits"], 4).tolist(), out["pred_labels"].tolist(), os.path.relpath(path, ROOT),
        os.path.getsize(path) / 1024))


def main(argv):
    warnings.filterwarnings("ignore")
    install_stubs()
    sys.path.insert(0, REFERENCE)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name, case in CASES.items():
        if len(argv) > 1 and name not in argv[1:]:
            continue
        run_case(name, *case)


if __name__ == "__main__":
    main(sys.argv)

"""GPU parity of the CompInvAdapter path (reference src/models.py:783-940, called at :546-547): the native in-place
adapter (`dfd_adapter_apply`: tcgen05 GEMM -> row kernel -> tcgen05 GEMM with bf16 reduce-add) between the encoder
taps and the decoder, against golden vectors of the unmodified reference and against the CPU oracle.

Tolerances are the north_star's: features cosine >= 0.999, clip logits within 2e-2, identical labels."""
import numpy as np
import pytest
import torch

from helpers import (TOL_FEATURE_COSINE, TOL_LOGIT_ABS, cosine, golden_inputs, golden_tensor, load_golden,
                     load_oracle)

pytestmark = pytest.mark.gpu

STRUCTS = ["768-x-768", "legacy-768-x-768", "768-x-768-nln", "768-x-768-ln", "768-x-768-z0", "768-xxx-768", "linear"]


def build_adapter_detector(arch, num_frames, struct, device, sd=None, inner=256, decode_indices=None, frozen=0):
    from dfdclip_b200 import synthetic
    from dfdclip_b200.config import CN
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    cfg.adapter.type = "normal"
    cfg.adapter.frozen = frozen
    cfg.adapter.struct = CN({"type": struct, "x": inner})
    if decode_indices is not None:
        cfg.decode_mode = "index"
        cfg.decode_indices = list(decode_indices)
    det = Detector(cfg, num_frames, None)
    if sd is None:
        sd = synthetic.detector_state_dict(arch, num_frames, out_dims=(2,), taps=det.layer_indices, seed=0,
                                           adapter=struct, adapter_inner=inner)
    det.load_state_dict(sd, strict=True)
    return det.to(device).eval(), sd


def check_labels(got, ref, margin=4 * TOL_LOGIT_ABS):
    """Identical labels, except where the reference's own class margin is inside the logit tolerance."""
    got_l, ref_l = got.argmax(-1), ref.argmax(-1)
    gap = np.abs(ref[:, 0] - ref[:, 1])
    assert np.array_equal(got_l[gap > margin], ref_l[gap > margin])


@pytest.mark.parametrize("case", ["tiny_ad_x", "tiny_ad_legacy", "tiny_ad_nln", "tiny_ad_ln", "tiny_ad_z0",
                                  "tiny_ad_xxx", "tiny_ad_linear", "vitb16_ad_nln", "vitb16_ad_z0", "vitb16_ad_bn"])
def test_adapter_predict_matches_reference_golden(cuda_device, case):
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    det, _ = build_adapter_detector(g["arch"], g["num_frames"], str(g["adapter"]), cuda_device, sd)
    assert det.layer_indices == g["layer_indices"]
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True,
                                with_adapt_features=True)
    torch.cuda.synchronize()
    got = logits[0].cpu().numpy()
    err = np.abs(got - g["logits"]).max()
    assert err <= TOL_LOGIT_ABS, f"{case}: max |dlogit| {err:.4f}"
    check_labels(got, g["logits"])
    assert cosine(feats["video"].cpu(), torch.from_numpy(g["video_feature"])) >= TOL_FEATURE_COSINE
    assert len(feats["adapt"]) == len(g["layer_indices"])
    for i, kv in enumerate(feats["adapt"]):
        assert tuple(kv["k"].shape[:2]) == (g["batch"], g["num_frames"]) and kv["k"].dtype == torch.float32
        for key in ("k", "v"):
            ref, val = golden_tensor(g, "adapt_" + key, i, kv[key])
            c = cosine(val, ref)
            assert c >= TOL_FEATURE_COSINE, f"{case} tap {i} {key}: cosine {c:.5f}"
            assert ((val - ref).norm() / ref.norm()).item() < 3e-2


@pytest.mark.parametrize("struct", STRUCTS)
@pytest.mark.parametrize("inner", [256, 1024])
def test_adapter_apply_unit_vitb_shapes(cuda_device, struct, inner):
    """dfd_adapter_apply on its own at ViT-B/16 tap shapes (D=768, 197 tokens/frame, K column block of a packed
    [rows, 3D] buffer) against the oracle fed the same bf16 tap values; the Q and V blocks must stay untouched."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import CompInvAdapter
    oracle = load_oracle()
    if struct == "linear" and inner != 256:
        pytest.skip("the linear struct has no inner width")
    arch, frames, seq, d, h = "ViT-B/16", 5, 197, 768, 12
    sd = synthetic.adapter_state_dict(arch, 1, struct, inner, seed=3)
    g = torch.Generator().manual_seed(5)
    qkv = (torch.randn(frames * seq, 3 * d, generator=g) * 1.5).to(torch.bfloat16)

    class _Enc:
        width, input_resolution, patch_size = d, 224, 16

    class _Det:
        encoder, layer_indices = _Enc, [0]

    from dfdclip_b200.config import CN
    cfg = CN({"adapter": {"struct": {"type": struct, "x": inner}}, "dropout": 0.0})
    ad = CompInvAdapter(cfg, _Det).to(cuda_device).eval()
    ad.load_state_dict(sd, strict=True)
    buf = qkv.to(cuda_device)
    before = buf.clone()
    ad.apply_packed({0: buf}, [0], frames, seq)
    torch.cuda.synchronize()
    view = qkv.view(1, frames, seq, 3, h, 64)
    kvs = [dict(k=view[:, :, 1:, 1].float(), v=view[:, :, 1:, 2].float())]
    with torch.no_grad():
        ref = oracle.adapter_forward({"adapter." + k: v for k, v in sd.items()}, kvs, struct)
    out = buf.cpu().view(1, frames, seq, 3, h, 64)
    assert torch.equal(out[:, :, :, 0], before.cpu().view(1, frames, seq, 3, h, 64)[:, :, :, 0])  # Q untouched
    for j, key in ((1, "k"), (2, "v")):
        got = out[:, :, 1:, j].float()
        c = cosine(got, ref[0][key])
        assert c >= 0.9995, (struct, key, c)
        assert ((got - ref[0][key]).norm() / ref[0][key].norm()).item() < 1.5e-2
    assert torch.isfinite(out.float()).all()


def test_adapter_forward_generic_kvs(cuda_device):
    """CompInvAdapter.forward(kvs) with the reference's signature on arbitrary [B,T,P,H,dh] tensors: returns new
    tensors, leaves the inputs untouched."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    det, sd = build_adapter_detector("small-512x6", 3, "768-x-768-nln", cuda_device)
    b, t, p, h = 2, 3, 16, 8
    g = torch.Generator().manual_seed(9)
    kvs_cpu = [dict(k=torch.randn(b, t, p, h, 64, generator=g), v=torch.randn(b, t, p, h, 64, generator=g))
               for _ in det.layer_indices]
    kvs = [{n: kv[n].to(cuda_device).to(torch.bfloat16) for n in kv} for kv in kvs_cpu]
    keep = [{n: kv[n].clone() for n in kv} for kv in kvs]
    out = det.adapter([dict(kv) for kv in kvs])
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = oracle.adapter_forward(sd, [{n: kv[n].to(torch.bfloat16).float() for n in kv} for kv in kvs_cpu],
                                     "768-x-768-nln")
    for i in range(len(kvs)):
        for n in ("k", "v"):
            assert torch.equal(kvs[i][n], keep[i][n])
            assert tuple(out[i][n].shape) == (b, t, p, h, 64)
            assert cosine(out[i][n].float().cpu(), ref[i][n]) >= 0.9995


def test_adapter_against_oracle_with_masks_and_host_pipeline(cuda_device):
    """Adapter + index taps + masked frames against the oracle, and the chunked host pipeline (adapter applied once
    on the whole-batch tap buffers) bit-identical to one predict() call."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import HostClipPipeline
    oracle = load_oracle()
    arch, t, b, taps = "small-512x6", 3, 9, [2, 3, 5]
    det, sd = build_adapter_detector(arch, t, "768-x-768-z0", cuda_device, decode_indices=taps)
    x, m = synthetic.make_clips(b, t, synthetic.vit_dims(arch)["image_size"], seed=13)
    m[3, 1:] = False
    with torch.no_grad():
        ref_logits, ref_feat = oracle.detector_predict(sd, x, m, taps, (2,), adapter="768-x-768-z0")
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
    torch.cuda.synchronize()
    got, ref = logits[0].cpu().numpy(), ref_logits[0].numpy()
    assert np.abs(got - ref).max() <= TOL_LOGIT_ABS
    check_labels(got, ref)
    assert cosine(feats["video"].cpu(), ref_feat) >= TOL_FEATURE_COSINE
    again, _ = det.predict(x.to(cuda_device), m.to(cuda_device))
    assert torch.equal(again[0], logits[0])  # deterministic; the taps of a call are adapted exactly once
    pipe = HostClipPipeline(det, chunk_clips=4)
    assert torch.equal(pipe(x.pin_memory(), m.pin_memory()), logits[0].cpu())
    # the cross-batch stream (adapter applied per batch on that batch's own tap buffers, decoder on its own stream)
    from dfdclip_b200.inference import HostClipStream
    xp, mp = x.pin_memory(), m.pin_memory()
    for overlap in (True, False):
        stream = HostClipStream(det, overlap_decoder=overlap)
        got = torch.cat(list(stream.run([(xp[:4], mp[:4]), (xp[4:], mp[4:]), (xp[:4], mp[:4])])))
        want = torch.cat([det.predict(xp[a:b_].to(cuda_device), mp[a:b_].to(cuda_device))[0][0].cpu()
                          for a, b_ in ((0, 4), (4, 9), (0, 4))])
        assert torch.equal(got, want)


def test_frozen_adapter_trains_decoder_only(cuda_device):
    """adapter.frozen = 1 of the reference (:479-480): the native in-place adapter runs inside the training step and
    only the decoder receives gradients. (The trainable adapter is covered by tests/test_train_gpu.py.)"""
    from dfdclip_b200 import synthetic
    arch, t, b = "tiny-256x4", 4, 3
    det, _ = build_adapter_detector(arch, t, "768-x-768-ln", cuda_device)
    x, m = synthetic.make_clips(b, t, 32, seed=1)
    x, m = x.to(cuda_device), m.to(cuda_device)
    y = torch.tensor([0, 1, 0], device=cuda_device)
    det.train()
    for p in det.adapter.parameters():
        p.requires_grad = False
    assert not det.adapter.needs_autograd()
    with torch.enable_grad():
        losses, logits, other = det(x, [y], m, train=True)
        losses[0].mean().backward()
    assert other == {}
    assert det.decoder.class_embedding.grad is not None and torch.isfinite(det.decoder.class_embedding.grad).all()
    assert all(p.grad is None for p in det.adapter.parameters())
    with torch.enable_grad():
        for p in det.adapter.parameters():
            p.requires_grad = True
        with pytest.raises(NotImplementedError):  # the in-place path refuses a trainable adapter under autograd
            det.adapter.apply_packed({}, [], 0, 5)


def test_shipped_config_shape_against_oracle(cuda_device):
    """The shape of the reference's shipped configs (configs/deepfake/*.yaml): ViT-B/16, 20 frames per clip,
    decode_mode index with taps 6..11, `768-x-768-nln` adapter (x = 256) — the CUDA path against the oracle."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    arch, t, b, taps = "ViT-B/16", 20, 2, [6, 7, 8, 9, 10, 11]
    det, sd = build_adapter_detector(arch, t, "768-x-768-nln", cuda_device, decode_indices=taps)
    x, m = synthetic.make_clips(b, t, 224, seed=41)
    m[1, 15:] = False
    with torch.no_grad():
        ref_logits, ref_feat = oracle.detector_predict(sd, x, m, taps, (2,), adapter="768-x-768-nln")
    logits, feats = det.predict(x.to(cuda_device), m.to(cuda_device), with_video_features=True)
    torch.cuda.synchronize()
    got, ref = logits[0].cpu().numpy(), ref_logits[0].numpy()
    assert np.abs(got - ref).max() <= TOL_LOGIT_ABS, np.abs(got - ref).max()
    check_labels(got, ref)
    assert cosine(feats["video"].cpu(), ref_feat) >= TOL_FEATURE_COSINE


def test_bn_adapter_eval_is_native_and_train_mode_uses_batch_statistics(cuda_device):
    """adapter.struct.type = "768-bn" (reference :877-887): in eval mode the native path applies the running statistics
    as one affine per frame index (checked against the oracle at a width the reference's hard-coded Linear(768, 768)
    cannot take); in train mode BatchNorm normalises with batch statistics, so the call must leave the native in-place
    path for the torch modules, and its running statistics move."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    arch, t, b = "small-512x6", 3, 4
    det, sd = build_adapter_detector(arch, t, "768-bn", cuda_device)
    assert any(k.endswith("l0_k.1.running_var") for k in det.state_dict())
    x, m = synthetic.make_clips(b, t, synthetic.vit_dims(arch)["image_size"], seed=17)
    with torch.no_grad():
        ref_logits, _ = oracle.detector_predict(sd, x, m, det.layer_indices, (2,), adapter="768-bn")
    assert not det.adapter.needs_autograd()
    logits, _ = det.predict(x.to(cuda_device), m.to(cuda_device))
    got, ref = logits[0].cpu().numpy(), ref_logits[0].numpy()
    assert np.abs(got - ref).max() <= TOL_LOGIT_ABS
    check_labels(got, ref)
    with pytest.raises(ValueError):  # a batch whose frame count is not a multiple of the BatchNorm's channels
        det.adapter.apply_packed(det.encoder.encode(x[:1, :2].flatten(0, 1).to(cuda_device),
                                                    keep_layers=det.layer_indices)[0], det.layer_indices, 2,
                                 det.encoder.tokens_per_frame)
    det.train()
    assert det.adapter.needs_autograd()
    before = det.adapter.l0_k[1].running_mean.clone()
    det.predict(x.to(cuda_device), m.to(cuda_device))
    assert not torch.equal(det.adapter.l0_k[1].running_mean, before)

"""CPU tests (-m "not gpu"): pin the oracle restatement (oracle/dfd_oracle.py) to golden vectors produced by the
UNMODIFIED reference (oracle/gen_golden.py, run in the build container where /root/reference exists)."""
import numpy as np
import pytest
import torch

from helpers import cosine, golden_inputs, golden_op_mode, golden_tensor, load_golden, load_oracle

CASES = ["tiny", "small", "vitb16", "vitl14"]


@pytest.fixture(scope="module")
def oracle():
    return load_oracle()


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(oracle, case):
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    with torch.no_grad():
        enc = oracle.encoder_forward(sd, x.flatten(0, 1), with_out=True, with_q=True)
        logits, feat = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,))
    # per-layer q/k/v/out of every encoder layer: fp32 restatement of the same torch ops -> ~1e-5
    for layer, a in enumerate(enc):
        for key in ("q", "k", "v", "out"):
            ref, got = golden_tensor(g, key, layer, a[key])
            scale = ref.abs().max().item()
            assert (got - ref).abs().max().item() <= 2e-4 * max(scale, 1.0), (case, layer, key)
            assert abs(a[key].float().norm().item() / float(g["norm_%s_%d" % (key, layer)]) - 1) < 1e-5
    assert np.abs(logits[0].numpy() - g["logits"]).max() < 1e-4
    assert np.abs(feat.numpy() - g["video_feature"]).max() < 2e-4
    assert np.array_equal(logits[0].argmax(-1).numpy(), g["pred_labels"])
    losses = oracle.detector_eval_losses(logits, [torch.from_numpy(g["labels"])])
    assert np.abs(losses[0].numpy() - g["losses"]).max() < 1e-4


ADAPTER_CASES = ["tiny_ad_x", "tiny_ad_legacy", "tiny_ad_nln", "tiny_ad_ln", "tiny_ad_z0", "tiny_ad_xxx",
                 "tiny_ad_linear", "vitb16_ad_nln", "vitb16_ad_z0", "vitb16_ad_bn"]


@pytest.mark.parametrize("case", ADAPTER_CASES)
def test_oracle_adapter_matches_reference_golden(oracle, case):
    """CompInvAdapter (src/models.py:783-940) between the taps and the decoder, every Sequential layout."""
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    struct = str(g["adapter"])
    with torch.no_grad():
        logits, feat, raw = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,), return_taps=True,
                                                    adapter=struct)
    assert np.abs(logits[0].numpy() - g["logits"]).max() < 1e-4
    assert np.abs(feat.numpy() - g["video_feature"]).max() < 2e-4
    assert np.array_equal(logits[0].argmax(-1).numpy(), g["pred_labels"])
    for i, kv in enumerate(raw):  # return_taps hands back the adapted taps the decoder consumed
        for key in ("k", "v"):
            ref, got = golden_tensor(g, "adapt_" + key, i, kv[key])
            assert (got - ref).abs().max().item() <= 2e-4 * max(ref.abs().max().item(), 1.0), (case, i, key)


MODE_CASES = ["tiny_aug_query", "tiny_global_pred", "small_gp_aq", "tiny_ema", "tiny_no_tpos", "tiny_attn_frame",
              "tiny_attn_tf", "small_attn_temporal", "small_pm_batch", "small_pm_sample", "tiny_pm_adapter"]


@pytest.mark.parametrize("case", MODE_CASES)
def test_oracle_decoder_modes_match_reference_golden(oracle, case):
    """Non-default op_mode / train_mode switches of the decoder (src/models.py:107-115, 250-267, 345-357, 511-544,
    572-578) against the unmodified reference. attn_mode with "frame" gives NaN for clips with a padded frame in the
    reference (softmax over an all -inf frame, :111): NaN positions must coincide."""
    g = load_golden(case)
    sd, x, m = golden_inputs(g)
    op = golden_op_mode(g)
    attn_mode = tuple(op["attn_mode"].split("+")) if "attn_mode" in op else ()
    adapter = str(g["adapter"]) if "adapter" in g else None
    with torch.no_grad():
        if op.get("ema_frame"):
            x, m = oracle.ema_frames(x, m, op["ema_frame"])
        pi = [torch.from_numpy(i) for i in g["patch_indices"]] if "patch_indices" in g else None
        logits, feat = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,), adapter=adapter,
                                               attn_mode=attn_mode, patch_indices=pi)
    got, ref = logits[0].numpy(), g["logits"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref)) < 1e-4
    if g["video_feature"].size:
        assert np.nanmax(np.abs(feat.numpy() - g["video_feature"])) < 2e-4
        assert feat.numpy().shape == g["video_feature"].shape


def test_z0_adapter_is_identity_at_init(oracle):
    """'768-x-768-z0' zero-initialises the LayerNorm weight and the up projection (src/models.py:856-858):
    a freshly constructed adapter leaves the taps unchanged."""
    from dfdclip_b200 import synthetic
    sd = synthetic.adapter_state_dict("tiny-256x4", 1, "768-x-768-z0", 256, seed=0)
    sd = {"adapter." + k: (torch.zeros_like(v) if k.endswith(("1.weight", "4.weight")) else v) for k, v in sd.items()}
    kv = [dict(k=torch.randn(2, 3, 4, 4, 64), v=torch.randn(2, 3, 4, 4, 64))]
    out = oracle.adapter_forward(sd, kv, "768-x-768-z0")
    assert torch.equal(out[0]["k"], kv[0]["k"]) and torch.equal(out[0]["v"], kv[0]["v"])


def test_logits_have_norm_five(oracle):
    g = load_golden("tiny")
    assert np.allclose(np.linalg.norm(g["logits"], axis=-1), 5.0, atol=1e-4)


def test_masked_frames_do_not_matter(oracle):
    """Frames masked out by m must not influence the clip logits (src/models.py:104, 124)."""
    g = load_golden("tiny")
    sd, x, m = golden_inputs(g)
    assert not m.all()
    x2 = x.clone()
    x2[~m] = torch.randn_like(x2[~m]) * 10
    with torch.no_grad():
        a, _ = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,))
        b, _ = oracle.detector_predict(sd, x2, m, g["layer_indices"], (2,))
    assert torch.allclose(a[0], b[0], atol=1e-5)


def test_all_masked_clip_is_nan(oracle):
    """An all-masked clip gives NaN in the reference (softmax over all -inf, SURVEY 8c caveat 4)."""
    qs = torch.randn(1, 1, 4, 128)
    k = torch.randn(1, 6, 4, 64)
    out = oracle.decoder_attention(qs, k, k, torch.zeros(1, 6, dtype=torch.bool))
    assert torch.isnan(out).all()


def test_video_scores_mean_of_probs(oracle):
    logits = torch.tensor([[5.0, 0.0], [0.0, 5.0], [3.0, 4.0]])
    s = oracle.video_scores(logits, [2, 1])
    assert torch.allclose(s[0], logits[:2].softmax(-1).mean(0))
    assert torch.allclose(s[1], logits[2].softmax(-1))


def test_cosine_helper():
    a = torch.randn(100)
    assert abs(cosine(a, a) - 1) < 1e-12


# ------------------------------------------------------------------ BASELINE-size fixtures (oracle/gen_golden_full.py)
@pytest.mark.parametrize("case,clips", [("vitb16_c2", slice(0, 3)), ("vitb16_c3", slice(8, 10)), ("vitl14_c4", slice(1, 2))])
def test_oracle_matches_fullsize_reference_golden(oracle, case, clips):
    """The oracle port on a few clips of each BASELINE-size fixture (the reference ran all of them): logits, labels,
    video features. Clips are independent, so a slice reproduces its rows of the reference's chunked run."""
    from helpers import fullsize_inputs
    g = load_golden(case)
    sd, x, m = fullsize_inputs(case, g, clips)
    with torch.no_grad():
        logits, feat = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,))
    assert np.abs(logits[0].numpy() - g["logits"][clips]).max() < 2e-4
    margin = np.abs(g["margin"][clips])
    agree = logits[0].argmax(-1).numpy() == g["pred_labels"][clips]
    assert agree[margin > 1e-3].all()
    if g["video_feature"].size:
        assert np.abs(feat.numpy() - g["video_feature"][clips]).max() < 5e-4


def test_fullsize_fixtures_stress_the_label_criterion():
    """C2 holds both classes and near ties (|margin| < 0.05); C3's per-video scores follow from its clip logits."""
    g = load_golden("vitb16_c2")
    assert g["batch"] == 64 and set(g["pred_labels"].tolist()) == {0, 1}
    assert (np.abs(g["margin"]) < 0.05).any() and (np.abs(g["margin"]) > 1.0).any()
    g3 = load_golden("vitb16_c3")
    probs = torch.from_numpy(g3["logits"]).softmax(-1)
    s = 0
    for v, n in enumerate(g3["counts"]):
        assert np.abs(probs[s:s + n].mean(0).numpy() - g3["video_scores"][v]).max() < 1e-6
        s += int(n)
    assert s == g3["batch"] and 5 * 8 <= s <= 5 * 32


def test_oracle_training_step_matches_c5_golden(oracle):
    """Detector.forward(train=True) + backward of the mean loss on the C5 fixture's first 3 clips is not what the
    reference ran (it ran all 12), so only quantities that do not mix clips are compared: logits and per-clip losses;
    the full-batch gradients are compared on the GPU (tests/test_fullsize_gpu.py)."""
    from helpers import fullsize_inputs
    g = load_golden("vitb16_c5")
    sd, x, m = fullsize_inputs("vitb16_c5", g, slice(0, 3))
    with torch.no_grad():
        logits, _ = oracle.detector_predict(sd, x, m, g["layer_indices"], (2,))
    assert np.abs(logits[0].numpy() - g["logits"][:3]).max() < 2e-4
    losses = oracle.detector_eval_losses(logits, [torch.from_numpy(g["labels"][:3])])
    assert np.abs(losses[0].numpy() - g["losses"][:3]).max() < 2e-4
    assert abs(float(g["loss"]) - g["losses"].mean()) < 1e-6
    assert len(g["grad_names"]) == 6 * 12 + 7   # 6 blocks x 12 tensors + class / positional embedding, ln_pre / ln_post, proj

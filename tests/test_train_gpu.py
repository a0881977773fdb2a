"""Training step (BASELINE config C5): frozen encoder forward with K/V taps + decoder forward/backward.
The CUDA path (native attention forward/backward inside torch autograd) is compared with the oracle's autograd."""
import pytest
import torch

from helpers import cosine, load_oracle

pytestmark = pytest.mark.gpu


def _attention_ref(qs, k, v, pe, mask):
    """fp32 restatement of src/models.py:99-146 (smax + coda, no projections), differentiable."""
    b, t, p, h, dh = k.shape
    k = k.float()
    v = v.float()
    if pe is not None:
        k = k + pe.view(1, t, 1, h, dh)
        v = v + pe.view(1, t, 1, h, dh)
    k = k.flatten(1, 2)
    v = v.flatten(1, 2)
    m = mask.repeat_interleave(p, dim=-1).unsqueeze(1).unsqueeze(-1)
    q0 = qs[:, :, :64].unsqueeze(1)
    q1 = qs[:, :, 64:].unsqueeze(1)
    aff0 = torch.einsum("nqhc,nkhc->nqkh", q0 / 8.0, k).masked_fill(~m, float("-inf")).softmax(dim=-2)
    aff1 = torch.einsum("nqhc,nkhc->nqkh", q1 / 8.0, k).tanh()
    gate = -(q1 - k).abs().sum(-1).unsqueeze(1) / 8.0
    gate = 2 * gate.sigmoid().masked_fill(~m, 0.0)
    aff = (aff0 + aff1 * gate) / 2
    return torch.einsum("nqlh,nlhc->nqhc", aff, v).flatten(-2).squeeze(1)


@pytest.mark.parametrize("b,t,p,h", [(2, 8, 196, 12), (3, 4, 50, 16), (2, 3, 17, 4), (4, 2, 196, 8)])
@pytest.mark.parametrize("use_pe", [True, False])
def test_decoder_attention_backward(cuda_device, b, t, p, h, use_pe):
    from dfdclip_b200 import _native as nat
    g = torch.Generator(device="cpu").manual_seed(b * 10 + t)
    d = h * 64
    buf = torch.randn(b * t * (p + 1), 3 * d, generator=g).to(cuda_device, torch.bfloat16)
    view = buf.view(b, t, p + 1, 3, h, 64)
    k, v = view[:, :, 1:, 1], view[:, :, 1:, 2]
    qs = (torch.randn(b, h, 128, generator=g) * 0.7).to(cuda_device)
    pe = (torch.randn(t, h, 64, generator=g) * 0.3).to(cuda_device) if use_pe else None
    mask = torch.ones(b, t, dtype=torch.bool, device=cuda_device)
    if t > 1:
        mask[0, -1] = False
    dmix = torch.randn(b, d, generator=g).to(cuda_device)

    mix, stats = nat.decoder_attention_train(qs, k, v, pe, mask)
    dqs, dpe = nat.decoder_attention_backward(qs, k, v, pe, mask, stats, dmix)
    torch.cuda.synchronize()

    with torch.enable_grad():
        qs_r = qs.clone().requires_grad_(True)
        pe_r = pe.clone().requires_grad_(True) if use_pe else None
        ref = _attention_ref(qs_r, k, v, pe_r, mask)
        ref.backward(dmix)
    assert (mix - ref.detach()).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert cosine(dqs, qs_r.grad) > 0.9999
    assert (dqs - qs_r.grad).abs().max().item() < 2e-3 * max(1e-3, qs_r.grad.abs().max().item())
    if use_pe:
        assert cosine(dpe, pe_r.grad) > 0.9999
        assert (dpe - pe_r.grad).abs().max().item() < 2e-3 * max(1e-3, pe_r.grad.abs().max().item())
    else:
        assert dpe is None


def test_training_step_matches_oracle_autograd(cuda_device):
    """loss = mean CE(5 l/|l|, y) through Detector.forward(train=True); gradients of every decoder parameter against
    the oracle's autograd (fp32, CPU) on the same weights and clips; one SGD step then lowers the loss."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    oracle = load_oracle()
    arch, frames, clips = "small-512x6", 3, 4
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    det = Detector(cfg, frames, None)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det.layer_indices, seed=0)
    det.load_state_dict(sd, strict=True)
    det = det.to(cuda_device).train()
    x, m = synthetic.make_clips(clips, frames, synthetic.vit_dims(arch)["image_size"], seed=5)
    y = torch.tensor([0, 1, 1, 0])

    with torch.enable_grad():
        losses, logits, other = det(x.to(cuda_device), [y.to(cuda_device)], m.to(cuda_device), train=True,
                                    single_task=0)
        assert other == {}
        loss = losses[0].mean()
        loss.backward()
        # oracle: same computation in fp32 on the CPU with autograd on the decoder parameters
        sd_r = {k_: (v_.clone().requires_grad_(True) if k_.startswith("decoder.") else v_) for k_, v_ in sd.items()}
        ref_logits, _ = oracle.detector_predict(sd_r, x, m, det.layer_indices, (2,))
        ref_loss = oracle.detector_eval_losses(ref_logits, [y])[0].mean()
        ref_loss.backward()

    assert abs(loss.item() - ref_loss.item()) < 2e-2
    assert all(not p.requires_grad and p.grad is None for p in det.encoder.parameters())
    checked = 0
    for name, p in det.decoder.named_parameters():
        ref_g = sd_r["decoder." + name].grad
        assert p.grad is not None, name
        assert ref_g is not None, name
        if ref_g.abs().max().item() < 1e-7:
            assert p.grad.abs().max().item() < 1e-4, name
            continue
        assert cosine(p.grad.cpu(), ref_g) > 0.995, (name, cosine(p.grad.cpu(), ref_g))
        rel = (p.grad.cpu() - ref_g).norm().item() / ref_g.norm().item()
        assert rel < 0.1, (name, rel)
        checked += 1
    assert checked >= 30

    # one SGD step (reference: configure_optimizers -> SGD momentum 0.95) lowers the loss on the same batch
    with torch.enable_grad():
        g2 = sum(p.grad.double().pow(2).sum().item() for p in det.decoder.parameters())
        opt = det.configure_optimizers(lr=0.05 / g2)  # first-order decrease of 0.05: a small step along -grad
        opt.step()
        opt.zero_grad()
        losses2, _, _ = det(x.to(cuda_device), [y.to(cuda_device)], m.to(cuda_device), train=True, single_task=0)
    assert losses2[0].mean().item() < loss.item()


@pytest.mark.parametrize("b,t,p,h", [(2, 8, 196, 12), (3, 4, 50, 16), (2, 3, 17, 4)])
def test_decoder_attention_backward_kv_grads(cuda_device, b, t, p, h):
    """dK / dV of the decoder attention (needed by a trainable adapter on the taps, src/models.py:546-547) against
    autograd of the fp32 restatement on the same bf16 K/V; keys of a masked frame get exactly zero."""
    from dfdclip_b200 import _native as nat
    g = torch.Generator(device="cpu").manual_seed(b * 7 + t)
    d = h * 64
    buf = torch.randn(b * t * (p + 1), 3 * d, generator=g).to(cuda_device, torch.bfloat16)
    view = buf.view(b, t, p + 1, 3, h, 64)
    k, v = view[:, :, 1:, 1], view[:, :, 1:, 2]
    qs = (torch.randn(b, h, 128, generator=g) * 0.7).to(cuda_device)
    pe = (torch.randn(t, h, 64, generator=g) * 0.3).to(cuda_device)
    mask = torch.ones(b, t, dtype=torch.bool, device=cuda_device)
    mask[0, -1] = False
    dmix = torch.randn(b, d, generator=g).to(cuda_device)
    mix, stats = nat.decoder_attention_train(qs, k, v, pe, mask)
    dqs, dpe, dk, dv = nat.decoder_attention_backward(qs, k, v, pe, mask, stats, dmix, need_kv_grad=True)
    dqs2, dpe2 = nat.decoder_attention_backward(qs, k, v, pe, mask, stats, dmix)
    torch.cuda.synchronize()
    assert torch.equal(dqs, dqs2) and torch.equal(dpe, dpe2)  # the extra outputs do not change the others
    with torch.enable_grad():
        k_r = k.float().clone().requires_grad_(True)
        v_r = v.float().clone().requires_grad_(True)
        _attention_ref(qs, k_r, v_r, pe, mask).backward(dmix)
    for got, ref, name in ((dk, k_r.grad, "dk"), (dv, v_r.grad, "dv")):
        assert tuple(got.shape) == (b, t, p, h, 64)
        assert cosine(got, ref) > 0.9999, name
        assert (got - ref).abs().max().item() < 2e-3 * max(1e-3, ref.abs().max().item()), name
        assert got[0, -1].abs().max().item() == 0.0
    # dpos_emb is the sum of dK and dV over clips and patches (pe is added to both, :326-329)
    assert torch.allclose((dk + dv).sum(dim=(0, 2)), dpe, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("struct", ["768-x-768-z0", "768-x-768-nln"])
def test_training_step_with_trainable_adapter(cuda_device, struct):
    """The shipped training configuration: a trainable CompInvAdapter between the taps and the decoder. Adapter and
    decoder gradients of one step against the oracle's autograd (fp32, CPU) on the same weights and clips."""
    from dfdclip_b200 import synthetic
    from test_adapter_gpu import build_adapter_detector
    oracle = load_oracle()
    arch, frames, clips = "small-512x6", 3, 4
    det, sd = build_adapter_detector(arch, frames, struct, cuda_device)
    det.train()
    x, m = synthetic.make_clips(clips, frames, synthetic.vit_dims(arch)["image_size"], seed=5)
    y = torch.tensor([0, 1, 1, 0])
    with torch.enable_grad():
        losses, logits, other = det(x.to(cuda_device), [y.to(cuda_device)], m.to(cuda_device), train=True,
                                    single_task=0)
        loss = losses[0].mean()
        loss.backward()
        sd_r = {k_: (v_.clone().requires_grad_(True) if not k_.startswith("encoder.") else v_) for k_, v_ in sd.items()}
        ref_logits, _ = oracle.detector_predict(sd_r, x, m, det.layer_indices, (2,), adapter=struct)
        ref_loss = oracle.detector_eval_losses(ref_logits, [y])[0].mean()
        ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 3e-2
    assert all(p.grad is None for p in det.encoder.parameters())
    checked = {"adapter": 0, "decoder": 0}
    for name, p in det.named_parameters():
        if name.startswith("encoder."):
            continue
        ref_g = sd_r[name].grad
        assert p.grad is not None and ref_g is not None, name
        if ref_g.abs().max().item() < 1e-7:
            continue
        c = cosine(p.grad.cpu(), ref_g)
        assert c > 0.99, (name, c)
        checked[name.split(".")[0]] += 1
    assert checked["adapter"] >= 4 * len(det.layer_indices) and checked["decoder"] >= 30
    # eval mode afterwards takes the native in-place adapter and agrees with the autograd path
    det.eval()
    a, _ = det.predict(x.to(cuda_device), m.to(cuda_device))
    assert (a[0] - logits[0].detach()).abs().max().item() < 2e-2


@pytest.mark.parametrize("mode,key", [("ranking", "speed/rank"), ("triplet", "speed/triplet")])
def test_temporal_train_mode_matches_reference_golden(cuda_device, mode, key):
    """train_mode.temporal (src/models.py:676-736): the speed ranking / triplet loss of Detector.forward(train=True)
    on the video features, against the unmodified reference (oracle/gen_golden.py, cases small_tm_*)."""
    import random
    import numpy as np
    from helpers import load_golden
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    g = load_golden("small_tm_" + mode)
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + g["arch"]
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    cfg.train_mode["temporal"] = mode
    det = Detector(cfg, g["num_frames"], None)
    sd = synthetic.detector_state_dict(g["arch"], g["num_frames"], out_dims=(2,), taps=det.layer_indices, seed=0,
                                       ranking=(mode == "ranking"))
    det.load_state_dict(sd, strict=True)
    det = det.to(cuda_device).eval()
    x, m = synthetic.make_clips(g["batch"], g["num_frames"], synthetic.vit_dims(g["arch"])["image_size"], seed=7)
    y = torch.from_numpy(g["labels"]).to(cuda_device)
    speed = torch.from_numpy(g["speed"]).to(cuda_device)
    comp = ["raw" if i % 2 == 0 else "c23" for i in range(g["batch"])]
    random.seed(11)
    with torch.enable_grad():
        losses, logits, other = det(x.to(cuda_device), [y], m.to(cuda_device), comp=comp, speed=speed, train=True,
                                    single_task=0)
        (losses[0].mean() + other[key]).backward()
    ref = float(g["other_" + key.replace("/", "_")])
    assert list(other) == [key]
    assert abs(other[key].item() - ref) <= 0.03 * abs(ref) + 1e-4, (other[key].item(), ref)
    assert np.abs(losses[0].detach().cpu().numpy() - g["losses"]).max() <= 4e-2
    if mode == "ranking":
        assert det.ranking_transform_param.grad is not None
    assert det.decoder.class_embedding.grad is not None


def test_train_modes_the_reference_cannot_run_raise():
    from dfdclip_b200.models import Detector
    for key, val in (("compression", "sync"), ("nerf_raw", 0.5), ("temporal", "order")):
        cfg = Detector.get_default_config()
        cfg.architecture = "synthetic:tiny-256x4"
        cfg.out_dim = [2]
        cfg.train_mode[key] = val
        with pytest.raises(NotImplementedError):
            Detector(cfg, 4, None)


@pytest.mark.parametrize("u8", [False, True])
def test_graphed_train_step_matches_eager_steps(cuda_device, u8):
    """GraphedTrainStep (forward(train=True) + backward + SGD step captured once, replayed per batch) against the same
    three steps run eagerly on an identical detector: losses, logits and every trained parameter agree, constructing the
    step does not change the weights, and a batch of another shape is refused."""
    import copy
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    from dfdclip_b200.training import GraphedTrainStep
    arch, frames, clips = "small-512x6", 3, 4
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    det_a = Detector(cfg, frames, None)
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det_a.layer_indices, seed=0)
    det_a.load_state_dict(sd, strict=True)
    det_b = copy.deepcopy(det_a)
    det_a, det_b = det_a.to(cuda_device).train(), det_b.to(cuda_device).train()
    res = synthetic.vit_dims(arch)["image_size"]
    batches = []
    for i in range(3):
        x, m = synthetic.make_clips(clips, frames, res, seed=60 + i)
        if u8:
            x = torch.randint(0, 256, tuple(x.shape), generator=torch.Generator().manual_seed(80 + i), dtype=torch.uint8)
        y = torch.randint(0, 2, (clips,), generator=torch.Generator().manual_seed(70 + i))
        batches.append((x.to(cuda_device), y.to(cuda_device), m.to(cuda_device)))

    opt_a = det_a.configure_optimizers(lr=0.05)
    eager = []
    for x, y, m in batches:
        with torch.enable_grad():
            losses, logits, _ = det_a(x, [y], m, train=True, single_task=0)
            loss = losses[0].mean()
            loss.backward()
        opt_a.step()
        opt_a.zero_grad(set_to_none=True)
        eager.append((loss.item(), logits[0].detach().clone()))

    opt_b = det_b.configure_optimizers(lr=0.05)
    before = [p.detach().clone() for p in det_b.parameters()]
    step = GraphedTrainStep(det_b, opt_b, *batches[2])  # example batch: any batch of the right shape
    torch.cuda.synchronize()
    assert all(torch.equal(p, q) for p, q in zip(det_b.parameters(), before))
    for (x, y, m), (ref_loss, ref_logits) in zip(batches, eager):
        loss, logits = step(x, y, m)
        assert abs(loss.item() - ref_loss) < 1e-4
        assert (logits - ref_logits).abs().max().item() < 1e-3
    moved = 0
    for (name, p), q, p0 in zip(det_b.named_parameters(), det_a.parameters(), before):
        if not p.requires_grad:
            assert torch.equal(p, p0), name
            continue
        delta = (q - p0).norm().item()
        assert (p - q).norm().item() <= 1e-3 * delta + 1e-6, (name, (p - q).norm().item(), delta)
        moved += delta > 0
    assert moved >= 30
    with pytest.raises(ValueError):
        step(batches[0][0][:2], batches[0][1][:2], batches[0][2][:2])


@pytest.mark.parametrize("graph", [True, False])
def test_pipelined_train_step_matches_the_plain_step(cuda_device, graph):
    """TrainStep(pipeline=True) — the frozen encoder of batch k+1 beside the decoder step of batch k — against the plain
    TrainStep on the same five batches: the loss / logits of batch k come back one call later and are identical, the
    first call returns nothing, flush() trains the last batch, and the trained parameters end up bit-identical; also
    through run_host from pinned batches."""
    import copy
    from dfdclip_b200 import synthetic
    from dfdclip_b200.training import TrainStep
    det_a, _ = _build_small_detector(cuda_device)
    det_b = copy.deepcopy(det_a)
    res = 64
    batches = []
    for i in range(5):
        x, m = synthetic.make_clips(4, 3, res, seed=90 + i)
        if i == 3:
            m[1, -1] = False
        y = torch.randint(0, 2, (4,), generator=torch.Generator().manual_seed(40 + i))
        batches.append((x, y, m))
    dev_batches = [tuple(t.to(cuda_device) for t in b) for b in batches]

    plain = TrainStep(det_a, det_a.configure_optimizers(lr=0.05), *dev_batches[0], graph=graph)
    ref = []
    for b in dev_batches:
        loss, logits = plain(*b)
        ref.append((loss.clone(), logits.clone()))

    before = [p.detach().clone() for p in det_b.parameters()]
    piped = TrainStep(det_b, det_b.configure_optimizers(lr=0.05), *dev_batches[4], graph=graph, pipeline=True)
    torch.cuda.synchronize()
    assert all(torch.equal(p, q) for p, q in zip(det_b.parameters(), before))   # constructing it does not train
    got = []
    for k, b in enumerate(dev_batches[:3]):
        loss, logits = piped(*b)
        if k == 0:
            assert loss is None and logits is None
        else:
            got.append((loss.clone(), logits.clone()))
    got.append(tuple(t.clone() for t in piped.flush()))
    assert piped.flush() == (None, None)
    # the last two batches from the host, H2D under the running step
    pinned = [tuple(t.pin_memory() for t in b) for b in batches[3:]]
    for loss, logits in piped.run_host(iter(pinned)):
        got.append((loss.clone(), logits.clone()))
    torch.cuda.synchronize()
    assert len(got) == len(ref) == 5
    for (l0, g0), (l1, g1) in zip(ref, got):
        assert torch.equal(l0, l1) and torch.equal(g0, g1)
    for (name, p), q in zip(det_a.named_parameters(), det_b.parameters()):
        assert torch.equal(p, q), name
    plain.close()
    piped.close()


# ------------------------------------------------------------------------- native decoder backward (dfd_decoder_train_*)
@pytest.mark.parametrize("b,n,k", [(12, 1536, 768), (12, 768, 3072), (3, 256, 256), (33, 3072, 768), (1, 768, 768)])
@pytest.mark.parametrize("gelu,add", [(False, False), (True, True)])
def test_linear_f32_backward(cuda_device, b, n, k, gelu, add):
    """dX = dY W (* quickgelu'(pre)) (+ add), dW = dY^T X, db = sum_b dY of one decoder nn.Linear against torch autograd."""
    from dfdclip_b200 import _native as nat
    g = torch.Generator(device="cpu").manual_seed(b + n + k)
    x = torch.randn(b, k, generator=g).to(cuda_device)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(cuda_device)
    dy = torch.randn(b, n, generator=g).to(cuda_device)
    pre = torch.randn(b, k, generator=g).to(cuda_device) if gelu else None
    extra = torch.randn(b, k, generator=g).to(cuda_device) if add else None
    dx, dw, db = nat.linear_f32_backward(x, w, dy, gelu_pre=pre, dx_add=extra)
    torch.cuda.synchronize()
    ref_dx = dy.double() @ w.double()
    if gelu:  # x here is quickgelu(pre) in the chain; the kernel only needs pre for the derivative
        with torch.enable_grad():
            p = pre.double().requires_grad_(True)
            (p * torch.sigmoid(1.702 * p)).backward(ref_dx)
        ref_dx = p.grad
    if add:
        ref_dx = ref_dx + extra.double()
    ref_dw = dy.double().t() @ x.double()
    assert (dx.double() - ref_dx).abs().max().item() < 1e-4 * max(1.0, ref_dx.abs().max().item())
    assert (dw.double() - ref_dw).abs().max().item() < 1e-4 * max(1.0, ref_dw.abs().max().item())
    assert (db.double() - dy.double().sum(0)).abs().max().item() < 1e-4
    only_dx, none_w, none_b = nat.linear_f32_backward(x, w, dy, gelu_pre=pre, dx_add=extra, need_dw=False)
    assert torch.equal(only_dx, dx) and none_w is None and none_b is None


def _build_small_detector(device, op_mode=None, arch="small-512x6", frames=3):
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    for key, val in (op_mode or {}).items():
        cfg.op_mode[key] = val
    det = Detector(cfg, frames, None)
    op = op_mode or {}
    sd = synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det.layer_indices, seed=0,
                                       aug_query=bool(op.get("aug_query")),
                                       global_prediction=bool(op.get("global_prediction")),
                                       temporal_position=bool(op.get("temporal_position", 1)))
    det.load_state_dict(sd, strict=True)
    return det.to(device).train(), sd


@pytest.mark.parametrize("op_mode", [None, {"aug_query": 1}, {"global_prediction": 1, "aug_query": 1},
                                     {"temporal_position": 0}])
def test_native_decoder_chain_backward_matches_the_torch_module_path(cuda_device, monkeypatch, op_mode):
    """The one-node native chain (dfd_decoder_train_forward / _backward) against the same step with the chain's
    LayerNorm / linear layers as torch modules under autograd (DFD_NATIVE_DECODER_BWD=0; same attention kernels):
    loss, logits and every gradient agree to fp32 round-off, for every op_mode that changes the chain's wiring."""
    from dfdclip_b200 import synthetic
    det, _ = _build_small_detector(cuda_device, op_mode)
    x, m = synthetic.make_clips(5, 3, 64, seed=21)
    y = torch.tensor([0, 1, 1, 0, 1], device=cuda_device)
    results = []
    for native in ("0", "1"):
        monkeypatch.setenv("DFD_NATIVE_DECODER_BWD", native)
        det.zero_grad(set_to_none=True)
        with torch.enable_grad():
            losses, logits, _ = det(x.to(cuda_device), [y], m.to(cuda_device), train=True, single_task=0)
            losses[0].mean().backward()
        results.append((losses[0].detach().clone(), logits[0].detach().clone(),
                        {n: p.grad.detach().clone() for n, p in det.named_parameters() if p.requires_grad}))
    (l0, g0, gr0), (l1, g1, gr1) = results
    assert (l0 - l1).abs().max().item() < 1e-5 and (g0 - g1).abs().max().item() < 1e-4
    assert set(gr0) == set(gr1) and len(gr1) >= 40
    for name in gr0:
        scale = max(gr0[name].abs().max().item(), 1e-6)
        err = (gr0[name] - gr1[name]).abs().max().item()
        assert err < 2e-4 * scale + 1e-7, (name, err, scale)


@pytest.mark.parametrize("op_mode", [None, {"global_prediction": 1, "aug_query": 1}])
def test_single_call_training_forward_is_bit_identical_to_two_calls(cuda_device, monkeypatch, op_mode):
    """dfd_train_forward (encoder with the activation-saving decoder blocks on the side stream behind their taps)
    against dfd_encoder_forward + dfd_decoder_train_forward (DFD_OVERLAP=0): same kernels on the same data, so loss,
    logits and every gradient are bit-identical — also for uint8 frames. The second setting also moves the backward's
    weight-gradient kernels onto the side stream (DFD_BWD_STREAMS)."""
    from dfdclip_b200 import synthetic
    det, _ = _build_small_detector(cuda_device, op_mode)
    x, m = synthetic.make_clips(5, 3, 64, seed=33)
    y = torch.tensor([1, 0, 1, 0, 0], device=cuda_device)
    mean = torch.tensor(det.encoder.input_mean).view(1, 1, 3, 1, 1)
    std = torch.tensor(det.encoder.input_std).view(1, 1, 3, 1, 1)
    x_u8 = ((x * std + mean).clamp(0, 1) * 255).round().to(torch.uint8)
    for frames in (x, x_u8):
        results = []
        for overlap in ("0", "1"):
            monkeypatch.setenv("DFD_OVERLAP", overlap)
            monkeypatch.setenv("DFD_BWD_STREAMS", overlap)  # weight gradients beside the dx chain of the backward
            det.zero_grad(set_to_none=True)
            with torch.enable_grad():
                assert det._train_single_call_ok(True) == (overlap == "1")
                losses, logits, _ = det(frames.to(cuda_device), [y], m.to(cuda_device), train=True, single_task=0)
                losses[0].mean().backward()
            torch.cuda.synchronize()
            results.append((losses[0].detach().clone(), logits[0].detach().clone(),
                            {n: p.grad.detach().clone() for n, p in det.named_parameters() if p.requires_grad}))
        (l0, g0, gr0), (l1, g1, gr1) = results
        assert torch.equal(l0, l1) and torch.equal(g0, g1)
        assert set(gr0) == set(gr1) and len(gr1) >= 40
        for name in gr0:
            assert torch.equal(gr0[name], gr1[name]), name


@pytest.mark.parametrize("mode", ["frame", "temporal+frame"])
def test_training_step_with_attn_mode_matches_oracle(cuda_device, mode):
    """op_mode.attn_mode in the training step (the shipped deepfake configs train with it): loss and decoder gradients
    against the oracle's autograd. Clips without padded frames (a padded frame gives NaN in 'frame' mode, :111)."""
    from dfdclip_b200 import synthetic
    oracle = load_oracle()
    det, sd = _build_small_detector(cuda_device, {"attn_mode": mode})
    x, m = synthetic.make_clips(4, 3, 64, seed=9, masked_tail=False)
    y = torch.tensor([1, 0, 0, 1])
    with torch.enable_grad():
        losses, logits, _ = det(x.to(cuda_device), [y.to(cuda_device)], m.to(cuda_device), train=True, single_task=0)
        loss = losses[0].mean()
        loss.backward()
        sd_r = {k_: (v_.clone().requires_grad_(True) if k_.startswith("decoder.") else v_) for k_, v_ in sd.items()}
        ref_logits, _ = oracle.detector_predict(sd_r, x, m, det.layer_indices, (2,), attn_mode=tuple(mode.split("+")))
        ref_loss = oracle.detector_eval_losses(ref_logits, [y])[0].mean()
        ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 3e-2
    checked = 0
    for name, p in det.decoder.named_parameters():
        ref_g = sd_r["decoder." + name].grad
        if ref_g is None or ref_g.abs().max().item() < 1e-7:
            continue
        assert cosine(p.grad.cpu(), ref_g) > 0.99, (name, cosine(p.grad.cpu(), ref_g))
        checked += 1
    assert checked >= 30


def test_train_step_follows_a_learning_rate_schedule(cuda_device):
    """The captured step reads lr from a device tensor: three replays under OneCycleLR (cycle_momentum=False, synced with
    step.sync_lr) equal three eager steps under the same scheduler; a cycled momentum is refused instead of ignored."""
    import copy
    from dfdclip_b200 import synthetic
    from dfdclip_b200.training import TrainStep
    det_a, _ = _build_small_detector(cuda_device)
    det_b = copy.deepcopy(det_a)
    batches = []
    for i in range(3):
        x, m = synthetic.make_clips(4, 3, 64, seed=60 + i)
        y = torch.randint(0, 2, (4,), generator=torch.Generator().manual_seed(70 + i))
        batches.append((x.to(cuda_device), y.to(cuda_device), m.to(cuda_device)))
    opt_a = det_a.configure_optimizers(lr=0.002)
    sch_a = torch.optim.lr_scheduler.OneCycleLR(opt_a, max_lr=0.05, total_steps=3, cycle_momentum=False)
    for x, y, m in batches:
        with torch.enable_grad():
            losses, _, _ = det_a(x, [y], m, train=True, single_task=0)
            losses[0].mean().backward()
        opt_a.step()
        sch_a.step()
        opt_a.zero_grad(set_to_none=True)
    opt_b = det_b.configure_optimizers(lr=0.002)
    sch_b = torch.optim.lr_scheduler.OneCycleLR(opt_b, max_lr=0.05, total_steps=3, cycle_momentum=False)
    before = [p.detach().clone() for p in det_b.parameters()]
    step = TrainStep(det_b, opt_b, *batches[0])
    step.sync_lr(sch_b)
    for x, y, m in batches:
        step(x, y, m)
        sch_b.step()
        step.sync_lr(sch_b)
    torch.cuda.synchronize()
    for (name, p), q, p0 in zip(det_b.named_parameters(), det_a.parameters(), before):
        if p.requires_grad:
            delta = (q - p0).norm().item()
            assert (p - q).norm().item() <= 2e-3 * delta + 1e-6, (name, (p - q).norm().item(), delta)
    opt_b.param_groups[0]["momentum"] = 0.5
    with pytest.raises(RuntimeError):
        step(*batches[0])


def test_captured_graphs_survive_a_larger_eager_call(cuda_device):
    """Workspaces handed out under stream capture are never freed: a captured predict keeps replaying correctly after an
    eager call with a bigger batch made the encoder / decoder workspaces grow."""
    from dfdclip_b200 import synthetic
    det, _ = _build_small_detector(cuda_device)
    det.eval()
    x, m = synthetic.make_clips(3, 3, 64, seed=3)
    xs, ms = x.to(cuda_device), m.to(cuda_device)
    with torch.no_grad():
        ref = det.predict(xs, ms)[0][0].clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            det.predict(xs, ms)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = det.predict(xs, ms)[0][0]
        small_ws = det.encoder._workspace.buf
        xb, mb = synthetic.make_clips(11, 3, 64, seed=4)
        det.predict(xb.to(cuda_device), mb.to(cuda_device))        # grows every workspace
        assert det.encoder._workspace.buf is not small_ws and small_ws in det.encoder._workspace.retired
        junk = [torch.randn(1 << 22, device=cuda_device) for _ in range(8)]  # would land on a freed block
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
        del junk


def test_train_step_run_host_matches_device_batches(cuda_device):
    """TrainStep.run_host (H2D of the next batch overlapped with the current step) trains exactly like feeding the same
    batches from the device."""
    import copy
    from dfdclip_b200 import synthetic
    from dfdclip_b200.training import TrainStep
    det_a, _ = _build_small_detector(cuda_device)
    det_b = copy.deepcopy(det_a)
    host = []
    for i in range(4):
        x, m = synthetic.make_clips(4, 3, 64, seed=90 + i)
        y = torch.randint(0, 2, (4,), generator=torch.Generator().manual_seed(95 + i))
        host.append((x.pin_memory(), y.pin_memory(), m.pin_memory()))
    dev0 = tuple(t.to(cuda_device) for t in host[0])
    step_a = TrainStep(det_a, det_a.configure_optimizers(lr=0.02), *dev0)
    step_b = TrainStep(det_b, det_b.configure_optimizers(lr=0.02), *dev0)
    losses_a = [step_a(*(t.to(cuda_device) for t in b))[0].item() for b in host]
    losses_b = [loss.item() for loss, _ in step_b.run_host(iter(host))]
    assert losses_a == losses_b
    for p, q in zip(det_a.parameters(), det_b.parameters()):
        assert torch.equal(p, q)

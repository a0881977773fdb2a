"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (ctypes), against plain
PyTorch fp32 references of the same op (the reference lines are cited in include/dfdclip_b200.h)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _native():
    import dfdclip_b200._native as n
    return n


def _rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 768), (256, 768, 768), (1576, 2304, 768),
                                   (197 * 8 + 3, 768, 3072), (640, 3072, 768), (100, 256, 640), (8192, 1024, 1024)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_bf16(cuda_device, m, n, k, epi):
    nat = _native()
    g = torch.Generator(device="cpu").manual_seed(m * 31 + n * 7 + k + epi)
    a = (torch.randn(m, k, generator=g) * 1.0).to(cuda_device, torch.bfloat16)
    w = (torch.randn(n, k, generator=g) * (k ** -0.5)).to(cuda_device, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda_device)
    ref = a.float() @ w.float().t() + bias
    if epi == nat.EPI_STORE_BF16_QGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    if epi in (nat.EPI_STORE_BF16, nat.EPI_STORE_BF16_QGELU):
        out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=cuda_device)
        tol = 6e-3
    elif epi == nat.EPI_STORE_F32:
        out = torch.full((m, n), float("nan"), dtype=torch.float32, device=cuda_device)
        tol = 2e-5
    else:
        res = torch.randn(m, n, generator=g).to(cuda_device)
        out = res.clone()
        ref = ref + res
        tol = 2e-5
    nat.gemm_bf16(a, w, bias, out, epi)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "unwritten or non-finite outputs"
    err = _rel_err(out, ref)
    maxabs = (out.float() - ref).abs().max().item()
    assert err < tol, f"rel err {err:.3e} max abs {maxabs:.3e}"


def test_gemm_no_bias_and_strided(cuda_device):
    nat = _native()
    g = torch.Generator(device="cpu").manual_seed(5)
    big = torch.randn(300, 1024, generator=g).to(cuda_device, torch.bfloat16)
    a = big[:, 128:128 + 512]  # row pitch 1024, K = 512
    w = (torch.randn(256, 512, generator=g) / 22).to(cuda_device, torch.bfloat16)
    out_big = torch.zeros(300, 512, dtype=torch.float32, device=cuda_device)
    out = out_big[:, 256:]
    nat.gemm_bf16(a, w, None, out, nat.EPI_STORE_F32)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    assert _rel_err(out, ref) < 2e-5
    assert out_big[:, :256].abs().max().item() == 0.0


def test_gemm_rejects_bad_shapes(cuda_device):
    nat = _native()
    a = torch.zeros(16, 64, dtype=torch.bfloat16, device=cuda_device)
    w = torch.zeros(100, 64, dtype=torch.bfloat16, device=cuda_device)
    out = torch.zeros(16, 100, dtype=torch.float32, device=cuda_device)
    with pytest.raises(nat.NativeError):
        nat.gemm_bf16(a, w, None, out, nat.EPI_STORE_F32)


@pytest.mark.parametrize("rows,d", [(1, 768), (197 * 5, 768), (1000, 1024), (33, 128)])
@pytest.mark.parametrize("with_pos", [False, True])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_layernorm(cuda_device, rows, d, with_pos, out_dtype):
    nat = _native()
    g = torch.Generator(device="cpu").manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(cuda_device)
    gamma = torch.randn(d, generator=g).to(cuda_device)
    beta = torch.randn(d, generator=g).to(cuda_device)
    pos = torch.randn(197, d, generator=g).to(cuda_device) if with_pos else None
    xin = x if pos is None else x + pos[torch.arange(rows, device=cuda_device) % 197]
    ref = torch.nn.functional.layer_norm(xin, (d,), gamma, beta, 1e-5)
    out = nat.layernorm(x, gamma, beta, pos, out_dtype)
    torch.cuda.synchronize()
    tol = 1e-5 if out_dtype == torch.float32 else 5e-3
    assert _rel_err(out, ref) < tol


def test_layernorm_in_place(cuda_device):
    nat = _native()
    x = torch.randn(64, 768, device=cuda_device)
    gamma = torch.ones(768, device=cuda_device)
    beta = torch.zeros(768, device=cuda_device)
    ref = torch.nn.functional.layer_norm(x, (768,), gamma, beta, 1e-5)
    nat.layernorm(x, gamma, beta, None, out=x)
    torch.cuda.synchronize()
    assert _rel_err(x, ref) < 1e-5


@pytest.mark.parametrize("r,patch", [(224, 16), (224, 14), (32, 16)])
def test_patchify(cuda_device, r, patch):
    nat = _native()
    f = 3
    x = torch.randn(f, 3, r, r, device=cuda_device)
    out = nat.patchify(x, patch)
    torch.cuda.synchronize()
    g = r // patch
    k = 3 * patch * patch
    ref = torch.nn.functional.unfold(x, kernel_size=patch, stride=patch).transpose(1, 2)  # [f, P, 3*p*p] in (c,i,j)
    got = out.view(f, g * g + 1, -1)
    assert got[:, 0].abs().max().item() == 0.0
    if got.shape[-1] > k:
        assert got[:, :, k:].abs().max().item() == 0.0
    assert torch.equal(got[:, 1:, :k], ref.to(torch.bfloat16))


@pytest.mark.parametrize("n_frames,seq,heads", [(2, 197, 12), (3, 257, 16), (1, 16, 4), (2, 50, 12), (40, 197, 12),
                                                (30, 129, 8), (26, 208, 12), (64, 160, 4), (2, 128, 12),
                                                # attention_sm100_v3: 208 < L <= 257, several items per SM
                                                (40, 257, 16), (21, 257, 4), (19, 256, 8), (23, 209, 12),
                                                (17, 240, 4), (1, 257, 16)])
def test_mha_fwd(cuda_device, n_frames, seq, heads):
    nat = _native()
    d = heads * 64
    g = torch.Generator(device="cpu").manual_seed(seq)
    qkv = (torch.randn(n_frames * seq, 3 * d, generator=g) * 1.5).to(cuda_device, torch.bfloat16)
    mix = nat.mha_fwd(qkv, n_frames, seq, heads)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_frames, seq, 3, heads, 64).unbind(2)
    aff = torch.einsum("nqhc,nkhc->nqkh", q / 8.0, k).softmax(dim=-2)
    ref = torch.einsum("nqlh,nlhc->nqhc", aff, v).reshape(n_frames * seq, d)
    assert torch.isfinite(mix.float()).all()
    assert _rel_err(mix, ref) < 1e-2
    assert (mix.float() - ref).abs().max().item() < 5e-2


def _decoder_attention_ref(qs, k, v, pe, mask):
    """fp32 restatement of src/models.py:99-146 (smax + coda, no projections)."""
    b, t, p, h, dh = k.shape
    k = k.float()
    v = v.float()
    if pe is not None:
        k = k + pe.view(1, t, 1, h, dh)
        v = v + pe.view(1, t, 1, h, dh)
    k = k.flatten(1, 2)
    v = v.flatten(1, 2)
    m = mask.repeat_interleave(p, dim=-1).unsqueeze(1).unsqueeze(-1)  # [B,1,S,1]
    q0 = qs[:, :, :64].unsqueeze(1)
    q1 = qs[:, :, 64:].unsqueeze(1)
    aff0 = torch.einsum("nqhc,nkhc->nqkh", q0 / 8.0, k).masked_fill(~m, float("-inf")).softmax(dim=-2)
    aff1 = torch.einsum("nqhc,nkhc->nqkh", q1 / 8.0, k).tanh()
    gate = -(q1 - k).abs().sum(-1).unsqueeze(1) / 8.0
    gate = 2 * gate.sigmoid().masked_fill(~m, 0.0)
    aff = (aff0 + aff1 * gate) / 2
    return torch.einsum("nqlh,nlhc->nqhc", aff, v).flatten(-2).squeeze(1)


@pytest.mark.parametrize("b,t,p,h", [(2, 8, 196, 12), (3, 4, 50, 16), (1, 1, 7, 4), (5, 3, 196, 8)])
@pytest.mark.parametrize("use_pe", [True, False])
def test_decoder_attention(cuda_device, b, t, p, h, use_pe):
    nat = _native()
    g = torch.Generator(device="cpu").manual_seed(b * 100 + t)
    # K/V as strided views of a packed [B*T*(P+1), 3D] buffer, like the encoder taps
    d = h * 64
    buf = torch.randn(b * t * (p + 1), 3 * d, generator=g).to(cuda_device, torch.bfloat16)
    view = buf.view(b, t, p + 1, 3, h, 64)
    k = view[:, :, 1:, 1]
    v = view[:, :, 1:, 2]
    qs = (torch.randn(b, h, 128, generator=g) * 0.7).to(cuda_device)
    pe = (torch.randn(t, h, 64, generator=g) * 0.3).to(cuda_device) if use_pe else None
    mask = torch.ones(b, t, dtype=torch.bool, device=cuda_device)
    if t > 1:
        mask[0, -1] = False
        if b > 1:
            mask[1, t // 2:] = False
    out = nat.decoder_attention(qs, k, v, pe, mask)
    torch.cuda.synchronize()
    ref = _decoder_attention_ref(qs, k, v, pe, mask)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert _rel_err(out, ref) < 1e-4


def test_project_logits(cuda_device):
    nat = _native()
    f = torch.randn(9, 768, device=cuda_device)
    proj = torch.randn(768, 2, device=cuda_device) * 768 ** -0.5
    out = nat.project_logits(f, proj, 5.0)
    torch.cuda.synchronize()
    l = f @ proj
    ref = 5 * l / (l.norm(dim=-1, keepdim=True) + 1e-10)
    assert (out - ref).abs().max().item() < 1e-4
    assert torch.allclose(out.norm(dim=-1), torch.full((9,), 5.0, device=cuda_device), atol=1e-4)


@pytest.mark.parametrize("m,n,k", [(300, 768, 768), (1000, 768, 3072), (2500, 1024, 1024)])
def test_gemm_resid_ln(cuda_device, m, n, k):
    """DFD_EPI_RESID_LN_F32: x += a w^T + bias with the full value in hand; also bf16(x) and, per row and 128-column
    block, the (sum, sum of squares) of the updated values."""
    nat = _native()
    g = torch.Generator(device="cpu").manual_seed(m + n)
    a = (torch.randn(m, k, generator=g)).to(cuda_device, torch.bfloat16)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(cuda_device, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda_device)
    x0 = (torch.randn(m, n, generator=g) * 3).to(cuda_device)
    x = x0.clone()
    xb, stats = nat.gemm_resid_ln(a, w, bias, x)
    torch.cuda.synchronize()
    ref = x0 + a.float() @ w.float().t() + bias
    assert (x - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    assert torch.equal(xb.view(torch.int16), x.to(torch.bfloat16).view(torch.int16))  # bf16 copy of exactly what was stored
    blocks = x.view(m, n // 128, 128)
    assert tuple(stats.shape) == (m, n // 128, 2)
    assert torch.allclose(stats[:, :, 0], blocks.sum(-1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[:, :, 1], (blocks * blocks).sum(-1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("m,n,quickgelu", [(300, 2304, False), (1000, 3072, True), (2000, 1024, False)])
def test_gemm_lnfold_equals_layernorm_then_linear(cuda_device, m, n, quickgelu):
    """LayerNorm folded into the GEMM (DFD_EPI_STORE_BF16[_QGELU]_LNFOLD) against LayerNorm -> Linear in fp32; the
    residual stream has a per-row offset and a few large channels, as real CLIP activations do."""
    nat = _native()
    d = 768
    g = torch.Generator(device="cpu").manual_seed(n)
    x = torch.randn(m, d, generator=g) * 2 + torch.randn(m, 1, generator=g)
    x[:, 5] += 25.0
    x[:, 300] -= 12.0
    gamma = 1 + 0.2 * torch.randn(d, generator=g)
    beta = 0.3 * torch.randn(d, generator=g)
    w = torch.randn(n, d, generator=g) * d ** -0.5
    b = torch.randn(n, generator=g) * 0.1
    ref = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5) @ w.t() + b
    if quickgelu:
        ref = ref * torch.sigmoid(1.702 * ref)
    blocks = x.view(m, d // 128, 128)
    stats = torch.stack([blocks.sum(-1), (blocks * blocks).sum(-1)], dim=-1).contiguous()
    wf = (gamma * w).to(torch.bfloat16)
    colsum = wf.float().sum(1)
    bf = b + w @ beta
    out = torch.empty(m, n, dtype=torch.bfloat16, device=cuda_device)
    nat.gemm_lnfold(x.to(cuda_device, torch.bfloat16), stats.to(cuda_device), wf.to(cuda_device),
                    colsum.to(cuda_device), bf.to(cuda_device), out, quickgelu=quickgelu)
    torch.cuda.synchronize()
    # same tolerance class as LayerNorm -> bf16 -> GEMM: compare with that pipeline's own error
    u = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5).to(torch.bfloat16).float()
    base = u @ w.to(torch.bfloat16).float().t() + b
    if quickgelu:
        base = base * torch.sigmoid(1.702 * base)
    err_fold = (out.float().cpu() - ref).norm() / ref.norm()
    err_base = (base - ref).norm() / ref.norm()
    assert err_fold.item() < 1.5e-2, err_fold
    assert err_fold.item() < 3 * err_base.item() + 4e-3, (err_fold, err_base)


@pytest.mark.parametrize("b,n,k", [(64, 1536, 768), (64, 768, 3072), (1, 768, 768), (3, 100, 36), (65, 3072, 768),
                                   (130, 2, 768), (12, 1024, 4096), (7, 250, 1000)])
@pytest.mark.parametrize("mode", ["plain", "bias", "bias_gelu", "bias_residual_in_place"])
def test_linear_f32(cuda_device, b, n, k, mode):
    """The decoder's one-token-per-clip linear (src/models.py:136, 145, 69-73) in exact fp32 arithmetic: every split-K
    configuration, ragged tiles in all three dimensions, bias / QuickGELU / in-place residual epilogues, against an
    fp64 reference; run twice for bit-exact determinism (the last-arriving CTA reduces in a fixed order)."""
    from dfdclip_b200 import _native as nat
    g = torch.Generator().manual_seed(b * 7 + n * 3 + k)
    x = torch.randn((b, k), generator=g).to(cuda_device)
    w = (torch.randn((n, k), generator=g) * k ** -0.5).to(cuda_device)
    bias = torch.randn((n,), generator=g).to(cuda_device) if mode != "plain" else None
    res = torch.randn((b, n), generator=g).to(cuda_device) if mode == "bias_residual_in_place" else None
    ref = x.double() @ w.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if mode == "bias_gelu":
        ref = ref * torch.sigmoid(1.702 * ref)
    if res is not None:
        ref = ref + res.double()
    outs = []
    for _ in range(2):
        out = res.clone() if res is not None else None
        got = nat.linear_f32(x, w, bias, out, quick_gelu=(mode == "bias_gelu"), out=out)
        outs.append(got.clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert (outs[0].double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())

"""GPU input side: T.Resize(n_px, BICUBIC) + T.CenterCrop(n_px) of the loader transform
(src/models.py:756-761) on the device, against torchvision's CPU result on the same uint8 frames."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed):
    """Frames with structure (smooth content + noise), like decoded video, not white noise."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, 3, h // 8 + 1, w // 8 + 1, generator=g)
    img = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear") * 200 + torch.rand(n, 3, h, w, generator=g) * 55
    return img.to(torch.uint8)


def _reference(frames, size):
    import torchvision.transforms as T
    return T.Compose([T.Resize(size, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(size)])(frames)


@pytest.mark.parametrize("h,w,size", [(300, 300, 224), (180, 240, 224), (401, 333, 224), (150, 150, 224),
                                      (224, 224, 224), (720, 1280, 224), (97, 64, 32), (256, 341, 224), (1080, 608, 224)])
def test_resize_crop_matches_torchvision(cuda_device, h, w, size):
    from dfdclip_b200 import _native as nat
    x = _frames(3, h, w, seed=h + w)
    ref = _reference(x, size)
    got = nat.resize_crop_u8(x.to(cuda_device), size).cpu()
    assert got.shape == ref.shape and got.dtype == torch.uint8
    diff = (got.int() - ref.int()).abs()
    assert diff.max().item() <= 1, "max |diff| %d" % diff.max().item()
    assert (diff > 0).float().mean().item() < 1e-3   # fp32 summation order moves a few values across x.5


def test_resize_keeps_leading_dims_and_handles_empty(cuda_device):
    from dfdclip_b200 import _native as nat
    x = _frames(6, 120, 160, seed=1).view(2, 3, 3, 120, 160).to(cuda_device)
    out = nat.resize_crop_u8(x, 64)
    assert tuple(out.shape) == (2, 3, 3, 64, 64)
    assert torch.equal(out[1, 2], nat.resize_crop_u8(x[1, 2], 64))
    assert tuple(nat.resize_crop_u8(x[:0], 64).shape) == (0, 3, 3, 64, 64)
    with pytest.raises(ValueError):
        nat.resize_crop_u8(x.float(), 64)


def test_predict_on_raw_frames_matches_the_cpu_transform(cuda_device):
    """Raw uint8 clips of another size straight into Detector.predict (resize + crop + normalise + patchify on the
    GPU) against the reference route: Detector.transform on the CPU (torchvision), fp32 frames to the device."""
    from dfdclip_b200 import synthetic
    from dfdclip_b200.models import Detector
    arch, frames, clips = "small-512x6", 3, 4
    cfg = Detector.get_default_config()
    cfg.architecture = "synthetic:" + arch
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    det = Detector(cfg, frames, None)
    det.load_state_dict(synthetic.detector_state_dict(arch, frames, out_dims=(2,), taps=det.layer_indices, seed=0))
    det = det.to(cuda_device).eval()
    raw = _frames(clips * frames, 150, 200, seed=5).view(clips, frames, 3, 150, 200)
    m = torch.ones(clips, frames, dtype=torch.bool)
    x_ref = det.transform(raw.flatten(0, 1)).view(clips, frames, 3, 64, 64)        # CPU, torchvision, fp32 normalised
    with torch.no_grad():
        want = det.predict(x_ref.to(cuda_device), m.to(cuda_device))[0][0]
        got = det.predict(raw.to(cuda_device), m.to(cuda_device))[0][0]
        small = det.transform_device(raw.to(cuda_device))
    assert tuple(small.shape) == (clips, frames, 3, 64, 64)
    assert (got - want).abs().max().item() <= 2e-2
    assert torch.equal(got.argmax(-1), want.argmax(-1))

"""Input side: uint8 frames with ConvertImageDtype(float32) + Normalize of the reference's CPU
transform (src/models.py:762-768) fused into the patch-extraction kernel. The fused path must be BIT-IDENTICAL to
feeding the fp32 frames the CPU transform produces."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def cpu_transform_tail(x_u8):
    """T.ConvertImageDtype(torch.float32) then T.Normalize(mean, std), as torchvision evaluates them on the CPU."""
    x = x_u8.to(torch.float32) / 255.0
    mean = torch.as_tensor(MEAN, dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(STD, dtype=torch.float32).view(-1, 1, 1)
    return x.sub_(mean).div_(std)


@pytest.mark.parametrize("res,patch", [(224, 16), (224, 14), (32, 16)])
def test_patchify_u8_is_bit_identical_to_cpu_transform(cuda_device, res, patch):
    from dfdclip_b200 import _native
    g = torch.Generator().manual_seed(3)
    x8 = torch.randint(0, 256, (3, 3, res, res), generator=g, dtype=torch.uint8)
    x8[0, :, :2, :16] = torch.tensor([0, 255] * 8, dtype=torch.uint8)  # extremes
    ref = _native.patchify(cpu_transform_tail(x8).to(cuda_device), patch)
    got = _native.patchify(x8.to(cuda_device), patch, mean=MEAN, std=STD)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


def test_predict_on_uint8_clips_is_bit_identical(cuda_device):
    from dfdclip_b200 import synthetic
    from dfdclip_b200.inference import HostClipPipeline
    from test_parity_gpu import build_detector
    arch, t, b = "small-512x6", 3, 5
    det, _ = build_detector(arch, t, [0, 2, 4], cuda_device)
    res = synthetic.vit_dims(arch)["image_size"]
    g = torch.Generator().manual_seed(17)
    x8 = torch.randint(0, 256, (b, t, 3, res, res), generator=g, dtype=torch.uint8)
    m = torch.ones(b, t, dtype=torch.bool)
    m[1, -1] = False
    xf = cpu_transform_tail(x8.flatten(0, 1)).view(b, t, 3, res, res)
    ref, _ = det.predict(xf.to(cuda_device), m.to(cuda_device))
    got, _ = det.predict(x8.to(cuda_device), m.to(cuda_device))
    torch.cuda.synchronize()
    assert torch.equal(got[0], ref[0])
    pipe = HostClipPipeline(det, chunk_clips=2)
    assert torch.equal(pipe(x8.pin_memory(), m.pin_memory()), ref[0].cpu())
    assert torch.equal(pipe(xf.pin_memory(), m.pin_memory()), ref[0].cpu())  # the fp32 route still works after it


def test_transform_uint8_keeps_pixels(cuda_device):
    """Detector.transform_uint8 = the reference transform without ConvertImageDtype/Normalize; composing it with the
    fused kernel equals Detector.transform (the full reference transform) on frames that need no resize."""
    from dfdclip_b200 import _native
    from test_parity_gpu import build_detector
    det, _ = build_detector("tiny-256x4", 4, [0, 2], cuda_device)
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (2, 3, 32, 32), generator=g, dtype=torch.uint8)
    full = det.transform(frames)
    kept = det.transform_uint8(frames)
    assert kept.dtype == torch.uint8 and torch.equal(kept, frames)
    a = _native.patchify(full.contiguous().to(cuda_device), 16)
    b = _native.patchify(kept.contiguous().to(cuda_device), 16, mean=det.encoder.input_mean, std=det.encoder.input_std)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))

"""Drop-in for the loader API of the reference's ``src/clip/clip.py``: ``available_models()`` (:89-91) and
``load(name, device, jit, download_root) -> (model, preprocess)`` (:94-142). Tokenisation and the JIT-patching
branch (text tower) are out of scope of the DFD-CLIP hot path."""
import hashlib
import os
import urllib.request
import warnings
from typing import List, Union

import torch

from .model import build_model

__all__ = ["available_models", "load", "tokenize"]

_MODELS = {
    "ViT-B/32": "https://openaipublic.azureedge.net/clip/models/40d365715913c9da98579312b702a82c18be219cc2a73407c4526f58eba950af/ViT-B-32.pt",
    "ViT-B/16": "https://openaipublic.azureedge.net/clip/models/5806e77cd80f8b59890b7e101eabd078d9fb84e6937f9e85e4ecb61988df416f/ViT-B-16.pt",
    "ViT-L/14": "https://openaipublic.azureedge.net/clip/models/b8cca3fd41ae0c99ba7e8951adf17d267cdb84cd88be6f7c2e0eca1737a03836/ViT-L-14.pt",
}


def available_models() -> List[str]:
    """Names of the CLIP ViT models this loader knows. The ResNet towers and ``ViT-L/14@336px`` (577 tokens per
    frame; the attention kernels hold a frame's keys on chip, at most 257) are not supported by the B200 path."""
    return list(_MODELS.keys())


def _download(url: str, root: str) -> str:
    os.makedirs(root, exist_ok=True)
    target = os.path.join(root, os.path.basename(url))
    expected = url.split("/")[-2]
    if os.path.isfile(target):
        with open(target, "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() == expected:
                return target
        warnings.warn("%s exists, but the SHA256 checksum does not match; re-downloading the file" % target)
    try:
        with urllib.request.urlopen(url) as source, open(target, "wb") as out:
            while True:
                chunk = source.read(1 << 20)
                if not chunk:
                    break
                out.write(chunk)
    except Exception as exc:  # no network on the build/bench boxes
        raise RuntimeError("cannot download %s (%s); pass the path of a local checkpoint to clip.load()" % (url, exc))
    with open(target, "rb") as fh:
        if hashlib.sha256(fh.read()).hexdigest() != expected:
            raise RuntimeError("Model has been downloaded but the SHA256 checksum does not not match")
    return target


def _transform(n_px):
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        CenterCrop(n_px),
        lambda image: image.convert("RGB"),
        ToTensor(),
        Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu",
         jit: bool = False, download_root: str = None):
    """Load a CLIP model by registry name or from a checkpoint path (a TorchScript archive or a plain
    ``torch.save(state_dict)`` file). Returns ``(model, preprocess)``; only ``model.visual`` is functional."""
    if name.startswith("synthetic:"):
        # "synthetic:<arch>[:<seed>]": seeded random-init checkpoint (no network for the real ones), see synthetic.py
        from .. import synthetic
        parts = name.split(":")
        state_dict = synthetic.clip_checkpoint_state_dict(parts[1], int(parts[2]) if len(parts) > 2 else 0)
        model = build_model(state_dict).to(device)
        return model, _transform(model.visual.input_resolution)
    if name in _MODELS:
        model_path = _download(_MODELS[name], download_root or os.path.expanduser("~/.cache/clip"))
    elif os.path.isfile(name):
        model_path = name
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if jit:
        raise NotImplementedError("jit=True (TorchScript inference) is not supported by dfdclip_b200")
    try:
        state_dict = torch.jit.load(model_path, map_location="cpu").eval().state_dict()
    except RuntimeError:
        state_dict = torch.load(model_path, map_location="cpu")
        if hasattr(state_dict, "state_dict"):
            state_dict = state_dict.state_dict()
    model = build_model(state_dict).to(device)
    return model, _transform(model.visual.input_resolution)


def tokenize(*args, **kwargs):
    raise NotImplementedError("the CLIP text path is out of scope of dfdclip_b200")

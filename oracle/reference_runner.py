"""Run the UNMODIFIED reference ``Detector`` (src/models.py:394-780) on the host CPU: from ``/root/reference`` in the
build container (golden-vector generation) or from the copy staged under ``baseline/_ref`` (``stage_reference.py``)
on the GPU box, where it is the CPU arm of ``bench.py`` (``kind: "reference"``).

ORACLE / test infrastructure: only ``oracle/``, ``tests/`` and ``bench.py``'s CPU legs import this; the product package
never does. The stand-ins (yacs, ftfy, accelerator) and the TorchScript parameter-holder route through the reference's
own ``clip.load`` + ``build_model`` are those of ``gen_golden.py``.
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import gen_golden as gg  # noqa: E402
import stage_reference  # noqa: E402
from dfdclip_b200 import synthetic  # noqa: E402


def reference_root():
    """Where the reference's ``src`` package can be imported from, or None (then only the oracle port exists)."""
    if os.path.isfile(os.path.join(gg.REFERENCE, "src", "models.py")):
        return gg.REFERENCE
    return stage_reference.staged_root()


def import_reference(root=None):
    root = root or reference_root()
    if root is None:
        raise RuntimeError("the reference is neither at %s nor staged under baseline/_ref" % gg.REFERENCE)
    gg.install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import src.models as ref_models
    if not os.path.abspath(ref_models.__file__).startswith(os.path.abspath(root)):
        raise RuntimeError("`src.models` resolved to %s, not to the reference under %s" % (ref_models.__file__, root))
    return ref_models


def build_reference_detector(arch, num_frames, proj=None, root=None):
    """The reference's own Detector (default config, out_dim [2], auc_roc loss) holding the seeded synthetic weights:
    encoder through ``clip.load(<TorchScript holder>)`` + ``build_model`` (fp16 rounding included), then the whole
    state dict through the reference's strict ``load_state_dict`` (inference.py:99)."""
    Detector = import_reference(root).Detector
    tmp = tempfile.mkdtemp(prefix="dfd_ref_")
    ckpt = os.path.join(tmp, "clip.pt")
    gg.write_jit_holder(synthetic.clip_checkpoint_state_dict(arch, seed=0), ckpt)
    cfg = Detector.get_default_config()
    cfg.architecture = ckpt
    cfg.out_dim = [2]
    cfg.losses = ["auc_roc"]
    torch.manual_seed(1)
    # the reference's clip.load defaults to device="cuda" when one is visible (src/clip/clip.py:94): this is the CPU arm
    det = Detector(cfg, num_frames, gg.FakeAccelerator()).to("cpu")
    os.remove(ckpt)
    os.rmdir(tmp)
    sd = synthetic.detector_state_dict(arch, num_frames, out_dims=(2,), taps=det.layer_indices, seed=0)
    if proj is not None:
        sd["decoder.proj0x2"] = proj.clone()
    for k, v in det.encoder.state_dict().items():
        assert torch.equal(v, sd["encoder." + k]), "encoder weight mismatch after the reference loader: " + k
    det.load_state_dict(sd, strict=True)
    return det.eval()

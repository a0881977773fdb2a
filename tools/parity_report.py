#!/usr/bin/env python
"""Print the parity margins of the CUDA path against the BASELINE-size reference goldens (tests/golden/reference_*_c2 /
_c4): max |dlogit|, flipped labels, worst per-clip video-feature cosine. `DFD_LN_FUSE=0|1|2 python tools/parity_report.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from helpers import cosine, fullsize_inputs, load_golden  # noqa: E402
from test_fullsize_gpu import build_detector  # noqa: E402

dev = torch.device("cuda:0")
for name in ("vitb16_c2", "vitl14_c4"):
    g = load_golden(name)
    sd, x, m = fullsize_inputs(name, g)
    det = build_detector(g, sd, dev).eval()
    with torch.no_grad():
        logits, feats = det.predict(x.to(dev), m.to(dev), with_video_features=True)
    got = logits[0].cpu().numpy()
    err = np.abs(got - g["logits"]).max()
    flips = int((got.argmax(-1) != g["pred_labels"]).sum())
    cos = min(cosine(feats["video"][b].cpu(), torch.from_numpy(g["video_feature"][b])) for b in range(x.shape[0]))
    print("DFD_LN_FUSE=%s %s: max |dlogit| %.5f (tolerance 2e-2), flipped labels %d of %d (min |margin| %.4f), "
          "worst video-feature cosine %.6f" % (os.environ.get("DFD_LN_FUSE", "default"), name, err, flips, x.shape[0],
                                               np.abs(g["margin"]).min(), cos))

# the small default-configuration goldens and the decoder-mode cases (tests/test_parity_gpu.py, tests/test_modes_gpu.py)
from helpers import golden_inputs  # noqa: E402
from test_modes_gpu import build_mode_detector  # noqa: E402

for name in ("tiny", "small", "vitb16", "vitl14", "tiny_attn_frame", "tiny_attn_tf", "small_attn_temporal", "small_gp_aq",
             "tiny_aug_query"):
    g = load_golden(name)
    sd, x, m = golden_inputs(g)
    det = build_mode_detector(g, dev, sd)
    with torch.no_grad():
        got = det.predict(x.to(dev), m.to(dev))[0][0].cpu().numpy()
    err = np.nanmax(np.abs(got - g["logits"]))
    print("DFD_LN_FUSE=%s %s: max |dlogit| %.5f, reference margins %s" % (
        os.environ.get("DFD_LN_FUSE", "default"), name, err,
        np.round(np.abs(g["logits"][:, 0] - g["logits"][:, 1]), 3).tolist()))

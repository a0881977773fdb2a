"""Shared test helpers: golden-vector loading, the oracle import, and north_star tolerances."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# BASELINE.json north_star tolerances for the bf16 CUDA path against the reference's fp32 path
TOL_FEATURE_COSINE = 0.999   # per-layer features (K/V taps)
TOL_LOGIT_ABS = 2e-2         # clip logits, absolute


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfd_oracle
    return dfd_oracle


def load_golden(name):
    data = np.load(os.path.join(GOLDEN_DIR, "reference_%s.npz" % name))
    out = {k: data[k] for k in data.files}
    out["arch"] = str(out["arch"])
    out["num_frames"] = int(out["num_frames"])
    out["batch"] = int(out["batch"])
    out["layer_indices"] = [int(i) for i in out["layer_indices"]]
    return out


def golden_op_mode(g):
    """op_mode overrides the golden case was generated with (empty for the default configuration)."""
    import ast
    return dict(ast.literal_eval(str(g["op_mode"]))) if "op_mode" in g else {}


def golden_inputs(g):
    """The exact (state_dict, clips, mask) the golden generator fed the reference."""
    from dfdclip_b200 import synthetic
    dims = synthetic.vit_dims(g["arch"])
    adapter = str(g["adapter"]) if "adapter" in g else None
    op = golden_op_mode(g)
    sd = synthetic.detector_state_dict(g["arch"], g["num_frames"], out_dims=(2,), taps=g["layer_indices"], seed=0,
                                       adapter=adapter, adapter_inner=256, aug_query=bool(op.get("aug_query")),
                                       global_prediction=bool(op.get("global_prediction")),
                                       temporal_position=bool(op.get("temporal_position", 1)))
    x, m = synthetic.make_clips(g["batch"], g["num_frames"], dims["image_size"], seed=7)
    assert np.array_equal(m.numpy(), g["mask"])
    return sd, x, m


# full-size cases of oracle/gen_golden_full.py: name -> seed of synthetic.make_varied_clips
FULLSIZE_SEEDS = {"vitb16_c2": 7, "vitl14_c4": 7, "vitb16_c5": 17, "vitb16_c3": 27}


def fullsize_inputs(name, g, clips=None):
    """(state_dict, clips, mask) of a BASELINE-size golden case: varied synthetic clips and the task head stored in
    the fixture (gen_golden_full.py centres it on the reference's own features). ``clips`` = slice of the batch."""
    from dfdclip_b200 import synthetic
    dims = synthetic.vit_dims(g["arch"])
    sd = synthetic.detector_state_dict(g["arch"], g["num_frames"], out_dims=(2,), taps=g["layer_indices"], seed=0)
    sd["decoder.proj0x2"] = torch.from_numpy(g["proj0x2"]).clone()
    video_ids = [int(v) for v in g["video_ids"]] if "video_ids" in g else None
    x, m = synthetic.make_varied_clips(g["batch"], g["num_frames"], dims["image_size"], seed=FULLSIZE_SEEDS[name],
                                       video_ids=video_ids)
    assert np.array_equal(m.numpy(), g["mask"])
    if clips is not None:
        x, m = x[clips], m[clips]
    return sd, x, m


def golden_tensor(g, key, layer, like):
    """Golden values for tap `key` of `layer`: (reference values, matching values taken from `like`)."""
    like = like.contiguous().float().cpu()
    full = "%s_%d" % (key, layer)
    if full in g:
        return torch.from_numpy(g[full]), like
    idx = torch.from_numpy(g["idx_%s_%d" % (key, layer)])
    return torch.from_numpy(g["val_%s_%d" % (key, layer)]), like.flatten()[idx]


def cosine(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()

"""Race hunt: the same small workload over and over through every caller path; any run whose scores differ from the
first run of the same path is reported (all paths are deterministic by construction)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_parity_gpu import build_detector  # noqa: E402
from dfdclip_b200 import synthetic  # noqa: E402
from dfdclip_b200.inference import HostClipStream, score_videos, score_videos_batched  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
arch, t = "small-512x6", 3
det, _ = build_detector(arch, t, [0, 2, 4], dev)
res = synthetic.vit_dims(arch)["image_size"]
counts = [5, 1, 9, 0, 4]
x, m = synthetic.make_clips(sum(counts), t, res, seed=23)
videos, masks, s = [], [], 0
for n in counts:
    videos.append(x[s:s + n])
    masks.append(m[s:s + n])
    s += n
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
xd, md = x.to(dev), m.to(dev)
pinned = (x.pin_memory(), m.pin_memory())


def path_predict_overlap():
    os.environ["DFD_OVERLAP"] = "1"
    return torch.cat([det.predict(xd[i:i + 4], md[i:i + 4])[0][0] for i in range(0, xd.shape[0], 4)])


def path_predict_two_call():
    os.environ["DFD_OVERLAP"] = "0"
    out = torch.cat([det.predict(xd[i:i + 4], md[i:i + 4])[0][0] for i in range(0, xd.shape[0], 4)])
    os.environ["DFD_OVERLAP"] = "1"
    return out


def path_score_videos():
    return torch.nan_to_num(score_videos(lambda a, b: det.predict(a, b)[0][0], videos, masks, chunk_clips=4, device=dev))


def path_batched():
    return torch.nan_to_num(score_videos_batched(det, videos, masks, batch_clips=7))


def path_stream(overlap):
    def run():
        pipe = HostClipStream(det, overlap_decoder=overlap)
        return torch.cat(list(pipe.run((pinned[0][i:i + 7], pinned[1][i:i + 7]) for i in range(0, x.shape[0], 7))))
    return run


paths = {"predict_overlap": path_predict_overlap, "predict_two_call": path_predict_two_call,
         "score_videos": path_score_videos, "score_videos_batched": path_batched,
         "stream_overlap": path_stream(True), "stream_serial": path_stream(False)}
ref = {k: f().clone().cpu() for k, f in paths.items()}
bad = {k: 0 for k in paths}
for it in range(iters):
    for k, f in paths.items():
        got = f().cpu()
        if not torch.equal(got, ref[k]):
            bad[k] += 1
            if bad[k] <= 3:
                d = (got - ref[k]).abs()
                print("MISMATCH", k, "iter", it, "max", d.max().item(), "rows", d.amax(-1).nonzero().flatten().tolist())
print("iters", iters, "mismatches", bad)
print("predict_overlap == two_call:", torch.equal(ref["predict_overlap"], ref["predict_two_call"]))

"""Diagnostic (GPU): error budget of the patch-mask + adapter case against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from helpers import load_golden, golden_inputs, cosine
from test_modes_gpu import build_mode_detector
import dfd_oracle as O
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
case = sys.argv[1] if len(sys.argv) > 1 else "small_pm_sample"
g = load_golden(case); sd, x, m = golden_inputs(g)
det = build_mode_detector(g, dev, sd)
b, t = x.shape[:2]
li = g["layer_indices"]
pi = [torch.from_numpy(i) for i in g["patch_indices"]]
adapter = str(g["adapter"]) if "adapter" in g else None
np.random.seed(1234)
logits, feats = det.predict(x.to(dev), m.to(dev), with_video_features=True, train=True)
print("gpu logits", logits[0].cpu().tolist(), "golden", g["logits"].tolist())
qkv, _ = det.encoder.encode(x.flatten(0, 1).to(dev), keep_layers=li)
raw = [{n: kv[n].float().cpu() for n in kv} for kv in det.taps_from_qkv(qkv, b, t)]
enc = O.encoder_forward(sd, x.flatten(0, 1), num_layers=max(li) + 1)
for i, l in enumerate(li):
    for n in ("k", "v"):
        ref = enc[l][n][:, 1:].unflatten(0, (b, t))
        print("tap", l, n, "rel err %.5f" % ((raw[i][n] - ref).norm() / ref.norm()).item())
def tail(kvs, label):
    kvs = [{n: kv[n][:, :, idx] for n in kv} for kv, idx in zip(kvs, pi)]
    if adapter: kvs = O.adapter_forward(sd, kvs, adapter)
    rawl, feat, _ = O.decoder_forward(sd, kvs, m, (2,), layer_indices=li)
    print(label, "raw", rawl[0].tolist(), "norm err", np.abs(O.normalise_logits(rawl)[0].numpy() - g["logits"]).max())
tail(raw, "oracle tail on GPU raw taps:")
tail([{n: enc[l][n][:, 1:].unflatten(0, (b, t)) for n in ("k", "v")} for l in li], "oracle tail on oracle taps:")

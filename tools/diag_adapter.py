"""Diagnostic (GPU): error budget of the native adapter path against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from helpers import load_golden, golden_inputs, cosine
from test_adapter_gpu import build_adapter_detector
from test_parity_gpu import build_detector
import dfd_oracle as O
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
for case in sys.argv[1:] or ["tiny_ad_ln", "tiny_ad_nln", "tiny_ad_xxx", "vitb16_ad_z0"]:
    g = load_golden(case); sd, x, m = golden_inputs(g); struct = str(g["adapter"])
    det, _ = build_adapter_detector(g["arch"], g["num_frames"], struct, dev, sd)
    b, t = x.shape[:2]
    logits, feats = det.predict(x.to(dev), m.to(dev), with_video_features=True, with_adapt_features=True)
    err = np.abs(logits[0].cpu().numpy() - g["logits"])
    print(case, "logit err per clip", err.max(-1), "ref", g["logits"].tolist())
    # raw taps from the GPU encoder, adapter in fp32 by the oracle on those taps
    qkv, _ = det.encoder.encode(x.flatten(0, 1).to(dev), keep_layers=det.layer_indices)
    raw = det.taps_from_qkv(qkv, b, t)
    raw_cpu = [{n: kv[n].float().cpu() for n in kv} for kv in raw]
    ref_ad = O.adapter_forward(sd, raw_cpu, struct)
    for i, kv in enumerate(feats["adapt"]):
        for n in ("k", "v"):
            got = kv[n].cpu()
            rel = ((got - ref_ad[i][n]).norm() / ref_ad[i][n].norm()).item()
            relf = ((got - raw_cpu[i][n]) - (ref_ad[i][n] - raw_cpu[i][n])).norm() / (ref_ad[i][n] - raw_cpu[i][n]).norm()
            print("  tap %d %s: rel err of adapted tap %.5f, of the adapter delta %.5f, |delta|/|tap| %.3f" % (
                i, n, rel, relf.item(), ((ref_ad[i][n] - raw_cpu[i][n]).norm() / raw_cpu[i][n].norm()).item()))
    # decoder (oracle, fp32) on GPU-adapted taps vs on oracle-adapted GPU taps
    l1, _, _ = O.decoder_forward(sd, [{n: kv[n].cpu() for n in kv} for kv in feats["adapt"]], m, (2,))
    l2, _, _ = O.decoder_forward(sd, ref_ad, m, (2,))
    l1, l2 = O.normalise_logits(l1)[0].numpy(), O.normalise_logits(l2)[0].numpy()
    print("  oracle decoder on GPU-adapted taps vs golden:", np.abs(l1 - g["logits"]).max(),
          "| on fp32-adapted GPU raw taps vs golden:", np.abs(l2 - g["logits"]).max())

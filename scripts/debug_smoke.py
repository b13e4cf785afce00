import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pixelpick_b200 import _lib
from oracle import acq_oracle as orc
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n, C, H, W = 2, 19, 64, 128
logits = (torch.randn((n, C, H, W), generator=g) * 3).float()
for strat in ("entropy", "least_confidence", "margin_sampling"):
    score = _lib.acq_score(logits.to(dev), strat).cpu()
    ref = orc.uncertainty(orc.probabilities(logits), strat)
    err = (score - ref).abs(); tol = 2e-6 + 1e-5 * ref.abs()
    i = int((err / tol).argmax())
    print(strat, "max err", err.max().item(), "viol", int((err > tol).sum()), "worst: got", score.flatten()[i].item(), "ref", ref.flatten()[i].item(),
          "nan got/ref", int(score.isnan().sum()), int(ref.isnan().sum()))
    if strat == "entropy":
        b, rem = divmod(i, H * W); y, x = divmod(rem, W)
        v = logits[b, :, y, x].double(); p = torch.softmax(v, 0); print("   f64 entropy", float(-(p * p.log()).sum()), "logits", [round(float(t), 3) for t in v])

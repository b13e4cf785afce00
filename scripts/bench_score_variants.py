"""GPU experiment: time the scoring kernel variants (PP_SCORE_VARIANT) and the top-k stages in isolation."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelpick_b200 import _lib
from bench import synth, C, H, W, K_TOP

dev = torch.device("cuda:0")
B = 256
logits, lab, void = synth(B, 1, device=dev)
score = torch.empty((B, H, W), dtype=torch.float32, device=dev)
ws = _lib.TopKWorkspace(B, H * W, K_TOP, dev)
res = {"variant": os.environ.get("PP_SCORE_VARIANT", "0")}
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for strat in ("margin_sampling", "entropy", "least_confidence"):
    for hist in (False, True):
        def f():
            if hist: ws.prepare()
            _lib.acq_score(logits, strat, lab, void, out=score, hist0_ws=ws if hist else None)
        ms = timeit(f)
        res[f"{strat}_hist{int(hist)}"] = {"ms": round(ms, 4), "GBps": round(B * H * W * (C * 4 + 2) / ms / 1e6, 1)}
def topk():
    ws.prepare(); _lib.acq_topk(score.view(B, -1), K_TOP, True, ws=ws, hist0_valid=False)
res["topk_ms_incl_hist0"] = round(timeit(topk), 4)
lb = logits.to(torch.bfloat16)
res["margin_bf16"] = round(timeit(lambda: _lib.acq_score(lb, "margin_sampling", lab, void, out=score)), 4)
print(json.dumps(res))

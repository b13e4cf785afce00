"""GPU: eval forward_lowres latency of DeepLab, fused eval encoder (hand-written conv epilogues) vs module path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixelpick_b200.deeplab import DeepLab

dev = torch.device("cuda:0")
for backbone in ("mobilenet", "resnet"):
    torch.manual_seed(0)
    m = DeepLab(bench.MARGS, backbone=backbone).to(dev).eval()
    for B in (1, 4, 16, 64):
        x = torch.randn((B, 3, 256, 512), device=dev)
        res = {}
        for fused in (True, False):
            m.fused_eval_encoder = fused
            with torch.no_grad():
                for _ in range(3):
                    m.forward_lowres(x)
                torch.cuda.synchronize()
                n = 20 if B <= 16 else 8
                t0 = time.perf_counter()
                for _ in range(n):
                    m.forward_lowres(x)
                torch.cuda.synchronize()
                res[fused] = (time.perf_counter() - t0) / n * 1e3
        print(f"{backbone:9s} B={B:3d}: fused {res[True]:7.2f} ms ({B / res[True] * 1e3:7.0f} img/s) | module path {res[False]:7.2f} ms ({B / res[False] * 1e3:7.0f} img/s)")

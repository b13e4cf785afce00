"""GPU: CUDA-event timing of the head's conv / dgrad / wgrad kernels at the bench shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelpick_b200 import _lib
dev = torch.device("cuda:0")

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

for name, B, H, W, cin, cin_pad, cout, k, dil in [("SegmentHead d1", 32, 64, 128, 304, 320, 256, 3, 1), ("SegmentHead d2", 32, 64, 128, 256, 256, 256, 3, 1),
                                                   ("ASPP rn50 3x3", 32, 32, 64, 2048, 2048, 256, 3, 12), ("ASPP rn50 1x1", 32, 32, 64, 2048, 2048, 256, 1, 1),
                                                   ("ASPP mnv2 3x3", 32, 16, 32, 320, 320, 256, 3, 6), ("ASPP proj", 32, 32, 64, 1024, 1024, 256, 1, 1)]:
    x = torch.randn((B, H, W, cin_pad), device=dev).to(torch.bfloat16)
    wt = torch.randn((cout, cin, k, k), device=dev) * 0.02
    w = _lib.pack_conv_weight(wt, cin_pad, 256)
    y = _lib.conv_igemm(x, w, 256, dil=dil)
    gf = 2.0 * B * H * W * cout * cin * k * k / 1e9
    tf = timed(lambda: _lib.conv_igemm(x, w, 256, dil=dil))
    tw = timed(lambda: _lib.conv_wgrad(x, cin, y, 256, k * k, dil))
    print(f"{name:16s} B={B} {gf:7.1f} GF: fwd {tf:7.1f} us ({gf / tf * 1e3:6.0f} TF/s) | wgrad {tw:7.1f} us ({gf / tw * 1e3:6.0f} TF/s, incl. the zero-fill of dW)")

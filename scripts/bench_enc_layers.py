"""GPU: per-layer inference time of the encoder convolutions — pp_conv_igemm with fused BatchNorm/activation epilogue
(conv_fused) vs the library conv (bf16 channels_last) + fused NHWC BatchNorm kernel of the module path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pixelpick_b200 import _lib

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


# (name, Cin, Cout, k, dil, H, W) at 256x512 input
rn = [("l1.c1", 256, 64, 1, 1, 64, 128), ("l1.c2", 64, 64, 3, 1, 64, 128), ("l1.c3", 64, 256, 1, 1, 64, 128),
      ("l2.c1", 512, 128, 1, 1, 32, 64), ("l2.c2", 128, 128, 3, 1, 32, 64), ("l2.c3", 128, 512, 1, 1, 32, 64),
      ("l3.c1", 1024, 256, 1, 1, 32, 64), ("l3.c2", 256, 256, 3, 2, 32, 64), ("l3.c3", 256, 1024, 1, 1, 32, 64),
      ("l4.c1", 2048, 512, 1, 1, 32, 64), ("l4.c2", 512, 512, 3, 4, 32, 64), ("l4.c3", 512, 2048, 1, 1, 32, 64)]
mn = [("exp16-96", 16, 96, 1, 1, 130, 258), ("proj96-24", 96, 24, 1, 1, 64, 128), ("exp24-144", 24, 144, 1, 1, 66, 130),
      ("proj144-32", 144, 32, 1, 1, 32, 64), ("exp32-192", 32, 192, 1, 1, 34, 66), ("exp64-384", 64, 384, 1, 1, 18, 34),
      ("proj384-96", 384, 96, 1, 1, 16, 32), ("exp96-576", 96, 576, 1, 1, 18, 34), ("exp160-960", 160, 960, 1, 1, 18, 34),
      ("proj960-320", 960, 320, 1, 1, 16, 32)]
print(f"B={B}: layer | GFLOP | ours fused us (TF/s) | library conv us + bn_apply us = total")
for name, ci, co, k, dil, H, W in rn + mn:
    x = torch.randn((B, H, W, ci), device=dev).to(torch.bfloat16)
    w = torch.randn((co, ci, k, k), device=dev) * 0.05
    cp = 32 if co <= 32 else -(-co // 64) * 64
    wp = _lib.pack_conv_weights(w, ci, fwd_pad=(cp, -(-ci // 64) * 64))[0]
    sc, sf = torch.ones(cp, device=dev), torch.zeros(cp, device=dev)
    t_ours = timed(lambda: _lib.conv_fused(x, wp, co, dil=dil, scale=sc, shift=sf, act=1))
    xc = x.permute(0, 3, 1, 2)  # channels_last view
    wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    pad = dil if k == 3 else 0
    y = F.conv2d(xc, wb, padding=pad, dilation=dil)
    t_lib = timed(lambda: F.conv2d(xc, wb, padding=pad, dilation=dil))
    yn = y.permute(0, 2, 3, 1).contiguous()
    out = torch.empty_like(yn)
    if co % 8 == 0:
        t_bn = timed(lambda: _lib.bn_apply(yn, 0, co, sc[:co].contiguous(), sf[:co].contiguous(), 1, out, 0))
    else:
        t_bn = float("nan")
    gf = 2.0 * B * H * W * co * ci * k * k / 1e9
    print(f"  {name:12s} | {gf:7.1f} | {t_ours:7.1f} ({gf / t_ours * 1e3:6.0f}) | {t_lib:7.1f} + {t_bn:6.1f} = {t_lib + t_bn:7.1f}")

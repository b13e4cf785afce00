"""GPU (under ncu): one encoder conv layer through both epilogues of pp_conv_igemm + the library conv, a few launches each.
usage: profile_conv_layer.py Cin Cout k dil H W [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pixelpick_b200 import _lib
from pixelpick_b200.deeplab import _cpad

ci, co, k, dil, H, W = [int(a) for a in sys.argv[1:7]]
B = int(sys.argv[7]) if len(sys.argv) > 7 else 32
dev = torch.device("cuda:0")
x = torch.randn((B, H, W, ci), device=dev).to(torch.bfloat16)
w = torch.randn((co, ci, k, k), device=dev) * 0.05
wp, wd = _lib.pack_conv_weights(w, ci, fwd_pad=(_cpad(co), -(-ci // 64) * 64), dgrad_pad=(_cpad(ci), -(-co // 64) * 64))
stats = torch.zeros((2, co), device=dev)
raw = torch.empty((B, H, W, co), dtype=torch.bfloat16, device=dev)
dy = torch.randn((B, H, W, co), device=dev).to(torch.bfloat16)
ent = [(0, 0, 0)] if k == 1 else [((t // 3 - 1) * dil, (t % 3 - 1) * dil, 0) for t in range(9)]
for _ in range(2):
    _lib.conv_fused(x, wp, co, dil=dil, out=raw, stats=stats)      # TMA-store epilogue + statistics
    _lib.conv_fused(x, wp, co, dil=dil, out=raw)                   # TMA-store epilogue
    prev = _lib.lib().pp_conv_set_epilogue(0)
    _lib.conv_fused(x, wp, co, dil=dil, out=raw)                   # direct-store epilogue
    _lib.lib().pp_conv_set_epilogue(prev)
    _lib.conv_fused(dy, wd, ci, dil=dil)                           # data gradient
    _lib.conv_wgrad_multi(x, ci, dy, co, ent)                      # weight gradient
    xc = x.permute(0, 3, 1, 2)
    wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    F.conv2d(xc, wb, padding=dil if k == 3 else 0, dilation=dil)
torch.cuda.synchronize()

"""GPU: per-layer TRAINING time of the encoder convolutions, ours vs the library path the module encoder uses.
  forward : pp_conv_igemm_stats (raw bf16 + BatchNorm sums in the epilogue)      vs  cuDNN conv (bf16 channels_last)
            [+ old epilogue: direct stores, no statistics]                         [+ the statistics pass it then needs: pp_bn_stats]
  dgrad   : pp_conv_igemm on the flipped / transposed weights                      vs  torch.ops.aten.convolution_backward (input)
  wgrad   : pp_conv_wgrad_multi                                                    vs  torch.ops.aten.convolution_backward (weight)
L2 is flushed between timed launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pixelpick_b200 import _lib
from pixelpick_b200.deeplab import _cpad

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


rn = [("l1.c1", 256, 64, 1, 1, 64, 128), ("l1.c2", 64, 64, 3, 1, 64, 128), ("l1.c3", 64, 256, 1, 1, 64, 128),
      ("l2.c1", 512, 128, 1, 1, 32, 64), ("l2.c2", 128, 128, 3, 1, 32, 64), ("l2.c3", 128, 512, 1, 1, 32, 64),
      ("l3.c1", 1024, 256, 1, 1, 32, 64), ("l3.c2", 256, 256, 3, 2, 32, 64), ("l3.c3", 256, 1024, 1, 1, 32, 64),
      ("l4.c1", 2048, 512, 1, 1, 32, 64), ("l4.c2", 512, 512, 3, 4, 32, 64), ("l4.c3", 512, 2048, 1, 1, 32, 64)]
mn = [("exp16-96", 16, 96, 1, 1, 130, 258), ("proj96-24", 96, 24, 1, 1, 64, 128), ("exp24-144", 24, 144, 1, 1, 66, 130),
      ("proj144-32", 144, 32, 1, 1, 32, 64), ("exp32-192", 32, 192, 1, 1, 34, 66), ("exp64-384", 64, 384, 1, 1, 18, 34),
      ("proj384-96", 384, 96, 1, 1, 16, 32), ("exp96-576", 96, 576, 1, 1, 18, 34), ("exp160-960", 160, 960, 1, 1, 18, 34),
      ("proj960-320", 960, 320, 1, 1, 16, 32)]
which = sys.argv[2] if len(sys.argv) > 2 else "all"
layers = rn if which == "rn" else mn if which == "mn" else rn + mn
print(f"B={B}: layer | GFLOP | fwd ours+stats / ours old-epi / lib conv (+ stats pass) | dgrad ours / lib | wgrad ours / lib   [us]")
tot = {k: 0.0 for k in ("f_ours", "f_old", "f_lib", "f_stats", "d_ours", "d_lib", "w_ours", "w_lib")}
for name, ci, co, k, dil, H, W in layers:
    x = torch.randn((B, H, W, ci), device=dev).to(torch.bfloat16)
    w = torch.randn((co, ci, k, k), device=dev) * 0.05
    wp, wd = _lib.pack_conv_weights(w, ci, fwd_pad=(_cpad(co), -(-ci // 64) * 64), dgrad_pad=(_cpad(ci), -(-co // 64) * 64))
    stats = torch.zeros((2, co), device=dev)
    raw = torch.empty((B, H, W, co), dtype=torch.bfloat16, device=dev)
    f_ours = timed(lambda: _lib.conv_fused(x, wp, co, dil=dil, out=raw, stats=stats))
    prev = _lib.lib().pp_conv_set_epilogue(0)
    f_old = timed(lambda: _lib.conv_fused(x, wp, co, dil=dil, out=raw))
    _lib.lib().pp_conv_set_epilogue(prev)
    xc = x.permute(0, 3, 1, 2)
    wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    pad = dil if k == 3 else 0
    f_lib = timed(lambda: F.conv2d(xc, wb, padding=pad, dilation=dil))
    f_stats = timed(lambda: _lib.bn_stats(raw, 0, co))
    dy = torch.randn((B, H, W, co), device=dev).to(torch.bfloat16)
    d_ours = timed(lambda: _lib.conv_fused(dy, wd, ci, dil=dil))
    ent = [(0, 0, 0)] if k == 1 else [((t // 3 - 1) * dil, (t % 3 - 1) * dil, 0) for t in range(9)]
    w_ours = timed(lambda: _lib.conv_wgrad_multi(x, ci, dy, co, ent))
    dyc = dy.permute(0, 3, 1, 2)
    cb = lambda mask: torch.ops.aten.convolution_backward(dyc, xc, wb, None, [1, 1], [pad, pad], [dil, dil], False, [0, 0], 1, mask)
    d_lib = timed(lambda: cb([True, False, False]))
    w_lib = timed(lambda: cb([False, True, False]))
    gf = 2.0 * B * H * W * co * ci * k * k / 1e9
    for key, v in (("f_ours", f_ours), ("f_old", f_old), ("f_lib", f_lib), ("f_stats", f_stats), ("d_ours", d_ours), ("d_lib", d_lib),
                   ("w_ours", w_ours), ("w_lib", w_lib)):
        tot[key] += v
    print(f"  {name:12s} | {gf:7.1f} | {f_ours:7.1f} / {f_old:7.1f} / {f_lib:7.1f} (+{f_stats:6.1f}) | {d_ours:7.1f} / {d_lib:7.1f} | {w_ours:7.1f} / {w_lib:7.1f}")
print("  totals [us]: " + "  ".join(f"{k} {v:.0f}" for k, v in tot.items()))

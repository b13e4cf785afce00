"""GPU: per-layer timing of the depthwise 3x3 kernels (MobileNetV2 shapes at 256x512) and the NHWC BatchNorm kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelpick_b200 import _lib

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()  # evict L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3  # us


# (C, Hi, Wi, stride, dil) of every depthwise conv at 256x512 (input already padded by fixed_padding)
layers = [(32, 130, 258, 1, 1), (96, 130, 258, 2, 1), (144, 66, 130, 1, 1), (144, 66, 130, 2, 1), (192, 34, 66, 1, 1),
          (192, 34, 66, 2, 1), (384, 18, 34, 1, 1), (576, 18, 34, 1, 1), (960, 18, 34, 1, 1), (960, 20, 36, 1, 2)]
tot = {"fwd": 0, "dgrad": 0, "wgrad": 0}
print(f"depthwise 3x3, B={B}: C Hi Wi s d | MB(in+out) | fwd us (GB/s) | dgrad us (GB/s) | wgrad us (GB/s)")
for C, Hi, Wi, s, d in layers:
    x = torch.randn((B, Hi, Wi, C), device=dev).to(torch.bfloat16)
    w = torch.randn((C, 1, 3, 3), device=dev)
    y = _lib.dwconv_fwd(x, w, s, d)
    dy = torch.randn_like(y)
    mb = (x.numel() + y.numel()) * 2 / 1e6
    tf = timed(lambda: _lib.dwconv_fwd(x, w, s, d))
    td = timed(lambda: _lib.dwconv_dgrad(dy, w, (Hi, Wi), s, d))
    tw = timed(lambda: _lib.dwconv_wgrad(x, dy, s, d))
    tot["fwd"] += tf; tot["dgrad"] += td; tot["wgrad"] += tw
    print(f"  {C:4d} {Hi:3d} {Wi:3d} {s} {d} | {mb:7.1f} | {tf:7.1f} ({mb / tf * 1e3:6.0f}) | {td:7.1f} ({mb / td * 1e3:6.0f}) | {tw:7.1f} ({mb / tw * 1e3:6.0f})")
print("  totals us:", {k: round(v, 1) for k, v in tot.items()})

print(f"BatchNorm NHWC kernels, B={B}: C H W | MB | stats us (GB/s) | apply us (GB/s, r+w) | bwd us (GB/s, 5 passes) | fused fwd us (3 passes) | fused bwd us (5 passes)")
for C, H, W in [(32, 128, 256), (96, 130, 258), (144, 66, 130), (256, 64, 128), (384, 18, 34), (960, 18, 34), (1024, 32, 64), (2048, 32, 64)]:
    x = torch.randn((B, H, W, C), device=dev).to(torch.bfloat16)
    dy = torch.randn_like(x)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    M = B * H * W
    mb = x.numel() * 2 / 1e6
    st = _lib.bn_finalize(_lib.bn_stats(x, 0, C), M, bn)
    out = torch.empty_like(x)
    ts = timed(lambda: _lib.bn_stats(x, 0, C))
    ta = timed(lambda: _lib.bn_apply(x, 0, C, st[0], st[1], 1, out, 0))
    import pixelpick_b200._lib as L
    tb = timed(lambda: _lib.bn_bwd(dy, 0, x, 0, C, st[0], st[1], st[2], st[3], 1))
    tff = timed(lambda: _lib.bn_fwd_fused(x, 0, C, bn, 1, out, 0))
    sc = _lib.bn_scratch(bn, C, dev)
    tfb = timed(lambda: _lib.bn_bwd(dy, 0, x, 0, C, st[0], st[1], st[2], st[3], 1, scratch=sc))
    print(f"  {C:4d} {H:3d} {W:3d} | {mb:7.1f} | {ts:7.1f} ({mb / ts * 1e3:6.0f}) | {ta:7.1f} ({2 * mb / ta * 1e3:6.0f}) | {tb:7.1f} ({5 * mb / tb * 1e3:6.0f})"
          f" | {tff:7.1f} ({3 * mb / tff * 1e3:6.0f}) | {tfb:7.1f} ({5 * mb / tfb * 1e3:6.0f})")

# launch-bound regime: what a CUDA-graph replay pays per BatchNorm layer (forward + backward), separate kernels vs the
# single-launch cooperative kernels
print(f"BatchNorm fwd+bwd per layer inside a CUDA graph (20 layers captured, replayed 5x), B={B}: C H W | separate us | fused us")
for C, H, W in [(96, 130, 258), (144, 66, 130), (192, 34, 66), (384, 18, 34), (960, 18, 34), (256, 64, 128), (64, 16, 32)]:
    x = torch.randn((B, H, W, C), device=dev).to(torch.bfloat16)
    dy = torch.randn_like(x)
    out = torch.empty_like(x)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    M = B * H * W
    res = {}
    for mode in ("separate", "fused"):
        def layer():
            if mode == "fused":
                st = _lib.bn_fwd_fused(x, 0, C, bn, 1, out, 0)
                _lib.bn_bwd(dy, 0, x, 0, C, st[0], st[1], st[2], st[3], 1, scratch=_lib.bn_scratch(bn, C, dev))
            else:
                st = _lib.bn_finalize(_lib.bn_stats(x, 0, C), M, bn)
                _lib.bn_apply(x, 0, C, st[0], st[1], 1, out, 0)
                _lib.bn_bwd(dy, 0, x, 0, C, st[0], st[1], st[2], st[3], 1)
        stream = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            layer()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(20):
                    layer()
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        res[mode] = a.elapsed_time(b) * 1e3 / 100
    print(f"  {C:4d} {H:3d} {W:3d} | {res['separate']:7.1f} | {res['fused']:7.1f}")

"""GPU microbench: achieved HBM bandwidth of the NHWC BatchNorm kernels (algorithmic bytes / CUDA-event time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelpick_b200 import _lib
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for shape in [(32, 64, 128, 256), (32, 32, 64, 2048), (32, 64, 128, 64), (32, 128, 256, 96), (4, 64, 128, 256)]:
    C = shape[-1]
    x = torch.randn(shape, device=dev).to(torch.bfloat16)
    dy = torch.randn(shape, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    nbytes = x.numel() * 2
    bn = torch.nn.BatchNorm2d(C).to(dev)
    M = x.numel() // C
    t_stats = timeit(lambda: _lib.bn_stats(x, 0, C))
    stats = _lib.bn_finalize(_lib.bn_stats(x, 0, C), M, bn)
    t_apply = timeit(lambda: _lib.bn_apply(x, 0, C, stats[0], stats[1], 1, y, 0))
    t_bwd = timeit(lambda: _lib.bn_bwd(dy, 0, x, 0, C, stats[0], stats[1], stats[2], stats[3], 1))
    xt = x.permute(0, 3, 1, 2)
    t_torch = timeit(lambda: torch.nn.functional.batch_norm(xt, None, None, bn.weight, bn.bias, True))
    print(f"{shape}: stats {t_stats*1e3:7.1f} us {nbytes/t_stats/1e6:7.0f} GB/s | apply {t_apply*1e3:7.1f} us {2*nbytes/t_apply/1e6:7.0f} GB/s | "
          f"bwd {t_bwd*1e3:7.1f} us {5*nbytes/t_bwd/1e6:7.0f} GB/s | torch bn fwd {t_torch*1e3:7.1f} us")

"""GPU: where does a train step's time go?  torch.profiler over a few steps (CPU launch overhead vs GPU time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from pixelpick_b200.deeplab import DeepLab
from pixelpick_b200.loss import sparse_cross_entropy

dev = torch.device("cuda:0")
backbone = sys.argv[1] if len(sys.argv) > 1 else "mobilenet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = DeepLab(bench.MARGS, backbone=backbone).to(dev).train()
from pixelpick_b200.graph import make_capturable_adam
opt = make_capturable_adam([{"params": list(model.parameters()), "lr": 5e-4, "weight_decay": 2e-4}])  # optim.FusedAdam
x, y, q = [t.to(dev) for t in bench.synth_train_batch(B, 1)]
def step():
    lr = model.forward_lowres(x)
    loss = sparse_cross_entropy(lr, y, q.bool(), 19)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
print(f"{backbone} B={B}: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms/step wall")
# phases
def timed(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n, out
with torch.no_grad():
    t_enc, (hi, lo) = timed(lambda: model._encode(x))
print(f"  encoder fwd (no grad): {t_enc:.2f} ms")
hi_g, lo_g = hi.detach().requires_grad_(True), lo.detach().requires_grad_(True)
from pixelpick_b200.deeplab import _HeadFn
def head_fb():
    gap = model.aspp.global_avg_pool
    out = _HeadFn.apply(model, 1, hi_g, lo_g, *model._head_params(), gap[1].weight, gap[2].weight, gap[2].bias)
    loss = sparse_cross_entropy(out, y, q.bool(), 19)
    loss.backward()
    return loss
t_head, _ = timed(head_fb)
print(f"  head fwd+bwd (+CE): {t_head:.2f} ms")
def enc_fb():
    h, l = model._encode(x)
    (h.float().mean() + l.float().mean()).backward()
t_encfb, _ = timed(enc_fb)
print(f"  encoder fwd+bwd: {t_encfb:.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = prof.key_averages()
cuda_total = sum(e.self_device_time_total for e in ka) / 3e3
print(f"  profiler: GPU kernel time {cuda_total:.2f} ms/step; kernels/step {sum(e.count for e in ka if e.self_device_time_total > 0 and e.device_type.name == 'CUDA') / 3:.0f}")
rows = sorted(ka, key=lambda e: -e.self_device_time_total)[:45]
for e in rows:
    print(f"    {e.key[:110]:110s} n={e.count / 3:6.1f} gpu {e.self_device_time_total / 3e3:7.3f} ms")
# which torch glue ops move the most bytes?  (grouped by input shape)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
    for _ in range(2): step()
    torch.cuda.synchronize()
glue = [e for e in prof2.key_averages(group_by_input_shape=True)
        if e.key.startswith("aten::") and e.self_device_time_total > 0 and "conv" not in e.key]
print("  torch glue ops by shape (per step):")
for e in sorted(glue, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"    {e.key:34s} n={e.count / 2:5.1f} gpu {e.self_device_time_total / 2e3:7.3f} ms  {str(e.input_shapes)[:150]}")

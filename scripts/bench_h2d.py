import torch, time
for mb in (16, 64, 160, 640):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(f"H2D {mb} MB pinned: {mb * 10 / 1024 / (a.elapsed_time(b) / 1e3):.1f} GiB/s")
# two streams
h1 = torch.empty(320 << 20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(320 << 20, dtype=torch.uint8).pin_memory()
d1 = torch.empty(320 << 20, dtype=torch.uint8, device="cuda"); d2 = torch.empty_like(d1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 2 streams: {640 * 5 / 1024 / dt:.1f} GiB/s")

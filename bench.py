#!/usr/bin/env python
"""bench.py — headline benchmark of the pixelpick_b200 hot paths (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

BASELINE.json metric: "train images/sec + query Mpixels/sec, Cityscapes 256x512" (configs[1]/[2]).  One JSON line:

  primary (metric/value/roofline/e2e/cpu_baseline) = the QUERY hot path, margin_sampling, top-5 % (k=6553), n=10:
      step     = one pass over a batch of B images per GPU: fused softmax+margin+mask score -> per-image sorted
                 top-k -> gather of the n picks                                   [all hand-written CUDA]
      value    = Mpixels/s over all ranks, logits resident in HBM (CUDA events, max over ranks)
      e2e      = the same metric through the C-ABI host-buffer call (pp_acq_session_run_host): pinned host
                 logits/masks -> H2D -> kernels -> D2H of the picks, incl. the host NumPy draw of the pick positions
      roofline = the scoring kernel alone, timed with CUDA events INSIDE the timed steps; algorithmic bytes
                 (C*4 + 2) B/px over MEASURED_PEAKS.json hbm_gbs
  "train"  = images/s of the DeepLabv3+ train step (encoder bf16 channels_last -> tcgen05 ASPP/decoder head ->
             fused upsample+sparse-CE -> custom backward -> fused Adam), dropout on, 10 labelled px/image,
             for MobileNetV2 (configs[1]) and ResNet-50 (configs[2]) at the reference batch (4/GPU) and a
             throughput batch; its own e2e (pinned host batch -> H2D every step, loss read back every step) and the
             tensor-core roofline of the SegmentHead 3x3 conv kernel (nominal FLOPs / CUDA-event time / measured bf16 peak)
  "query_model" = Mpixels/s of QuerySelector-style querying with the model in the loop (forward_lowres + fused
             upsample/score + top-k), images resident in HBM and from pinned host memory.
  cpu_baseline / --impl reference = the oracle port of the reference's CPU path (torch CPU + NumPy, all host
             threads) on a bounded sample of the same workloads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, H, W = 19, 256, 512
STRATEGY = "margin_sampling"
TOP_N_PERCENT, N_SEL = 0.05, 10
K_TOP = int(H * W * TOP_N_PERCENT)
ALG_BYTES_PER_PX = C * 4 + 2  # logits + labelled mask + void mask
WORKLOAD = f"cityscapes 256x512 C={C} {STRATEGY} top-5% (k={K_TOP}) n={N_SEL}: logits -> score -> top-k select -> n random ranks"
MARGS = Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=C)
OPT = {"lr": 5e-4, "weight_decay": 2e-4}  # args.py:101-106 (cs)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d["bf16_tflops"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth(n_img, seed, device=None, pin=False):
    """Synthetic Cityscapes-shape logits N(0, 3^2), 10 labelled px / image, 1 % void."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.empty((n_img, C, H, W), dtype=torch.float32, pin_memory=pin)
    for i in range(0, n_img, 16):  # chunked to bound the host RNG working set
        logits[i:i + 16] = torch.randn((min(16, n_img - i), C, H, W), generator=g) * 3.0
    rs = np.random.RandomState(seed)
    lab = np.zeros((n_img, H * W), dtype=np.uint8)
    for i in range(n_img):
        lab[i, rs.choice(H * W, 10, replace=False)] = 1
    void = (rs.rand(n_img, H * W) < 0.01).astype(np.uint8)
    lab_t, void_t = torch.from_numpy(lab).view(n_img, H, W), torch.from_numpy(void).view(n_img, H, W)
    if pin:
        lab_t, void_t = lab_t.pin_memory(), void_t.pin_memory()
    if device is not None:
        return logits.to(device), lab_t.to(device), void_t.to(device)
    return logits, lab_t, void_t


def synth_train_batch(B, seed, pin=False, size=None):
    H, W = size or (globals()["H"], globals()["W"])
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, 3, H, W), generator=g)
    rs = np.random.RandomState(seed)
    y = torch.from_numpy(rs.randint(0, C, size=(B, H, W)).astype(np.int64))
    y[torch.from_numpy(rs.rand(B, H, W) < 0.01)] = C
    q = np.zeros((B, H * W), dtype=np.uint8)
    for i in range(B):
        q[i, rs.choice(H * W, 10, replace=False)] = 1
    q = torch.from_numpy(q).view(B, H, W)
    if pin:
        x, y, q = x.pin_memory(), y.pin_memory(), q.pin_memory()
    return x, y, q


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle ports of the reference CPU path
# ----------------------------------------------------------------------------------------------
def cpu_query_rate(n_img, reps, threads):
    from oracle import acq_oracle as orc  # checker / baseline only
    torch.set_num_threads(threads)
    logits, lab, void = synth(n_img, 1234)
    lab_np, void_np = lab.numpy().astype(bool), void.numpy().astype(bool)
    imgs = [logits[i:i + 1] for i in range(n_img)]
    names = [f"img_{i:05d}.png" for i in range(n_img)]
    np.random.seed(0)
    orc.query_images(imgs[:2], STRATEGY, lab_np, void_np, names, N_SEL, TOP_N_PERCENT)  # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.query_images(imgs, STRATEGY, lab_np, void_np, names, N_SEL, TOP_N_PERCENT)
        times.append(time.perf_counter() - t0)
    return n_img * H * W / 1e6 / min(times), times


def cpu_train_rate(backbone, B, steps, threads):
    """oracle train step (fp32, torch CPU, Adam) — model.py:103-122 restated; returns images/s."""
    from oracle import deeplab_oracle as dorc
    from pixelpick_b200.deeplab import DeepLab  # parameter names / shapes only
    torch.set_num_threads(threads)
    shapes = {k: tuple(v.shape) for k, v in DeepLab(MARGS, backbone=backbone).state_dict().items()}
    sd = dorc.synthetic_state_dict(shapes, seed=1)
    params = []
    for k in sorted(sd):
        if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var")) and \
                not k.startswith(("backbone.low_level_features.", "backbone.high_level_features.")):
            sd[k] = sd[k].clone().requires_grad_(True)
            params.append(sd[k])
    opt = torch.optim.Adam(params, lr=OPT["lr"], weight_decay=OPT["weight_decay"])
    x, y, q = synth_train_batch(B, 7)
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        out = dorc.deeplab_forward(sd, x, backbone=backbone, training=True, drop=(0.5, 0.5, 0.2))
        loss = dorc.sparse_ce_loss(out["pred"], y, q, C)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if i > 0:
            times.append(time.perf_counter() - t0)
    return B / float(np.mean(times)), times


def cpu_query_model_rate(backbone, n_img, threads):
    """oracle port of query.py:159-212 WITH the model forward (fp32 torch CPU), one image per forward as the reference."""
    from oracle import acq_oracle as orc
    from oracle import deeplab_oracle as dorc
    from pixelpick_b200.deeplab import DeepLab  # parameter names / shapes only
    torch.set_num_threads(threads)
    shapes = {k: tuple(v.shape) for k, v in DeepLab(MARGS, backbone=backbone).state_dict().items()}
    sd = dorc.synthetic_state_dict(shapes, seed=1)
    x, y, q = synth_train_batch(n_img, 21)
    lab_np, void_np = q.numpy().astype(bool), (y == C).numpy()
    names = [f"img_{i:05d}.png" for i in range(n_img)]
    np.random.seed(0)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_img):
            pred = dorc.deeplab_forward(sd, x[i:i + 1], backbone=backbone)["pred"]
            orc.query_images([pred], STRATEGY, lab_np[i:i + 1], void_np[i:i + 1], names[i:i + 1], N_SEL, TOP_N_PERCENT)
    dt = time.perf_counter() - t0
    return n_img * H * W / 1e6 / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_img = 16
    W_, K_ = max(args.warmup, 1), args.steps
    from oracle import acq_oracle as orc
    torch.set_num_threads(threads)
    logits, lab, void = synth(n_img, 1234)
    lab_np, void_np = lab.numpy().astype(bool), void.numpy().astype(bool)
    imgs = [logits[i:i + 1] for i in range(n_img)]
    names = [f"img_{i:05d}.png" for i in range(n_img)]
    np.random.seed(0)
    for _ in range(W_):
        orc.query_images(imgs, STRATEGY, lab_np, void_np, names, N_SEL, TOP_N_PERCENT)
    t0 = time.perf_counter()
    for _ in range(K_):
        orc.query_images(imgs, STRATEGY, lab_np, void_np, names, N_SEL, TOP_N_PERCENT)
    dt = time.perf_counter() - t0
    val = K_ * n_img * H * W / 1e6 / dt
    sample = f"{n_img} images of 256x512x19 fp32 per step (logits pre-computed), torch CPU + NumPy, {threads} threads"
    out = {
        "impl": "reference", "metric": "query_mpixels_per_sec", "value": val, "unit": "Mpixels/s", "n_gpus": world,
        "steps": K_, "warmup": W_, "ms_per_step": dt / K_ * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": n_img},
        "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_train:
        qm, qdt = cpu_query_model_rate("mobilenet", 8, threads)
        out["query_model"] = {"unit": "Mpixels/s", "mobilenetv2_bs1_host": {"value": qm, "images_per_s": 8 / qdt},
                              "sample": "oracle port of query.py:159-212 incl. the MobileNetV2-DeepLab forward (fp32 torch CPU), "
                                        "8 images, one per forward"}
        r_mn, t_mn = cpu_train_rate("mobilenet", 4, 2, threads)
        out["train"] = {"metric": "train_images_per_sec", "unit": "images/s", "dtype": "f32",
                        "mobilenetv2_b4": {"value": r_mn, "ms_per_step": float(np.mean(t_mn)) * 1e3, "steps": len(t_mn)},
                        "sample": "oracle port of model.py:103-122 (fp32, torch CPU, Adam, dropout on), B=4, 2 timed steps"}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def bench_train_graph(backbone, B, steps, warmup, dev, world, e2e=False, loop=False):
    """Same train step captured ONCE as a CUDA graph (pixelpick_b200/graph.py) and replayed: removes the ~1000 kernel
    launches / step of CPU overhead that bound the reference batch size."""
    import torch.distributed as dist
    from pixelpick_b200 import _lib, dist as ppdist
    from pixelpick_b200.deeplab import DeepLab
    from pixelpick_b200.graph import GraphedTrainStep, make_capturable_adam
    torch.manual_seed(0)
    model = DeepLab(MARGS, backbone=backbone).to(dev)
    model.train()
    ppdist.broadcast_parameters(model)
    groups = [{"params": model.backbone.parameters(), "lr": OPT["lr"] / 10, "weight_decay": OPT["weight_decay"]}]
    for part in (model.aspp, model.low_level_conv, model.seg_head):
        groups.append({"params": part.parameters(), "lr": OPT["lr"], "weight_decay": OPT["weight_decay"]})
    opt = make_capturable_adam(groups)
    reducer = ppdist.GradAllReducer(model) if world > 1 else None
    hx, hy, hq = synth_train_batch(B, 11 + ppdist.rank(), pin=True)
    # loop=True is what pixelpick_b200.Model._train_epoch runs: host batch -> H2D every step (prefetched during the
    # previous step's graph), running metrics accumulated on the device, ONE device->host read at the end of the epoch
    gs = GraphedTrainStep(model, opt, (B, H, W), C, capacity=B * 16, device=dev, reducer=reducer,
                          n_classes=C if loop else None)
    gs.load(hx, hy, hq)
    gs.capture()
    loss_host = torch.zeros(1).pin_memory()

    if e2e or loop:
        gs.prefetch(hx, hy, hq)

    def step():
        if e2e or loop:
            gs.commit()  # staging -> static inputs (the H2D of this batch ran during the previous step's graph)
        loss, _ = gs()
        if e2e or loop:
            gs.prefetch(hx, hy, hq)  # next host batch -> staging on the copy stream, overlapping the graph (H2D every step)
        if e2e:
            loss_host.copy_(loss.detach().reshape(1))  # D2H read of this step's loss every step
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    l0 = _lib.lib().pp_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        last = step()
    if loop:
        gs.metrics.read()  # the epoch's confusion matrix + loss sum: the loop's only device->host read
    b.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = a.elapsed_time(b) if not (e2e or loop) else wall * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    host = e2e or loop
    return {"value": world * B * steps / (ms / 1e3), "ms_per_step": ms / steps, "batch_per_gpu": B, "steps": steps,
            "cuda_graph": True, "final_loss": float(last.item()),
            "h2d_bytes_per_step": B * (3 * H * W * 4) + 3 * B * 16 * 4 + 4 if host else 0,
            "d2h_bytes_per_step": 4 if e2e else (C * C * 8 + 16) / steps if loop else 0}


def bench_train(backbone, B, steps, warmup, dev, world, e2e=False):
    """images/s of the full train step; returns dict(value, ms_per_step, ...)."""
    import torch.distributed as dist
    from pixelpick_b200 import _lib, dist as ppdist
    from pixelpick_b200.deeplab import DeepLab
    from pixelpick_b200.loss import sparse_cross_entropy
    torch.manual_seed(0)
    model = DeepLab(MARGS, backbone=backbone).to(dev)
    model.train()
    ppdist.broadcast_parameters(model)
    groups = [{"params": model.backbone.parameters(), "lr": OPT["lr"] / 10, "weight_decay": OPT["weight_decay"]}]
    for part in (model.aspp, model.low_level_conv, model.seg_head):
        groups.append({"params": part.parameters(), "lr": OPT["lr"], "weight_decay": OPT["weight_decay"]})
    opt = torch.optim.Adam(groups, fused=True)
    reducer = ppdist.GradAllReducer(model) if world > 1 else None
    hx, hy, hq = synth_train_batch(B, 11 + ppdist.rank(), pin=True)
    dx, dy, dq = hx.to(dev), hy.to(dev), hq.to(dev)
    loss_host = torch.zeros(1).pin_memory()

    def step(from_host):
        if from_host:
            x, y, q = hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True), hq.to(dev, non_blocking=True)
        else:
            x, y, q = dx, dy, dq
        lowres = model.forward_lowres(x)
        loss = sparse_cross_entropy(lowres, y, q.bool(), C)
        opt.zero_grad(set_to_none=True)
        if world > 1:
            n_local = torch.tensor(float(B * 10), device=dev)
            (loss * ppdist.global_mean_loss_scale(n_local)).backward()
            reducer()
        else:
            loss.backward()
        opt.step()
        if from_host:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=False)  # D2H read of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(e2e)
    barrier()
    l0 = _lib.lib().pp_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        last = step(e2e)
    b.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = a.elapsed_time(b) if not e2e else wall * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": world * B * steps / (ms / 1e3), "ms_per_step": ms / steps, "batch_per_gpu": B, "steps": steps,
            "our_kernel_launches_per_step": (_lib.lib().pp_launch_count() - l0) / steps, "final_loss": float(last.item()),
            "h2d_bytes_per_step": B * (3 * H * W * 4 + H * W * 8 + H * W) if e2e else 0, "d2h_bytes_per_step": 4 if e2e else 0}


def bench_conv_roofline(dev, peak_tf):
    """SegmentHead conv #1 (3x3, 304(320)->256) at B=32, 64x128: nominal FLOPs / CUDA-event time."""
    from pixelpick_b200 import _lib
    B = 32
    x = torch.randn((B, 64, 128, 320), device=dev).to(torch.bfloat16)
    w = _lib.pack_conv_weight(torch.randn((256, 304, 3, 3), device=dev) * 0.02, 320, 256)
    sc, sf = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    out = torch.empty((B, 64, 128, 256), dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        _lib.conv_igemm(x, w, 256, scale=sc, shift=sf, relu=True, out=out)
    torch.cuda.synchronize()
    n = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        _lib.conv_igemm(x, w, 256, scale=sc, shift=sf, relu=True, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    flops = 2.0 * B * 64 * 128 * 256 * 9 * 304
    ach = flops / (ms / 1e3) / 1e12
    return {"kernel": "conv_igemm_kernel<256> (SegmentHead 3x3 304->256, B=32, 64x128, BN+ReLU epilogue)", "bound": "tensor",
            "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "kernel_ms": ms,
            "nominal_flops_per_launch": flops, "traffic": None}


def bench_query_model(backbone, n_img, steps, dev, from_host, size=None):
    """QuerySelector-style querying with the model in the loop (eval forward_lowres + fused upsample/score + top-k).
    n_img = 1 is the reference's own query batch (model.py:36-37: the query dataloader has batch_size 1)."""
    from pixelpick_b200 import _lib
    from pixelpick_b200.deeplab import DeepLab
    H, W = size or (globals()["H"], globals()["W"])
    K_TOP = int(H * W * TOP_N_PERCENT)
    torch.manual_seed(0)
    model = DeepLab(MARGS, backbone=backbone).to(dev).eval()
    hx, hy, hq = synth_train_batch(n_img, 21, pin=True, size=(H, W))
    hvoid = (hy == C).pin_memory()
    dx, dq, dvoid = hx.to(dev), hq.to(dev), hvoid.to(dev)
    ws = _lib.TopKWorkspace(n_img, H * W, K_TOP, dev)
    pos = torch.from_numpy(np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(n_img)]).astype(np.int32)).to(dev)
    out_host = torch.empty((n_img, N_SEL), dtype=torch.int32).pin_memory()

    def step():
        with torch.no_grad():
            if from_host:
                x, q, v = hx.to(dev, non_blocking=True), hq.to(dev, non_blocking=True), hvoid.to(dev, non_blocking=True)
            else:
                x, q, v = dx, dq, dvoid
            lr = model.forward_lowres(x)
            ws.prepare()
            score = _lib.acq_score_upsampled(lr, (H, W), STRATEGY, q, v, hist0_ws=ws)
            sel = _lib.acq_select_pick(score.view(n_img, -1), K_TOP, False, pos, ws=ws, hist0_valid=True)
            if from_host:
                out_host.copy_(sel)
        return sel

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n_img * H * W * steps / 1e6 / dt, "images_per_s": n_img * steps / dt, "ms_per_step": dt / steps * 1e3,
            "images_per_step": n_img}


def bench_strategy_sweep(dev, hbm_peak, n_img=8, Hs=1024, Ws=2048, steps=5):
    """BASELINE configs[4]: Cityscapes 1024x2048 query sweep over entropy / margin (= BvSB) / least-confidence on logits
    resident in HBM: scoring-kernel GB/s (algorithmic bytes, CUDA events) and whole-step Mpixels/s (k = 5 % = 104857)."""
    from pixelpick_b200 import _lib
    HWs = Hs * Ws
    k = int(HWs * TOP_N_PERCENT)
    g = torch.Generator().manual_seed(5)
    logits = torch.empty((n_img, C, Hs, Ws), dtype=torch.float32, device=dev)
    for i in range(n_img):
        logits[i] = (torch.randn((C, Hs, Ws), generator=g) * 3.0).to(dev)
    rs = np.random.RandomState(5)
    lab = torch.from_numpy((rs.rand(n_img, Hs, Ws) < 100.0 / HWs).astype(np.uint8)).to(dev)
    void = torch.from_numpy((rs.rand(n_img, Hs, Ws) < 0.01).astype(np.uint8)).to(dev)
    ws = _lib.TopKWorkspace(n_img, HWs, k, dev)
    score = torch.empty((n_img, Hs, Ws), dtype=torch.float32, device=dev)
    pos = torch.from_numpy(np.stack([rs.permutation(k)[:N_SEL] for _ in range(n_img)]).astype(np.int32)).to(dev)
    out = {"workload": f"cityscapes {Hs}x{Ws} C={C} top-5% (k={k}) n={N_SEL}, {n_img} images / step, logits in HBM "
                       f"({n_img * C * HWs * 4 / 1e6:.0f} MB > L2)"}
    for strat in ("entropy", "margin_sampling", "least_confidence"):
        largest = _lib.LARGEST[strat]

        def step(ea=None, eb=None):
            ws.prepare()
            if ea is not None:
                ea.record()
            _lib.acq_score(logits, strat, lab, void, out=score, hist0_ws=ws)
            if eb is not None:
                eb.record()
            return _lib.acq_select_pick(score.view(n_img, HWs), k, largest, pos, ws=ws, hist0_valid=True)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for ea, eb in ev:
            step(ea, eb)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        sms = float(np.mean([x.elapsed_time(y) for x, y in ev]))
        ach = n_img * HWs * ALG_BYTES_PER_PX / (sms / 1e3) / 1e9
        out[strat] = {"step_mpixels_per_sec": n_img * HWs / 1e6 / (ms / 1e3), "ms_per_step": ms, "score_kernel_ms": sms,
                      "score_kernel_gbs": ach, "score_kernel_frac_of_hbm_peak": ach / hbm_peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per step per GPU (device-resident query leg)")
    ap.add_argument("--e2e-batch", type=int, default=64, help="images per step per GPU (host-buffer query leg)")
    ap.add_argument("--train-batch", type=int, default=32, help="throughput batch of the train leg (per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sorted-topk", action="store_true", help="query step materialises the sorted top-k list (pp_acq_topk)")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--overlap-select", action="store_true",
                    help="EXPERIMENTAL (written in round 1 without GPU time left to verify it; default off): run the "
                         "latency-bound select + pick of step i on a high-priority side stream while step i+1 is scored")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from pixelpick_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout = ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = _lib.lib()
    hbm_peak, tf_burst, tf_sust, peak_src = peaks()

    B, K, Wm = args.batch, args.steps, args.warmup
    HW = H * W
    largest = _lib.LARGEST[STRATEGY]
    logits, lab, void = synth(B, 100 + rank, device=dev)
    ws = _lib.TopKWorkspace(B, HW, K_TOP, dev)
    score = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    np.random.seed(rank)
    pos = torch.from_numpy(np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(B)]).astype(np.int32)).to(dev)
    # multi-GPU: the path's ONE exchange is an all-gather of the per-rank picks per query ROUND: the K
    # timed steps are one round, so every step parks its picks in sel_round and the round ends with one all_gather
    sel_round = torch.empty((K, B, N_SEL), dtype=torch.int32, device=dev) if world > 1 else None
    gathered = [torch.empty((K, B, N_SEL), dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]

    if args.overlap_select and not args.sorted_topk:
        # two slots (workspace + score map); slot p's select + pick run on `side` behind an event while the main stream
        # already scores the next batch into slot p ^ 1.  DESIGN.md §7 item 1.
        side = torch.cuda.Stream(device=dev, priority=-1)
        slots = [(ws, score), (_lib.TopKWorkspace(B, HW, K_TOP, dev), torch.empty_like(score))]
        scored = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        turn = [0]

        def step(i=None):
            p = turn[0] & 1
            turn[0] += 1
            ws_p, score_p = slots[p]
            main = torch.cuda.current_stream()
            main.wait_event(free[p])  # the select + pick that used this slot two steps ago (no-op before its first record)
            ws_p.prepare()
            if i is not None:
                ev_a[i].record()
            _lib.acq_score(logits, STRATEGY, lab, void, out=score_p, hist0_ws=ws_p)
            if i is not None:
                ev_b[i].record()
            scored[p].record(main)
            with torch.cuda.stream(side):
                side.wait_event(scored[p])
                sel = _lib.acq_select_pick(score_p.view(B, HW), K_TOP, largest, pos, ws=ws_p, hist0_valid=True)
                if world > 1 and i is not None:
                    sel_round[i].copy_(sel)
                free[p].record(side)
            return sel
    else:
        side = None

    def step_serial(i=None):
        ws.prepare()
        if i is not None:
            ev_a[i].record()
        _lib.acq_score(logits, STRATEGY, lab, void, out=score, hist0_ws=ws)
        if i is not None:
            ev_b[i].record()
        if args.sorted_topk:  # materialise the whole sorted top-k list, then gather the drawn ranks
            topk = _lib.acq_topk(score.view(B, HW), K_TOP, largest, ws=ws, hist0_valid=True)
            sel = _lib.acq_gather(topk, pos)
        else:  # what QuerySelector runs: only the n drawn ranks are needed (query.py:63-64) -> radix pick, no sort
            sel = _lib.acq_select_pick(score.view(B, HW), K_TOP, largest, pos, ws=ws, hist0_valid=True)
        if world > 1 and i is not None:
            sel_round[i].copy_(sel)
        return sel

    if side is None:
        step = step_serial

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = lib.pp_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        step(i)
    if side is not None:  # the last steps' select + pick belong to the timed region
        torch.cuda.current_stream().wait_stream(side)
    if world > 1:  # the round's exchange, inside the timed region: per-rank picks -> every rank (rank 0 builds the dict)
        dist.all_gather(gathered, sel_round)
    t_end.record()
    barrier()
    launches = lib.pp_launch_count() - l0
    ms_total = t_start.elapsed_time(t_end)
    score_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]))

    # ---- e2e: host buffers through the C-ABI session ------------------------------------------
    Be = args.e2e_batch
    h_logits, h_lab, h_void = synth(Be, 500 + rank, pin=True)
    # one chunk per call: on this pod every host<->device copy pays ~0.9 ms before its first byte (scripts/bench_h2d.py),
    # so one 650 MB logits copy (48.7 GiB/s) beats four 160 MB ones (33 GiB/s); the kernels take ~0.15 ms
    sess = _lib.AcqSession(Be, C, H, W, K_TOP, N_SEL)
    h_pos = torch.empty((Be, N_SEL), dtype=torch.int32).pin_memory()
    h_sel = torch.empty((Be, N_SEL), dtype=torch.int32).pin_memory()

    def e2e_step():
        sess.begin(h_logits, h_lab, h_void, STRATEGY)  # H2D + score + select in flight ...
        h_pos.numpy()[:] = np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(Be)])  # ... while the host draws
        sess.finish(h_pos, h_sel)
        return h_sel

    Ke = max(3, min(K, 10))
    for _ in range(3):
        e2e_step()
    barrier()
    le0 = lib.pp_launch_count()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = lib.pp_launch_count() - le0
    barrier()
    sess.close()
    del logits, score, ws, h_logits
    torch.cuda.empty_cache()

    t = torch.tensor([ms_total, score_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, score_ms, e2e_s = [float(v) for v in t.cpu()]

    # ---- train leg + model-in-the-loop query -------------------------------------------------------
    train = None
    qmodel = None
    if not args.no_train:
        train = {"metric": "train_images_per_sec", "unit": "images/s", "dtype": "bf16 (fp32 accumulate, fp32 master weights)",
                 "config": "cityscapes 256x512, 10 labelled px/image, Adam lr 5e-4 (encoder lr/10) wd 2e-4, dropout on; "
                           "synthetic batch resident in HBM (value) / pinned host -> H2D every step + loss D2H (e2e)"}
        ts = max(5, min(K, 20))
        for name, bb in (("mobilenetv2", "mobilenet"), ("resnet50", "resnet")):
            train[f"{name}_b4"] = bench_train(bb, 4, ts, 5, dev, world)
            train[f"{name}_b4_graph"] = bench_train_graph(bb, 4, ts, 5, dev, world)
            train[f"{name}_b4_graph_e2e"] = bench_train_graph(bb, 4, ts, 5, dev, world, e2e=True)
            train[f"{name}_b4_graph_loop"] = bench_train_graph(bb, 4, ts, 5, dev, world, loop=True)
            train[f"{name}_b{args.train_batch}"] = bench_train(bb, args.train_batch, ts, 3, dev, world)
            train[f"{name}_b{args.train_batch}_graph"] = bench_train_graph(bb, args.train_batch, ts, 3, dev, world)
            train[f"{name}_b{args.train_batch}_graph_e2e"] = bench_train_graph(bb, args.train_batch, ts, 3, dev, world, e2e=True)
            torch.cuda.empty_cache()
        if rank == 0:
            train["roofline_tensor"] = bench_conv_roofline(dev, tf_burst)
            train["roofline_tensor"]["peak_source"] = peak_src + " bf16_tflops (burst; kernel timed alone)"
            qmodel = {"unit": "Mpixels/s",
                      "mobilenetv2_hbm": bench_query_model("mobilenet", 64, 5, dev, False),
                      "mobilenetv2_host": bench_query_model("mobilenet", 64, 5, dev, True),
                      "resnet50_hbm": bench_query_model("resnet", 32, 5, dev, False),
                      "resnet50_host": bench_query_model("resnet", 32, 5, dev, True),
                      "mobilenetv2_bs1_host": bench_query_model("mobilenet", 1, 50, dev, True),
                      "resnet50_bs1_host": bench_query_model("resnet", 1, 50, dev, True),
                      "resnet50_1024x2048_bs1_host": bench_query_model("resnet", 1, 10, dev, True, size=(1024, 2048)),
                      "note": "bs1 = the reference's query loop (one image per forward, query.py:159-212); "
                              "hbm/host = 64 (MobileNetV2) / 32 (ResNet-50) images per step"}
    barrier()
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        value = world * B * HW * K / 1e6 / (ms_total / 1e3)
        achieved = B * HW * ALG_BYTES_PER_PX / (score_ms / 1e3) / 1e9
        e2e_val = world * Be * HW * Ke / 1e6 / e2e_s
        h2d = Be * (C * HW * 4 + 2 * HW + N_SEL * 4)
        d2h = Be * N_SEL * 4
        out = {
            "metric": "query_mpixels_per_sec", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step_per_gpu": B, "e2e_images_per_step_per_gpu": Be,
                       "selection": "sorted top-k list + gather" if args.sorted_topk else "radix select + order statistics at the drawn ranks (no sort; identical picks)"
                                    + ("; select + pick of step i overlapped with the scoring of step i+1 (side stream)" if side is not None else ""),
                       "l2": f"inputs larger than L2 ({B * C * HW * 4 / 1e6:.0f} MB of logits per step)",
                       "parallelism": f"images sharded over {world} rank(s); one all_gather of the round's picks "
                                      f"({K} steps = one query round)" if world > 1 else "single GPU"},
            "roofline": {"kernel": "acq_score_vec_kernel<19, margin, f32, fused hist0>", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "peak_source": peak_src, "kernel_ms": score_ms,
                         "alg_bytes_per_launch": B * HW * ALG_BYTES_PER_PX,
                         "share_of_step": score_ms / (ms_total / K)},
            "e2e": {"value": e2e_val, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": e2e_s / Ke * 1e3,
                    "api": "pp_acq_session_begin_host / finish_host (pinned host buffers) + host np.random.permutation draws "
                           "overlapped with the H2D copy"},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
            "clocks": clk,
        }
        if train is not None:
            out["train"] = train
            out["query_model"] = qmodel
            out["strategy_sweep_1024x2048"] = bench_strategy_sweep(dev, hbm_peak)
        if not args.no_cpu_baseline and world == 1:  # the contract: CPU baseline on rank 0 at N=1 only
            threads = os.cpu_count() or 1
            n_cpu = 32
            rate, times = cpu_query_rate(n_cpu, 3, threads)
            out["cpu_baseline"] = {"value": rate, "unit": "Mpixels/s", "cores": threads, "kind": "port",
                                   "sample": f"{n_cpu} images of the same workload x 3 repetitions (best), "
                                             f"oracle port of query.py:159-212 on torch CPU + NumPy, {sum(times):.1f} s"}
            if train is not None:
                qm, qdt = cpu_query_model_rate("mobilenet", 8, threads)
                out["query_model"]["cpu_baseline"] = {"value": qm, "unit": "Mpixels/s", "cores": threads, "kind": "port",
                                                      "sample": f"MobileNetV2-DeepLab forward + query, 8 images one per forward "
                                                                f"(oracle port, fp32 torch CPU), {qdt:.1f} s"}
                r_mn, t_mn = cpu_train_rate("mobilenet", 4, 2, threads)
                out["train"]["cpu_baseline"] = {"value": r_mn, "unit": "images/s", "cores": threads, "kind": "port",
                                                "sample": f"MobileNetV2-DeepLab B=4, 2 timed steps of the oracle train "
                                                          f"step (fp32 torch CPU, Adam), {sum(t_mn):.1f} s"}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                out["roofline"]["traffic"] = tj.get("acq_score_bytes_per_px") * B * HW
                if train is not None and tj.get("conv_d1_b32_bytes"):
                    out["train"]["roofline_tensor"]["traffic"] = tj.get("conv_d1_b32_bytes")
            except Exception:
                pass
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

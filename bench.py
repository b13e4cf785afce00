#!/usr/bin/env python
"""bench.py - headline benchmark of pixelpick_b200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

BASELINE.json metric: "train images/sec + query Mpixels/sec, Cityscapes 256x512 RN50 @1/2/4/8 B200" (configs[2]).
ONE JSON line:

  metric / value = train images/s of the RN50-DeepLabv3+ step (model.py:103-142 through networks/deeplab.py, dilated
             ResNet-50 encoder, ASPP OS8, SegmentHead) on synthetic Cityscapes-shape batches: forward_lowres -> fused x4
             upsample + sparse CE -> backward -> bucketed gradient all-reduce (NCCL, launched from backward hooks) ->
             Adam, replayed from ONE captured CUDA graph per step; batch resident in HBM; CUDA events, max over ranks;
             whole-job aggregate over the N ranks (weak scaling: the per-GPU batch is fixed).
  e2e      = the same metric with the batch coming from PINNED HOST memory every step (H2D on a copy stream under the
             previous step) and the loss read back every step - the way pixelpick_b200.Model._train_epoch drives it.
  config   = the workload plus the secondary numbers of the metric as scalars: the reference batch (4 / GPU), the
             MobileNetV2 network (BASELINE configs[1]), query Mpixels/s with the model in the loop through
             QuerySelector.__call__ (host images in, dict of picks out: query.py:159-221), the acquisition-kernel step.
  roofline = the fused acquisition scoring kernel (north star: >= 80 % of the HBM roofline), timed with CUDA events
             inside its own timed steps: algorithmic bytes (C*4 + 2) B/px over MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference = the oracle port of the reference's CPU train step (fp32 torch CPU, every host
             thread) on a bounded sample of the same workload.
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
WORKLOAD = (f"cityscapes 256x512 C={C} ResNet50-DeepLabv3+ train step (10 labelled px/image, Adam lr 5e-4 / encoder lr/10, wd 2e-4, "
            f"dropout on) + {STRATEGY} query top-5% (k={K_TOP}) n={N_SEL}")
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



# ----------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's CPU path, same metric / config as our arm
# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the train step (model.py:103-122 through
    networks/deeplab.py with the dilated ResNet-50 encoder), restated by oracle/deeplab_oracle.py on torch CPU with every
    host thread.  The reference is pure Python and cannot travel to the GPU box, so kind = "port" (the port is pinned to
    the imported reference by tests/golden/make_golden_model.py)."""
    if rank != 0:
        return
    from oracle import deeplab_oracle as dorc
    from pixelpick_b200.deeplab import DeepLab  # parameter names / shapes only
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    W_, K_ = max(args.warmup, 1), args.steps
    shapes = {k: tuple(v.shape) for k, v in DeepLab(MARGS, backbone="resnet").state_dict().items()}
    sd = dorc.synthetic_state_dict(shapes, seed=1)
    params = []
    for k in sorted(sd):
        if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            sd[k] = sd[k].clone().requires_grad_(True)
            params.append(sd[k])
    opt = torch.optim.Adam(params, lr=OPT["lr"], weight_decay=OPT["weight_decay"])

    def step(x, y, q):
        out = dorc.deeplab_forward(sd, x, backbone="resnet", training=True, drop=(0.5, 0.5, 0.2))
        loss = dorc.sparse_ce_loss(out["pred"], y, q, C)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)

    # bounded sample: the reference batch (4) when the whole run fits ~3 minutes on this host, else 2 images / step
    B = 4
    x, y, q = synth_train_batch(B, 7)
    t0 = time.perf_counter()
    step(x, y, q)
    t_first = time.perf_counter() - t0
    if t_first * (K_ + W_) > 150.0:
        B = 2
        x, y, q = synth_train_batch(B, 7)
    for _ in range(W_ - 1 if B == 4 else W_):
        step(x, y, q)
    t0 = time.perf_counter()
    for _ in range(K_):
        step(x, y, q)
    dt = time.perf_counter() - t0
    val = K_ * B / dt
    sample = (f"RN50-DeepLabv3+ train step on {B} images of 256x512 per step (fp32, torch CPU, Adam, dropout on), "
              f"{threads} threads, {K_} timed steps, {dt:.1f} s")
    cfg = {"workload": WORKLOAD, "batch_per_step": B, "parallelism": "CPU, one process"}
    if not args.no_query:
        qm, qdt = cpu_query_model_rate("resnet", 2, threads)
        cfg["query_rn50_mpix_s"] = _r(qm)
        cfg["query_sample"] = f"oracle port of query.py:159-212 incl. the RN50-DeepLab forward, 2 images one per forward, {qdt:.1f} s"
    out = {
        "impl": "reference", "metric": "train_images_per_sec", "value": _r(val), "unit": "images/s", "n_gpus": world,
        "steps": K_, "warmup": W_, "ms_per_step": _r(dt / K_ * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": _r(val), "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": _r(val), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out, separators=(",", ":")))


def _r(v, sig=5):
    """round to `sig` significant digits (keeps the JSON line short enough for the driver's tail)."""
    if v is None or isinstance(v, (int, str, bool)):
        return v
    v = float(v)
    if v == 0 or not np.isfinite(v):
        return v
    return float(f"{v:.{sig}g}")


def _round_tree(o):
    if isinstance(o, dict):
        return {k: _round_tree(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_round_tree(v) for v in o]
    if isinstance(o, float):
        return _r(o)
    return o


def bench_augment(dev, B=32):
    """datasets/base_dataset.py:174-183 on the device for one raw uint8 batch of B Cityscapes crops (256x512): host draws +
    resampling tables + pp_augment_geometric_u8 + pp_augment_photometric, wall time per batch in ms."""
    import random
    from pixelpick_b200.augment import GpuAugment
    random.seed(0)
    torch.manual_seed(0)
    np.random.seed(0)
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    y = torch.randint(0, C + 1, (B, H, W), dtype=torch.uint8, device=dev)
    q = (torch.rand((B, H, W), device=dev) < 0.01).to(torch.uint8) * 255
    aug = GpuAugment((H, W), [0.28689554, 0.32513303, 0.28389177], [0.18696375, 0.19017339, 0.18720214], C)
    for _ in range(2):
        aug(x, y, q, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        aug(x, y, q, q)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 5 * 1e3


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

# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def ppdist_rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(vals, dev, world):
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def build_train(backbone, B, dev, world, loop=False):
    """model + capturable Adam (args.py:101-106 cs groups) + bucketed gradient all-reduce + the captured step."""
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
    reducer = ppdist.GradAllReducer(model) if (world > 1 or os.environ.get("PP_FORCE_REDUCER")) else None
    hx, hy, hq = synth_train_batch(B, 11 + ppdist.rank(), pin=True)
    gs = GraphedTrainStep(model, opt, (B, H, W), C, capacity=B * 16, device=dev, reducer=reducer,
                          n_classes=C if loop else None)
    gs.load(hx, hy, hq)
    l0 = _lib.lib().pp_launch_count()
    gs.capture()
    # our kernels are counted when they are LAUNCHED (captured); a replay re-issues the same list: 3 warm-up steps + 1 capture
    per_step = (_lib.lib().pp_launch_count() - l0) / 4.0
    return model, gs, (hx, hy, hq), per_step, reducer


def bench_train_graph(backbone, B, steps, warmup, dev, world, mode="hbm"):
    """images/s of the whole train step (forward_lowres -> fused upsample + sparse CE -> backward -> bucketed gradient
    all-reduce -> Adam) replayed from ONE captured CUDA graph - what pixelpick_b200.Model._train_epoch runs.
      mode "hbm"  : batch resident in HBM, CUDA events around the K replays (device time, max over ranks)
      mode "e2e"  : every step uploads its batch from PINNED HOST memory (prefetched on a copy stream under the previous
                    step's graph) and reads the step's loss back to the host; wall clock, max over ranks
      mode "loop" : as Model._train_epoch: host batch every step, metrics accumulated on the device, ONE read per epoch"""
    model, gs, (hx, hy, hq), per_step, reducer = build_train(backbone, B, dev, world, loop=(mode == "loop"))
    loss_host = torch.zeros(1).pin_memory()
    host = mode in ("e2e", "loop")
    if host:
        gs.prefetch(hx, hy, hq)

    def step():
        if host:
            gs.commit()  # staging -> static inputs (the H2D of this batch ran during the previous step's graph)
        loss, _ = gs()
        if host:
            gs.prefetch(hx, hy, hq)  # next host batch -> staging on the copy stream, overlapping the graph
        if mode == "e2e":
            loss_host.copy_(loss.detach().reshape(1))  # D2H read of this step's loss every step
        return loss

    for _ in range(warmup):
        step()
    _barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        last = step()
    if mode == "loop":
        gs.metrics.read()  # the epoch's confusion matrix + loss sum: the loop's only device->host read
    b.record()
    _barrier(world)
    wall = time.perf_counter() - t0
    ms = a.elapsed_time(b) if not host else wall * 1e3
    (ms,) = _max_over_ranks([ms], dev, world)
    final_loss = float(last.item())
    del gs, model, reducer
    torch.cuda.empty_cache()
    return {"value": world * B * steps / (ms / 1e3), "ms_per_step": ms / steps, "batch_per_gpu": B, "steps": steps,
            "launches_per_step": per_step, "final_loss": final_loss,
            "h2d_bytes_per_step": B * (3 * H * W * 4) + 3 * B * 16 * 4 + 4 if host else 0,
            "d2h_bytes_per_step": 4 if mode == "e2e" else (C * C * 8 + 16) / steps if mode == "loop" else 0}


class _HostImages(torch.utils.data.Dataset):
    """Reference dataset interface (datasets/base_dataset.py:18-46) over pre-generated HOST tensors: items
    {'x','y','p_img'}, `.queries`, `.label_queries` - what QuerySelector.__call__ reads (query.py:144-221)."""

    def __init__(self, n, seed):
        self.x, self.y, q = synth_train_batch(n, seed)
        self.queries = [m.numpy().astype(bool) for m in q]
        self.labelled = None

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {"x": self.x[i], "y": self.y[i], "p_img": f"synthetic/{i:06d}.png"}

    def label_queries(self, dict_queries, nth_query=None):
        self.labelled = dict_queries


def bench_query_selector(backbone, n_img_total, dev, world, tmpdir, reps=2):
    """Query Mpixels/s with the model in the loop, through the public API: QuerySelector.__call__(nth_query, model) over a
    DataLoader of HOST images (batch_size 1, as model.py:36-37) -> dict of picks (query.py:159-221): forward, fused
    upsample + margin score, per-image top-k select, n random ranks, QueryStats, the round's all-gather at N > 1.
    Wall clock, max over ranks; best of `reps` rounds after one warm-up round."""
    from pixelpick_b200.deeplab import DeepLab
    from pixelpick_b200.query import QuerySelector
    torch.manual_seed(0)
    model = DeepLab(MARGS, backbone=backbone).to(dev).eval()
    ds = _HostImages(n_img_total, 21)
    dl = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0)
    qargs = Namespace(dataset_name="cs", debug=False, dir_root=tmpdir, experim_name=f"bench_{backbone}", ignore_index=C,
                      mc_n_steps=20, n_classes=C, n_pixels_by_us=N_SEL, network_name="deeplab", query_strategy=STRATEGY,
                      reverse_order=False, stride_total=16, top_n_percent=TOP_N_PERCENT, use_mc_dropout=False, vote_type="soft")
    qs = QuerySelector(qargs, dl, device=dev)
    times = []
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        sys.stdout.flush()
        os.dup2(devnull, 1)  # QuerySelector prints the reference's progress lines: keep stdout = ONE JSON line
        for r in range(reps + 1):
            np.random.seed(0)
            _barrier(world)
            t0 = time.perf_counter()
            if os.environ.get("PP_QUERY_PROFILE") and r == reps and ppdist_rank() == 0:
                import cProfile, pstats
                prof = cProfile.Profile()
                picks = prof.runcall(qs, r, model)
                pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(35)
            else:
                picks = qs(r, model)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    assert len(picks) == n_img_total and all(len(v["x_coords"]) == N_SEL for v in picks.values())
    (dt,) = _max_over_ranks([min(times[1:])], dev, world)
    del model, qs
    torch.cuda.empty_cache()
    return {"value": n_img_total * H * W / 1e6 / dt, "images_per_s": n_img_total / dt, "images": n_img_total, "s_per_round": dt,
            "h2d_bytes_per_image": 3 * H * W * 4 + 2 * H * W, "d2h_bytes_per_image": N_SEL * 8 + N_SEL * 4}


def bench_acq_pipeline(B, K, Wm, dev, world, rank, overlap, sorted_topk=False, fused=False):
    """The acquisition kernels alone on logits resident in HBM (B images / step / GPU, 2.55 GB of fp32 logits per step at
    B = 256 >> L2): fused softmax + margin + mask fills + level-0 histogram -> radix select -> order statistics at the n
    drawn ranks; at N > 1 one all-gather of the round's picks.  Returns step / score-kernel times (CUDA events)."""
    import torch.distributed as dist
    from pixelpick_b200 import _lib
    lib = _lib.lib()
    HW = H * W
    largest = _lib.LARGEST[STRATEGY]
    logits, lab, void = synth(B, 100 + rank, device=dev)
    np.random.seed(rank)
    pos = torch.from_numpy(np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(B)]).astype(np.int32)).to(dev)
    sel_round = torch.empty((K, B, N_SEL), dtype=torch.int32, device=dev) if world > 1 else None
    gathered = torch.empty((world, K, B, N_SEL), dtype=torch.int32, device=dev) if world > 1 else None
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    n_slots = 2 if overlap else 1
    slots = [(_lib.TopKWorkspace(B, HW, K_TOP, dev), torch.empty((B, H, W), dtype=torch.float32, device=dev)) for _ in range(n_slots)]
    side = torch.cuda.Stream(device=dev, priority=-1) if overlap else None
    scored = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    turn = [0]

    fused = fused and not overlap and not sorted_topk and _lib.acq_score_select_supported(logits, C, H, W)

    def step_fused(i=None):
        # ONE pass over the logits: a cluster per image scores, keeps the scores in shared memory, merges the level-0 histograms
        # and classifies (pp_acq_score_select), then the radix tail + the order statistics at the drawn ranks
        ws_p, _ = slots[0]
        ws_p.prepare()
        if i is not None:
            ev_a[i].record()
        sel = _lib.acq_score_select_pick(logits, STRATEGY, K_TOP, pos, lab, void, ws=ws_p, mark=(ev_b[i] if i is not None else None))
        if world > 1 and i is not None:
            sel_round[i].copy_(sel)
        return sel

    def step(i=None):
        if fused:
            return step_fused(i)
        p = (turn[0] % n_slots)
        turn[0] += 1
        ws_p, score_p = slots[p]
        main = torch.cuda.current_stream()
        if overlap:
            main.wait_event(free[p])  # the select + pick that used this slot two steps ago
        ws_p.prepare()
        if i is not None:
            ev_a[i].record()
        _lib.acq_score(logits, STRATEGY, lab, void, out=score_p, hist0_ws=ws_p)
        if i is not None:
            ev_b[i].record()

        def tail():
            if sorted_topk:
                sel = _lib.acq_gather(_lib.acq_topk(score_p.view(B, HW), K_TOP, largest, ws=ws_p, hist0_valid=True), pos)
            else:  # what QuerySelector runs: only the n drawn ranks are needed (query.py:63-64) -> radix pick, no sort
                sel = _lib.acq_select_pick(score_p.view(B, HW), K_TOP, largest, pos, ws=ws_p, hist0_valid=True)
            if world > 1 and i is not None:
                sel_round[i].copy_(sel)
            return sel

        if not overlap:
            return tail()
        scored[p].record(main)
        with torch.cuda.stream(side):
            side.wait_event(scored[p])
            sel = tail()
            free[p].record(side)
        return sel

    if world > 1:  # pre-warm the communicator for this shape: the first call of a collective builds its plan
        dist.all_gather_into_tensor(gathered, sel_round)
    for _ in range(Wm):
        step()
    if overlap:
        torch.cuda.current_stream().wait_stream(side)
    _barrier(world)
    l0 = lib.pp_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        step(i)
    if overlap:  # the last steps' select + pick belong to the timed region
        torch.cuda.current_stream().wait_stream(side)
    if world > 1:  # the round's exchange: per-rank picks -> every rank, ONE collective
        dist.all_gather_into_tensor(gathered, sel_round)
    t_end.record()
    _barrier(world)
    launches = lib.pp_launch_count() - l0
    ms_total = t_start.elapsed_time(t_end)
    score_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]))
    ms_total, score_ms = _max_over_ranks([ms_total, score_ms], dev, world)
    del logits, slots
    torch.cuda.empty_cache()
    return {"ms_per_step": ms_total / K, "score_ms": score_ms, "launches": int(launches), "fused": bool(fused),
            "mpix_s": world * B * HW * K / 1e6 / (ms_total / 1e3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-batch", type=int, default=32, help="throughput batch of the headline train step (per GPU)")
    ap.add_argument("--acq-batch", type=int, default=256, help="images per step per GPU of the acquisition-kernel leg")
    ap.add_argument("--query-images", type=int, default=96, help="images per GPU of the QuerySelector.__call__ leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-query", action="store_true", help="skip the model-in-the-loop query legs")
    ap.add_argument("--no-extras", action="store_true", help="headline train leg + acquisition roofline only")
    ap.add_argument("--overlap-select", action="store_true",
                    help="acquisition leg: run the select + pick of step i on a side stream under the scoring of step i+1 "
                         "(measured SLOWER on B200: 0.61 vs 0.55 ms/step - the concurrent kernels take HBM bandwidth from the "
                         "scoring kernel; kept for the record, default is one stream)")
    ap.add_argument("--fused-acq", action="store_true",
                    help="acquisition leg through pp_acq_score_select (scoring + level-0 select in one pass, a cluster per image, the "
                         "score map never written).  Bit-identical picks; measured SLOWER on B200 (0.583 vs 0.536 ms / 256 images: the "
                         "fused kernel takes 512 us against 407 + 5 + 53 us - 18 %% of every CTA's life is cluster barriers, the "
                         "leader's bucket pick and the classification, during which it does not stream), so the default stays the "
                         "three-kernel form")
    ap.add_argument("--sweep", action="store_true", help="also run the 1024x2048 strategy sweep (BASELINE configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import tempfile
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
    _lib.lib()
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1: the host side of the query round (DataLoader collate, 1.5 MB staging copies) would
        # run on one thread per rank and take twice as long as in the single-process run; give every rank its share of the cores
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
    hbm_peak, tf_burst, tf_sust, peak_src = peaks()
    K, Wm, Bt = args.steps, args.warmup, args.train_batch
    tmpdir = tempfile.mkdtemp(prefix="pp_bench_")

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- headline: RN50-DeepLabv3+ train step, throughput batch, graph replay; then the same through host batches ----
    head = bench_train_graph("resnet", Bt, K, Wm, dev, world, mode="hbm")
    e2e = bench_train_graph("resnet", Bt, K, Wm, dev, world, mode="e2e")
    cfg = {"workload": WORKLOAD, "batch_per_gpu": Bt, "global_batch": Bt * world, "parallelism": f"dp{world}",
           "step": "forward_lowres (bf16 NHWC, tcgen05 head) -> fused x4 upsample + sparse CE -> backward -> bucketed grad "
                   "all-reduce from backward hooks -> Adam; ONE captured CUDA graph per step, dropout on, 10 labelled px/image",
           "l2": "per-step working set (activations + 160 MB of fp32 gradients) far larger than the 126 MB L2; no flush needed",
           "train_rn50_ms_per_step": head["ms_per_step"], "train_rn50_e2e_ms_per_step": e2e["ms_per_step"]}
    extras = {}
    if not args.no_extras:
        ts = max(5, min(K, 20))
        r4 = bench_train_graph("resnet", 4, ts, 5, dev, world, mode="hbm")
        r4e = bench_train_graph("resnet", 4, ts, 5, dev, world, mode="e2e")
        m32 = bench_train_graph("mobilenet", Bt, ts, 5, dev, world, mode="hbm")
        m4 = bench_train_graph("mobilenet", 4, ts, 5, dev, world, mode="hbm")
        m4l = bench_train_graph("mobilenet", 4, ts, 5, dev, world, mode="loop")
        cfg.update({"train_rn50_b4_img_s": r4["value"], "train_rn50_b4_e2e_img_s": r4e["value"],
                    f"train_mnv2_b{Bt}_img_s": m32["value"], "train_mnv2_b4_img_s": m4["value"],
                    "train_mnv2_b4_loop_img_s": m4l["value"]})
        extras["train"] = {"resnet50_b4": r4, "resnet50_b4_e2e": r4e, f"mobilenetv2_b{Bt}": m32, "mobilenetv2_b4": m4,
                           "mobilenetv2_b4_loop": m4l}
    # ---- query with the model in the loop, through QuerySelector.__call__ (host images in, dict of picks out) ----
    if not args.no_query:
        q_rn = bench_query_selector("resnet", args.query_images * world, dev, world, tmpdir)
        cfg.update({"query_rn50_mpix_s": q_rn["value"], "query_rn50_img_s": q_rn["images_per_s"], "query_images": q_rn["images"]})
        extras["query_model"] = {"resnet50": q_rn}
        if not args.no_extras:
            q_mn = bench_query_selector("mobilenet", args.query_images * world, dev, world, tmpdir)
            cfg["query_mnv2_mpix_s"] = q_mn["value"]
            extras["query_model"]["mobilenetv2"] = q_mn
            if rank == 0 and world == 1:
                b1 = bench_query_model("resnet", 1, 50, dev, True)
                cfg["query_rn50_bs1_mpix_s"] = b1["value"]  # the reference's own query batch (one image per forward)
    # ---- acquisition kernels alone (north star: >= 80 % of the HBM roofline on the fused scoring kernel) ----
    acq = bench_acq_pipeline(args.acq_batch, K, Wm, dev, world, rank, overlap=args.overlap_select, fused=args.fused_acq)
    HW = H * W
    alg = args.acq_batch * HW * ALG_BYTES_PER_PX
    achieved = alg / (acq["score_ms"] / 1e3) / 1e9
    cfg.update({"acq_images_per_step_per_gpu": args.acq_batch, "acq_step_ms": acq["ms_per_step"], "acq_step_gpix_s": acq["mpix_s"] / 1e3,
                "acq_step_frac_of_hbm": alg / (acq["ms_per_step"] / 1e3) / 1e9 / hbm_peak,
                "acq_select": ("scoring fused with the level-0 select (scores stay in shared memory), radix tail, order statistics at "
                               "the drawn ranks; " if acq["fused"] else "radix select (persistent level-0 kernel) + order statistics at the drawn ranks; ") +
                              ("select+pick of step i on a side stream under the scoring of step i+1" if args.overlap_select else "one stream")})
    roof = {"kernel": ("acq_score_select_kernel<19, margin> (softmax + margin + mask fills + level-0 select in one pass, cluster of 8 "
                       "CTAs per image)" if acq["fused"] else "acq_score_pf_kernel<19, margin, f32, fused hist0> (next tile in flight while scoring)"), "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": acq["score_ms"],
            "alg_bytes_per_launch": alg, "share_of_acq_step": acq["score_ms"] / acq["ms_per_step"],
            "peak_note": "peak = the driver's COPY bandwidth (half reads, half writes); this kernel's traffic is 95 % reads, so frac may pass 1"}
    if rank == 0 and not args.no_extras:
        cfg["augment_b32_ms"] = bench_augment(dev)  # the device input pipeline (geometric + photometric) for one batch of 32
        conv = bench_conv_roofline(dev, tf_burst)
        cfg.update({"conv_seghead_tflops": conv["achieved"], "conv_seghead_frac_of_bf16_peak": conv["frac"]})
        extras["roofline_tensor"] = conv
        if args.sweep and world == 1:
            extras["strategy_sweep_1024x2048"] = bench_strategy_sweep(dev, hbm_peak)
    _barrier(world)
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        out = {
            "metric": "train_images_per_sec", "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "gpu_launches": int(round(head["launches_per_step"] * K)),
            "config": cfg, "roofline": roof,
            "e2e": {"value": e2e["value"], "unit": "images/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "ms_per_step": e2e["ms_per_step"],
                    "api": "GraphedTrainStep.prefetch/commit/replay as Model._train_epoch drives it: pinned host batch -> H2D on a "
                           "copy stream under the previous step, loss read back every step"},
            "clocks": clk,
        }
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                out["roofline"]["traffic"] = json.load(open(tpath)).get("acq_score_bytes_per_px") * args.acq_batch * HW
                out["roofline"]["traffic_source"] = "ncu --set full capture of the same launch (profiles/), not re-measured in this run"
            except Exception:
                pass
        if not args.no_cpu_baseline and world == 1:  # the contract: CPU baseline on rank 0 at N=1 only
            threads = os.cpu_count() or 1
            r_rn, t_rn = cpu_train_rate("resnet", 4, 2, threads)
            out["cpu_baseline"] = {"value": r_rn, "unit": "images/s", "cores": threads, "kind": "port",
                                   "sample": f"RN50-DeepLabv3+ B=4, 2 timed steps of the oracle train step (fp32 torch CPU, Adam, "
                                             f"dropout on), {sum(t_rn):.1f} s"}
            if not args.no_query:
                qm, qdt = cpu_query_model_rate("resnet", 2, threads)
                out["config"]["cpu_query_rn50_mpix_s"] = qm
        out.update(extras)
        print(json.dumps(_round_tree(out), separators=(",", ":")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

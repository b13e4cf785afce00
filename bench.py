#!/usr/bin/env python
"""bench.py — headline benchmark of the pixelpick_b200 hot paths (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], Cityscapes 256x512, C=19, margin_sampling, top-5 % -> k=6553,
n_pixels_by_us=10; synthetic logits/masks, seeded):
  step      = one pass of the query/acquisition hot path over one batch of B images per GPU:
              fused softmax+margin+mask score -> per-image sorted top-k -> gather of the n picks
  value     = Mpixels/s over all ranks, inputs resident in HBM (CUDA events, max over ranks)
  e2e       = the same metric through the C-ABI host-buffer call (pp_acq_session_run_host): pinned host
              logits/masks -> H2D -> kernels -> D2H of the selected indices, every step, including the
              host-side NumPy draw of the pick positions (np.random.choice semantics)
  roofline  = the scoring kernel alone, timed with CUDA events INSIDE the timed steps, algorithmic
              bytes = (C*4 + 2) B/px over MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference = the oracle port of the reference's CPU path (torch CPU kernels +
              NumPy RNG, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, H, W = 19, 256, 512
STRATEGY = "margin_sampling"
TOP_N_PERCENT, N_SEL = 0.05, 10
K_TOP = int(H * W * TOP_N_PERCENT)
ALG_BYTES_PER_PX = C * 4 + 2  # logits + labelled mask + void mask
WORKLOAD = f"cityscapes 256x512 C={C} {STRATEGY} top-5% (k={K_TOP}) n={N_SEL}: logits -> score -> top-k -> picks"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def cpu_reference_rate(n_img, reps, threads):
    """Oracle port of the reference CPU path (query.py:159-212 loop body over pre-computed logits)."""
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
    print(json.dumps({
        "impl": "reference", "metric": "query_mpixels_per_sec", "value": val, "unit": "Mpixels/s", "n_gpus": world,
        "steps": K_, "warmup": W_, "ms_per_step": dt / K_ * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": n_img},
        "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per step per GPU (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=64, help="images per step per GPU (host-buffer leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    B, K, Wm = args.batch, args.steps, args.warmup
    HW = H * W
    largest = _lib.LARGEST[STRATEGY]
    logits, lab, void = synth(B, 100 + rank, device=dev)
    ws = _lib.TopKWorkspace(B, HW, K_TOP, dev)
    score = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    np.random.seed(rank)
    pos = torch.from_numpy(np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(B)]).astype(np.int32)).to(dev)
    gathered = [torch.empty((B, N_SEL), dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None

    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]

    def step(i=None):
        ws.prepare()
        if i is not None:
            ev_a[i].record()
        _lib.acq_score(logits, STRATEGY, lab, void, out=score, hist0_ws=ws)
        if i is not None:
            ev_b[i].record()
        topk = _lib.acq_topk(score.view(B, HW), K_TOP, largest, ws=ws, hist0_valid=True)
        sel = _lib.acq_gather(topk, pos)
        if world > 1:  # the path's one exchange: per-rank picks -> every rank (rank 0 builds the dict)
            dist.all_gather(gathered, sel)
        return sel

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
    t_end.record()
    barrier()
    launches = lib.pp_launch_count() - l0
    ms_total = t_start.elapsed_time(t_end)
    score_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]))

    # ---- e2e: host buffers through the C-ABI session ------------------------------------------
    Be = args.e2e_batch
    h_logits, h_lab, h_void = synth(Be, 500 + rank, pin=True)
    sess = _lib.AcqSession(min(16, Be), C, H, W, K_TOP, N_SEL)
    h_pos = torch.empty((Be, N_SEL), dtype=torch.int32).pin_memory()
    h_sel = torch.empty((Be, N_SEL), dtype=torch.int32).pin_memory()

    def e2e_step():
        h_pos.numpy()[:] = np.stack([np.random.permutation(K_TOP)[:N_SEL] for _ in range(Be)])
        sess.run(h_logits, h_lab, h_void, STRATEGY, h_pos, h_sel)
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
    clk = clocks.stop() if rank == 0 else None
    sess.close()

    t = torch.tensor([ms_total, score_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, score_ms, e2e_s = [float(v) for v in t.cpu()]

    if rank == 0:
        peak, peak_src = peaks()
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
                       "l2": f"inputs larger than L2 ({B * C * HW * 4 / 1e6:.0f} MB of logits per step)",
                       "parallelism": f"images sharded over {world} rank(s); all_gather of picks" if world > 1 else "single GPU"},
            "roofline": {"kernel": "acq_score_vec_kernel<19, margin, f32, fused hist0>", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel_ms": score_ms,
                         "alg_bytes_per_launch": B * HW * ALG_BYTES_PER_PX,
                         "share_of_step": score_ms / (ms_total / K)},
            "e2e": {"value": e2e_val, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": e2e_s / Ke * 1e3,
                    "api": "pp_acq_session_run_host (pinned host buffers) + host np.random.permutation draws"},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
            "clocks": clk,
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = 32
            rate, times = cpu_reference_rate(n_cpu, 3, threads)
            out["cpu_baseline"] = {"value": rate, "unit": "Mpixels/s", "cores": threads, "kind": "port",
                                   "sample": f"{n_cpu} images of the same workload x 3 repetitions (best), "
                                             f"oracle port of query.py:159-212 on torch CPU + NumPy, {sum(times):.1f} s"}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                out["roofline"]["traffic"] = json.load(open(tpath)).get("acq_score_bytes_per_px") * B * HW
            except Exception:
                pass
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

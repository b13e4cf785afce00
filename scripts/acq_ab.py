"""GPU: same-box A/B of the acquisition pipeline between library builds, per phase (CUDA events) + a checksum of the picks.

    python scripts/acq_ab.py [other_build.so ...]      # e.g. pixelpick_b200/csrc/build/libpp_prev.so (an earlier build)

Every build is loaded into this one process (separate CDLL instances) and run on the same device tensors, so equal
`picks` checksums mean bit-identical selections; the phase times are score / select (pick_bucket0 + select_l0 +
select_rest) / pick, means over `steps` steps.  The tree's build is also run with PP_SELECT_L0=1 (one-chunk-per-CTA
level-0 kernel; PP_AB_L0=1,2,3 for more of acq.cu's variants) from a copy of the file, because the switch is read once per
loaded library."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import shutil  # noqa: E402

from pixelpick_b200 import _lib  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
C = bench.C
TREE_LIB = _lib.LIB_PATH


SELF_CLEANING = True  # the loaded build's select leaves the workspace prepared (False for builds older than this feature)


def use(path, select_l0=None, score=None):
    """make `path` the library behind _lib (a fresh CDLL instance with its own statics)"""
    global SELF_CLEANING
    SELF_CLEANING = "prev" not in os.path.basename(path)
    os.environ.pop("PP_SELECT_L0", None)
    os.environ.pop("PP_SCORE_VARIANT", None)
    if select_l0 is not None:
        os.environ["PP_SELECT_L0"] = select_l0
    if score is not None:
        os.environ["PP_SCORE_VARIANT"] = score
    _lib._lib = None
    _lib.LIB_PATH = os.path.abspath(path)
    return _lib.lib()


def run(n_img, H, W, strat, steps=10, seed=7, scale=3.0):
    lib = _lib.lib()
    HW = H * W
    k = int(HW * bench.TOP_N_PERCENT)
    g = torch.Generator(device=dev).manual_seed(seed)
    logits = torch.randn((n_img, C, H, W), generator=g, device=dev) * scale
    rs = np.random.RandomState(seed)
    lab = torch.from_numpy((rs.rand(n_img, H, W) < 100.0 / HW).astype(np.uint8)).to(dev)
    void = torch.from_numpy((rs.rand(n_img, H, W) < 0.01).astype(np.uint8)).to(dev)
    pos = torch.from_numpy(np.stack([rs.permutation(k)[:bench.N_SEL] for _ in range(n_img)]).astype(np.int32)).to(dev)
    ws = _lib.TopKWorkspace(n_img, HW, k, dev)
    score = torch.empty((n_img, H, W), dtype=torch.float32, device=dev)
    out = torch.empty((n_img, bench.N_SEL), dtype=torch.int32, device=dev)
    largest = int(bool(_lib.LARGEST[strat]))
    st = _lib._stream(score)

    def step(ev=None):
        ws.prepare()  # skipped by the wrapper when the previous select handed the workspace back zeroed
        if ev:
            ev[0].record()
        _lib.acq_score(logits, strat, lab, void, out=score, hist0_ws=ws)
        if ev:
            ev[1].record()
        _lib.check(lib.pp_acq_select(_lib._ptr(score), n_img, HW, k, largest, 1, _lib._ptr(ws.buf), ws.nbytes, st), "select")
        ws._clean = n_img if SELF_CLEANING else None  # what _lib.acq_select_pick does for the tree's build
        if ev:
            ev[2].record()
        _lib.check(lib.pp_acq_pick(_lib._ptr(ws.buf), ws.nbytes, n_img, HW, k, _lib._ptr(pos), bench.N_SEL, _lib._ptr(out), st), "pick")
        if ev:
            ev[3].record()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for ev in evs:
        step(ev)
    b.record()
    torch.cuda.synchronize()
    ph = [float(np.mean([e[i].elapsed_time(e[i + 1]) for e in evs])) * 1e3 for i in range(3)]
    picks = out.cpu().numpy()
    return {"shape": f"{n_img}x{H}x{W}", "strategy": strat, "k": k, "step_us": round(a.elapsed_time(b) / steps * 1e3, 1),
            "score_us": round(ph[0], 1), "select_us": round(ph[1], 1), "pick_us": round(ph[2], 1),
            "picks": hashlib.sha1(np.ascontiguousarray(picks).tobytes()).hexdigest()[:12]}


if __name__ == "__main__":
    cfgs = [(256, 256, 512, "margin_sampling"), (256, 256, 512, "entropy"), (32, 360, 480, "least_confidence"),
            (8, 1024, 2048, "entropy"), (8, 1024, 2048, "margin_sampling"), (8, 1024, 2048, "least_confidence"),
            (1, 256, 512, "margin_sampling")]
    builds = [(os.path.basename(p), p, None) for p in sys.argv[1:]]
    for v in os.environ.get("PP_AB_L0", "1").split(","):  # the tree's build with the other level-0 kernels
        copy = os.path.join(os.path.dirname(TREE_LIB), "build", f"libpp_tree_l0_{v}.so")
        shutil.copyfile(TREE_LIB, copy)
        builds.append((f"tree, PP_SELECT_L0={v}", copy, v))
    for v in [x for x in os.environ.get("PP_AB_SCORE", "").split(",") if x]:  # ... and with the other scoring kernels
        copy = os.path.join(os.path.dirname(TREE_LIB), "build", f"libpp_tree_sc_{v}.so")
        shutil.copyfile(TREE_LIB, copy)
        builds.append((f"tree, PP_SCORE_VARIANT={v}", copy, ("score", v)))
    builds.append(("tree", TREE_LIB, None))
    if os.environ.get("PP_AB_CFGS"):
        cfgs = cfgs[:int(os.environ["PP_AB_CFGS"])]
    for name, path, sel in builds:
        if isinstance(sel, tuple):
            use(path, None, sel[1])
        else:
            use(path, sel)
        print("==", name, flush=True)
        for c in cfgs:
            print(json.dumps(run(*c)), flush=True)

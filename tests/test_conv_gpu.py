"""GPU: tcgen05 implicit-GEMM convolution (pp_conv_igemm) vs torch fp32 conv2d on bf16-rounded operands.

Tolerance: operands are rounded to bf16 on both sides, products are exact in fp32 and only the
accumulation order differs, so f32 outputs agree to 1e-3 relative of the output scale (K up to 18432);
bf16 outputs add one bf16 rounding (2^-9 relative)."""
import pytest
import torch
import torch.nn.functional as F

from pixelpick_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _case(N, H, W, Cin, Cout, k, dil, seed=0, ld_in=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((N, Cin, H, W), generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.float(), w.float(), padding=dil if k == 3 else 0, dilation=dil)
    cin_pad = -(-Cin // 64) * 64
    ld = ld_in or cin_pad
    x_nhwc = torch.zeros((N, H, W, ld), dtype=torch.bfloat16)
    x_nhwc[..., :Cin] = x.permute(0, 2, 3, 1)
    return x_nhwc.to(DEV), w, ref


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,dil,bn", [
    (2, 64, 128, 320, 256, 3, 1, 0),     # SegmentHead conv #1 (304 padded to 320)
    (2, 64, 128, 256, 256, 3, 1, 128),   # SegmentHead conv #2
    (2, 16, 32, 320, 256, 1, 1, 0),      # ASPP 1x1 (MobileNetV2)
    (2, 16, 32, 320, 256, 3, 6, 0),      # ASPP dilated branches
    (2, 16, 32, 320, 256, 3, 12, 64),
    (2, 16, 32, 320, 256, 3, 18, 32),
    (1, 32, 64, 2048, 256, 3, 12, 0),    # ASPP ResNet-50 OS8
    (1, 32, 64, 1024, 256, 1, 1, 256),   # ASPP projection (without the pooled branch)
    (2, 23, 30, 320, 256, 3, 6, 0),      # CamVid 360x480 / 16: ragged tiles
    (1, 90, 120, 256, 256, 3, 1, 0),     # CamVid decoder
    (3, 5, 7, 64, 32, 3, 2, 0),          # tiny
])
def test_conv_matches_torch(N, H, W, Cin, Cout, k, dil, bn):
    x, w, ref = _case(N, H, W, Cin, Cout, k, dil)
    wp = _lib.pack_conv_weight(w.to(DEV))
    out = _lib.conv_igemm(x, wp, Cout, dil=dil, block_n=bn)
    got = out.float().permute(0, 3, 1, 2).cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 8e-3 * scale, ((got - ref).abs().max().item(), scale)
    out32 = _lib.conv_igemm(x, wp, Cout, dil=dil, out_mode=1, block_n=bn).cpu()
    assert (out32 - ref).abs().max().item() < 1e-3 * scale


def test_classifier_bias_f32_nchw():
    x, w, ref = _case(2, 64, 128, 256, 19, 1, 1, seed=3)
    g = torch.Generator().manual_seed(9)
    bias = torch.randn(19, generator=g)
    wp = _lib.pack_conv_weight(w.to(DEV))
    shift = torch.zeros(32)
    shift[:19] = bias
    out = _lib.conv_igemm(x, wp, 19, shift=shift.to(DEV), out_mode=1).cpu()
    assert out.shape == (2, 19, 64, 128)
    assert torch.allclose(out, ref + bias.view(1, -1, 1, 1), atol=2e-3, rtol=1e-3)


def test_epilogue_bn_relu_prebias_and_concat_slice():
    x, w, ref = _case(2, 16, 32, 320, 256, 3, 6, seed=5)
    g = torch.Generator().manual_seed(6)
    scale, shift = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    pre = torch.randn((2, 256), generator=g)
    wp = _lib.pack_conv_weight(w.to(DEV))
    buf = torch.full((2, 16, 32, 1024), 7.0, dtype=torch.bfloat16, device=DEV)
    _lib.conv_igemm(x, wp, 256, dil=6, pre_bias=pre.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), relu=True,
                    out=buf, c_off=512)
    want = torch.relu((ref + pre.view(2, 256, 1, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    got = buf[..., 512:768].float().permute(0, 3, 1, 2).cpu()
    assert (got - want).abs().max().item() < 1e-2 * want.abs().max().item()
    assert bool((buf[..., :512] == 7).all()) and bool((buf[..., 768:] == 7).all())  # neighbours untouched


def test_dgrad_is_conv_with_flipped_transposed_weights():
    N, H, W, Cin, Cout, dil = 2, 16, 32, 320, 256, 6
    g = torch.Generator().manual_seed(2)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) / 50).to(torch.bfloat16)
    gy = torch.randn((N, Cout, H, W), generator=g).to(torch.bfloat16)
    x = torch.zeros((N, Cin, H, W), requires_grad=True)
    F.conv2d(x, w.float(), padding=dil, dilation=dil).backward(gy.float())
    wp = _lib.pack_conv_weight(w.to(DEV), transpose_for_dgrad=True)
    gx = _lib.conv_igemm(gy.permute(0, 2, 3, 1).contiguous().to(DEV), wp, Cin, dil=dil, out_mode=1).cpu()
    assert (gx - x.grad).abs().max().item() < 1e-3 * x.grad.abs().max().item()


@pytest.mark.parametrize("N,H,W,Cin,dils", [(2, 16, 32, 320, (1, 6, 12, 18)), (1, 32, 64, 2048, (1, 12, 24, 36)),
                                            (1, 23, 30, 320, (1, 6, 12, 18))])
def test_fused_aspp_dgrad_multi_tap_table(N, H, W, Cin, dils):
    """pp_conv_igemm_multi: the data gradient of the four ASPP branches (aspp.py:49-52,64-68) as ONE implicit GEMM over
    1 + 9 + 9 + 9 tap entries reading channel slices of a 1024-wide gradient buffer == sum of four conv2d input grads."""
    g = torch.Generator().manual_seed(Cin + H)
    ws = [(torch.randn((256, Cin, 1 if d == 1 else 3, 1 if d == 1 else 3), generator=g) / 40).to(torch.bfloat16) for d in dils]
    gy = torch.randn((N, 1024, H, W), generator=g).to(torch.bfloat16)
    x = torch.zeros((N, Cin, H, W), requires_grad=True)
    tot = 0
    for i, (w, d) in enumerate(zip(ws, dils)):
        y = F.conv2d(x, w.float(), padding=0 if d == 1 else d, dilation=d)
        tot = tot + (y * gy[:, 256 * i:256 * (i + 1)].float()).sum()
    tot.backward()
    cin_pad = -(-Cin // 64) * 64
    n_taps = sum(w.shape[2] * w.shape[3] for w in ws)
    w_all = torch.empty((n_taps, cin_pad, 256), dtype=torch.bfloat16, device=DEV)
    entries, t0 = [], 0
    for i, (w, d) in enumerate(zip(ws, dils)):
        taps = w.shape[2] * w.shape[3]
        _lib.pack_conv_weights(w.float().to(DEV), Cin, dgrad_pad=(cin_pad, 256), dgrad_out=w_all[t0:t0 + taps])
        entries += [((t // 3 - 1) * d, (t % 3 - 1) * d, 256 * i) if taps == 9 else (0, 0, 256 * i) for t in range(taps)]
        t0 += taps
    gx = _lib.conv_igemm_multi(gy.permute(0, 2, 3, 1).contiguous().to(DEV), w_all, entries, cin_pad)
    got = gx[..., :Cin].float().permute(0, 3, 1, 2).cpu()
    assert (got - x.grad).abs().max().item() < 8e-3 * x.grad.abs().max().item()


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,dil,act,with_res", [
    (2, 18, 34, 24, 144, 1, 1, 2, False),    # MobileNetV2 expansion on the padded tensor: Cin, Cout not multiples of 64
    (2, 16, 32, 144, 24, 1, 1, 0, True),     # projection + residual, ragged 24-channel output
    (1, 130, 258, 16, 96, 1, 1, 2, False),   # flattened pixel list (M % 16 == 0)
    (1, 7, 9, 32, 16, 1, 1, 0, False),       # M % 16 != 0: real geometry
    (2, 32, 64, 256, 256, 3, 4, 1, False),   # ResNet layer4 conv2 (dilated)
    (2, 32, 64, 512, 2048, 1, 1, 1, True),   # bottleneck tail relu(bn3(conv3) + identity)
    (1, 16, 32, 960, 320, 1, 1, 0, False),
])
def test_conv_fused_epilogue_ragged_channels(N, H, W, Cin, Cout, k, dil, act, with_res):
    """conv_fused: conv + folded BatchNorm (scale/shift) + activation (+ residual) on UNPADDED activations (TMA zero-fills K)."""
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn((N, Cin, H, W), generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
    sc, sf = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.5
    res = torch.randn((N, Cout, H, W), generator=g).to(torch.bfloat16) if with_res else None
    ref = F.conv2d(x.float(), w.float(), padding=dil if k == 3 else 0, dilation=dil) * sc.view(1, -1, 1, 1) + sf.view(1, -1, 1, 1)
    if with_res:
        ref = ref + res.float()
    ref = F.relu(ref) if act == 1 else (F.relu6(ref) if act == 2 else ref)
    cp = 32 if Cout <= 32 else -(-Cout // 64) * 64
    wp = _lib.pack_conv_weight(w.to(DEV), -(-Cin // 64) * 64, cp)
    pad = lambda t: F.pad(t, (0, cp - Cout)).contiguous().to(DEV)
    out = _lib.conv_fused(x.permute(0, 2, 3, 1).contiguous().to(DEV), wp, Cout, dil=dil, scale=pad(sc), shift=pad(sf), act=act,
                          res=res.permute(0, 2, 3, 1).contiguous().to(DEV) if with_res else None)
    assert tuple(out.shape) == (N, H, W, Cout)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("co,ci_tot,cin,k", [(256, 64, 64, 1), (64, 256, 256, 1), (2048, 512, 512, 1), (24, 144, 144, 1), (256, 1280, 1024, 1),
                                            (256, 304, 304, 3), (128, 128, 128, 3), (96, 40, 40, 3), (19, 256, 256, 1), (48, 24, 24, 5)])
def test_weight_packing_kernels_equal_the_torch_layout(co, ci_tot, cin, k):
    """pp_pack_conv_weight (tiled through shared memory for 1x1 / 3x3, gather form otherwise): both operand images, bit for bit,
    with ragged channel counts, a channel sub-range (cin < Cin_total) and the padding written as zeros"""
    w = torch.randn((co, ci_tot, k, k), generator=torch.Generator().manual_seed(co + ci_tot)).to(DEV)
    taps = k * k
    fwd_pad = (-(-co // 32) * 32 if co > 32 else 32, -(-cin // 64) * 64)
    dgr_pad = (-(-cin // 64) * 64, -(-co // 64) * 64)
    fwd, dgr = _lib.pack_conv_weights(w, cin, fwd_pad=fwd_pad, dgrad_pad=dgr_pad)
    ws = w[:, :cin].contiguous()
    assert torch.equal(fwd, _lib.pack_conv_weight(ws, fwd_pad[1], fwd_pad[0]))
    assert torch.equal(dgr, _lib.pack_conv_weight(ws, dgr_pad[1], dgr_pad[0], transpose_for_dgrad=True))
    only_d = torch.full((taps,) + dgr_pad, 7.0, dtype=torch.bfloat16, device=DEV)
    _lib.pack_conv_weights(w, cin, dgrad_pad=dgr_pad, dgrad_out=only_d)
    assert torch.equal(only_d, dgr)


def test_batched_weight_packing_equals_the_per_conv_launches():
    import torch.nn as nn
    from pixelpick_b200.deeplab import _EncoderTrainPlan
    torch.manual_seed(0)
    convs = [nn.Conv2d(64, 256, 1, bias=False), nn.Conv2d(256, 64, 1, bias=False), nn.Conv2d(64, 64, 3, padding=1, bias=False),
             nn.Conv2d(24, 144, 1, bias=False), nn.Conv2d(512, 512, 3, padding=4, dilation=4, bias=False), nn.Conv2d(960, 320, 1, bias=False)]
    convs = [c.to(DEV) for c in convs]
    plan = _EncoderTrainPlan(convs, DEV)
    plan.wp.fill_(3.0)
    plan.wd.fill_(3.0)
    plan.begin_step()
    for c in convs:
        sl = plan.slots[id(c)]
        f, d = _lib.pack_conv_weights(c.weight, c.in_channels, fwd_pad=tuple(sl.wp.shape[1:]), dgrad_pad=tuple(sl.wd.shape[1:]))
        assert torch.equal(sl.wp, f) and torch.equal(sl.wd, d), c

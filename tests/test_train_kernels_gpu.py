"""GPU: wgrad (tcgen05, MN-major operands), BatchNorm/ReLU/Dropout forward+backward, NHWC upsample and layout
kernels vs torch fp32 on CPU.  Tolerances are relative to the tensor scale: 2e-3 for f32 outputs of bf16-operand
contractions, 1e-2 for bf16 outputs (one bf16 rounding = 2^-9)."""
import pytest
import torch
import torch.nn.functional as F

from pixelpick_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _nhwc(x, ld=None):
    N, C, H, W = x.shape
    ld = ld or C
    out = torch.zeros((N, H, W, ld), dtype=torch.bfloat16)
    out[..., :C] = x.permute(0, 2, 3, 1)
    return out.to(DEV)


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,dil,splits", [
    (2, 64, 128, 304, 256, 3, 1, 0),   # SegmentHead conv #1 (ld 320)
    (2, 64, 128, 256, 256, 3, 1, 3),   # SegmentHead conv #2
    (2, 16, 32, 320, 256, 3, 6, 0),    # ASPP
    (2, 16, 32, 320, 256, 3, 18, 0),   # ASPP, most taps outside
    (2, 16, 32, 320, 256, 1, 1, 1),
    (1, 32, 64, 1024, 256, 1, 1, 0),   # projection
    (2, 64, 128, 24, 48, 1, 1, 0),     # low-level conv (Cin 24 -> ld 64, Cout 48 -> 64)
    (2, 64, 128, 256, 19, 1, 1, 0),    # classifier (Cout 19 -> 64)
    (1, 23, 30, 320, 256, 3, 12, 2),   # ragged
])
def test_wgrad_matches_torch(N, H, W, Cin, Cout, k, dil, splits):
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn((N, Cin, H, W), generator=g).to(torch.bfloat16)
    gy = torch.randn((N, Cout, H, W), generator=g).to(torch.bfloat16)
    w = torch.zeros((Cout, Cin, k, k), requires_grad=True)
    F.conv2d(x.float(), w, padding=dil if k == 3 else 0, dilation=dil).backward(gy.float())
    cout_pad = 64 if Cout <= 64 else (128 if Cout <= 128 else 256)
    ld_x = -(-Cin // 64) * 64
    dw = _lib.conv_wgrad(_nhwc(x, ld_x), Cin, _nhwc(gy, cout_pad), cout_pad, k * k, dil, splits)
    got = dw[:, :Cin, :Cout].permute(2, 1, 0).reshape(Cout, Cin, k, k).cpu()
    scale = w.grad.abs().max().item()
    assert (got - w.grad).abs().max().item() < 2e-3 * scale, ((got - w.grad).abs().max().item(), scale)
    assert bool((dw[:, Cin:, :] == 0).all()) and bool((dw[:, :, Cout:].abs() < 1e-6 * scale).all() if Cout < cout_pad else True)


@pytest.mark.parametrize("M,ld,c_off,C", [(8192, 256, 0, 256), (4100, 1024, 256, 256), (32768, 64, 0, 48), (1000, 1024, 0, 1024)])
def test_bn_stats(M, ld, c_off, C):
    g = torch.Generator().manual_seed(M)
    raw = (torch.randn((M, ld), generator=g) * 2 + 0.5).to(torch.bfloat16)
    sums = _lib.bn_stats(raw.to(DEV), c_off, C).cpu()
    ref = raw[:, c_off:c_off + C].double()
    assert torch.allclose(sums[0].double(), ref.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(sums[1].double(), (ref * ref).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("relu,p", [(True, 0.0), (True, 0.5), (False, 0.2)])
def test_bn_apply_and_backward_match_autograd(relu, p):
    M, C = 4096, 256
    g = torch.Generator().manual_seed(3)
    raw = torch.randn((M, C), generator=g).to(torch.bfloat16)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    x = raw.float()
    mean, var = x.mean(0), x.var(0, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    scale, shift = gamma * rstd, beta - mean * gamma * rstd
    out = torch.empty((M, 320), dtype=torch.bfloat16, device=DEV)
    _lib.bn_apply(raw.to(DEV), 0, C, scale.to(DEV), shift.to(DEV), relu, out, 64, drop_p=p, seed=11, offset=5)
    y = out[:, 64:64 + C].float().cpu()
    keep = (y != 0) if p > 0 else torch.ones_like(y, dtype=torch.bool)
    if p > 0:
        frac = 1.0 - keep.float().mean().item() if not relu else None
        if frac is not None:
            assert abs(frac - p) < 0.02
    # autograd reference with the SAME dropout mask (recovered from the kernel output)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = F.batch_norm(xr, None, None, gr, br, training=True, eps=1e-5)
    act = torch.relu(bn) if relu else bn
    if p > 0:
        mask = keep if not relu else ((y != 0) | (act.detach() <= 0))
        act = act * mask / (1 - p)
    assert (act.detach() - y).abs().max().item() < 2e-2 * act.detach().abs().max().item()
    dy = torch.randn((M, C), generator=g).to(torch.bfloat16)
    act.backward(dy.float())
    draw, sums = _lib.bn_bwd(dy.to(DEV), 0, raw.to(DEV), 0, C, scale.to(DEV), shift.to(DEV), mean.to(DEV), rstd.to(DEV),
                             relu, drop_p=p, seed=11, offset=5)
    assert (draw.float().cpu() - xr.grad).abs().max().item() < 2e-2 * xr.grad.abs().max().item()
    assert torch.allclose(sums[0].cpu(), br.grad, rtol=2e-2, atol=2e-2 * br.grad.abs().max().item())
    assert torch.allclose(sums[1].cpu(), gr.grad, rtol=2e-2, atol=2e-2 * gr.grad.abs().max().item())


@pytest.mark.parametrize("N,h,w,H,W,C", [(2, 16, 32, 64, 128, 256), (1, 23, 30, 90, 120, 256), (2, 32, 64, 64, 128, 64),
                                         (1, 9, 7, 9, 7, 64), (1, 1, 5, 4, 17, 8), (1, 3, 3, 64, 2, 16)])
def test_upsample_nhwc_fwd_bwd(N, h, w, H, W, C):
    g = torch.Generator().manual_seed(h)
    x = torch.randn((N, C, h, w), generator=g).to(torch.bfloat16)
    out = torch.zeros((N, H, W, 320), dtype=torch.bfloat16, device=DEV)
    _lib.upsample_nhwc(_nhwc(x), out, 0, C)
    xr = x.float().requires_grad_(True)
    ref = F.interpolate(xr, size=(H, W), mode="bilinear", align_corners=True)
    got = out[..., :C].float().permute(0, 3, 1, 2).cpu()
    assert (got - ref.detach()).abs().max().item() < 1e-2 * ref.abs().max().item()
    go = torch.randn((N, C, H, W), generator=g).to(torch.bfloat16)
    ref.backward(go.float())
    gin = _lib.upsample_nhwc_bwd(_nhwc(go, 320), 0, C, (h, w)).permute(0, 3, 1, 2).cpu()
    assert (gin - xr.grad).abs().max().item() < 1e-3 * xr.grad.abs().max().item()


def test_to_nhwc_bf16_layouts():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 24, 17, 33), generator=g)
    got = _lib.to_nhwc_bf16(x.to(DEV))  # padded to 64 channels
    assert got.shape == (2, 17, 33, 64) and bool((got[..., 24:] == 0).all())
    assert torch.equal(got[..., :24].cpu(), x.to(torch.bfloat16).permute(0, 2, 3, 1))
    xc = x.to(DEV).to(memory_format=torch.channels_last).to(torch.bfloat16)
    got2 = _lib.to_nhwc_bf16(xc)
    assert torch.equal(got2.cpu(), got.cpu())


@pytest.mark.parametrize("act,C", [(0, 24), (1, 256), (2, 144), (2, 960)])
def test_fused_bn_act_module_matches_torch(act, C):
    """Encoder BatchNorm(+ReLU/ReLU6) on the NHWC kernels vs nn.BatchNorm2d + activation (fp32, CPU)."""
    from pixelpick_b200.deeplab import FusedBNAct
    g = torch.Generator().manual_seed(C)
    x = (torch.randn((3, C, 12, 20), generator=g) * 2 + 1).to(torch.bfloat16)
    go = torch.randn((3, C, 12, 20), generator=g).to(torch.bfloat16)
    ref_bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref_bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref_bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
    m = FusedBNAct(C, act=act)
    m.load_state_dict(ref_bn.state_dict())
    m = m.to(DEV)
    xr = x.float().requires_grad_(True)
    yr = ref_bn(xr)
    yr = F.relu(yr) if act == 1 else (F.relu6(yr) if act == 2 else yr)
    yr.backward(go.float())
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(xg)
    y.backward(go.to(DEV))
    assert y.dtype == torch.bfloat16 and y.shape == yr.shape
    assert (y.float().cpu() - yr.detach()).abs().max().item() < 2e-2 * yr.abs().max().item()
    assert (xg.grad.float().cpu() - xr.grad).abs().max().item() < 3e-2 * xr.grad.abs().max().item()
    assert torch.allclose(m.weight.grad.cpu(), ref_bn.weight.grad, rtol=3e-2, atol=3e-2 * ref_bn.weight.grad.abs().max().item())
    assert torch.allclose(m.bias.grad.cpu(), ref_bn.bias.grad, rtol=3e-2, atol=3e-2 * ref_bn.bias.grad.abs().max().item())
    assert torch.allclose(m.running_mean.cpu(), ref_bn.running_mean, atol=1e-2)
    assert torch.allclose(m.running_var.cpu(), ref_bn.running_var, rtol=2e-2, atol=1e-2)
    assert int(m.num_batches_tracked) == 1
    m.eval(); ref_bn.eval()
    with torch.no_grad():
        ye = m(x.to(DEV).contiguous(memory_format=torch.channels_last)).float().cpu()
        yre = ref_bn(x.float())
        yre = F.relu(yre) if act == 1 else (F.relu6(yre) if act == 2 else yre)
    assert (ye - yre).abs().max().item() < 2e-2 * yre.abs().max().item()


@pytest.mark.parametrize("C,train", [(256, True), (2048, True), (64, False)])
def test_fused_bn_residual_relu_matches_torch(C, train):
    """Bottleneck tail relu(bn3(x) + identity) (resnet_models.py:88-92) as ONE fused pass, forward and backward."""
    from pixelpick_b200.deeplab import FusedBNAct
    g = torch.Generator().manual_seed(C + 1)
    x = (torch.randn((2, C, 9, 14), generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    r = torch.randn((2, C, 9, 14), generator=g).to(torch.bfloat16)
    go = torch.randn((2, C, 9, 14), generator=g).to(torch.bfloat16)
    ref_bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref_bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref_bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
        ref_bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        ref_bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    m = FusedBNAct(C, act=1)
    m.load_state_dict(ref_bn.state_dict())
    m = m.to(DEV)
    if not train:
        m.eval(); ref_bn.eval()
        with torch.no_grad():
            y = m(x.to(DEV).contiguous(memory_format=torch.channels_last), residual=r.to(DEV).contiguous(memory_format=torch.channels_last))
            yr = F.relu(ref_bn(x.float()) + r.float())
        assert (y.float().cpu() - yr).abs().max().item() < 2e-2 * yr.abs().max().item()
        return
    xr, rr = x.float().requires_grad_(True), r.float().requires_grad_(True)
    yr = F.relu(ref_bn(xr) + rr)
    yr.backward(go.float())
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    rg = r.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(xg, residual=rg)
    y.backward(go.to(DEV))
    assert (y.float().cpu() - yr.detach()).abs().max().item() < 2e-2 * yr.abs().max().item()
    # gate disagreements can only happen where |bn(x)+r| is within bf16 rounding of 0: compare away from the kink
    near = ((ref_bn(x.float()) + r.float()).abs() < 0.05)
    ok = ~near
    assert ((rg.grad.float().cpu() - rr.grad).abs() * ok).max().item() < 1e-2 * go.float().abs().max().item()
    assert ((xg.grad.float().cpu() - xr.grad).abs() * ok).max().item() < 4e-2 * xr.grad.abs().max().item()
    assert torch.allclose(m.weight.grad.cpu(), ref_bn.weight.grad, rtol=3e-2, atol=4e-2 * ref_bn.weight.grad.abs().max().item())
    assert torch.allclose(m.bias.grad.cpu(), ref_bn.bias.grad, rtol=3e-2, atol=4e-2 * ref_bn.bias.grad.abs().max().item())


@pytest.mark.parametrize("N,C,Hi,Wi,stride,dil", [(2, 32, 18, 34, 1, 1), (2, 96, 19, 35, 2, 1), (1, 144, 10, 11, 2, 1),
                                                  (2, 960, 12, 20, 1, 2), (1, 576, 7, 9, 1, 1), (1, 8, 3, 3, 1, 1),
                                                  (1, 2048, 9, 9, 1, 4)])
def test_depthwise_conv3x3_fwd_dgrad_wgrad(N, C, Hi, Wi, stride, dil):
    """Hand-written depthwise 3x3 (mobilenet_v2.py:33-35,46-48: groups=C, padding 0 on the pre-padded tensor) vs
    F.conv2d in fp32 on the same bf16-rounded operands."""
    g = torch.Generator().manual_seed(C + Hi)
    x = torch.randn((N, C, Hi, Wi), generator=g).to(torch.bfloat16)
    w = torch.randn((C, 1, 3, 3), generator=g) * 0.4
    xr, wr = x.float().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride, 0, dil, groups=C)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = _lib.dwconv_fwd(xn, w.to(DEV), stride, dil)
    assert tuple(y.shape) == (N, yr.shape[2], yr.shape[3], C)
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert (got - yr.detach()).abs().max().item() < 1e-2 * yr.abs().max().item()
    go = torch.randn(yr.shape, generator=g).to(torch.bfloat16)
    yr.backward(go.float())
    gon = go.permute(0, 2, 3, 1).contiguous().to(DEV)
    dx = _lib.dwconv_dgrad(gon, w.to(DEV), (Hi, Wi), stride, dil).float().permute(0, 3, 1, 2).cpu()
    assert (dx - xr.grad).abs().max().item() < 1e-2 * xr.grad.abs().max().item()
    dw = _lib.dwconv_wgrad(xn, gon, stride, dil).cpu()
    assert dw.shape == w.shape
    assert (dw - wr.grad).abs().max().item() < 2e-3 * wr.grad.abs().max().item() + 1e-4


def test_depthwise_module_matches_conv2d_through_autograd():
    from pixelpick_b200.deeplab import DepthwiseConv3x3
    g = torch.Generator().manual_seed(3)
    m = DepthwiseConv3x3(64, 2, 1)
    ref = torch.nn.Conv2d(64, 64, 3, 2, 0, 1, groups=64, bias=False)
    ref.load_state_dict(m.state_dict())
    x = torch.randn((2, 64, 21, 17), generator=g).to(torch.bfloat16)
    go = torch.randn((2, 64, 10, 8), generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    ref(xr).backward(go.float())
    m = m.to(DEV)
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(xg)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (2, 64, 10, 8)
    y.backward(go.to(DEV))
    assert (xg.grad.float().cpu() - xr.grad).abs().max().item() < 1e-2 * xr.grad.abs().max().item()
    assert (m.weight.grad.cpu() - ref.weight.grad).abs().max().item() < 2e-3 * ref.weight.grad.abs().max().item()


@pytest.mark.parametrize("M,C,relu,p,with_res", [(4 * 16 * 32, 960, 2, 0.0, False), (2 * 64 * 128, 256, 1, 0.5, False),
                                                 (3 * 9 * 14, 2048, 1, 0.0, True), (1000, 24, 0, 0.0, False),
                                                 (4 * 16 * 32, 96, 0, 0.0, True), (300, 160, 2, 0.5, False),  # cluster path
                                                 (32 * 64 * 128, 64, 1, 0.0, True)])
def test_single_launch_bn_matches_separate_kernels(M, C, relu, p, with_res):
    """pp_bn_fwd_fused / pp_bn_bwd_fused (one cooperative launch, grid barrier) == stats + finalize + apply / reduce +
    apply; repeated launches reuse the self-zeroing scratch; also replayed from a CUDA graph."""
    g = torch.Generator().manual_seed(C + M)
    raw = (torch.randn((M, C), generator=g) * 1.5 + 0.4).to(torch.bfloat16).to(DEV)
    dy = torch.randn((M, C), generator=g).to(torch.bfloat16).to(DEV)
    res = torch.randn((M, C), generator=g).to(torch.bfloat16).to(DEV) if with_res else None

    def mk():
        bn = torch.nn.BatchNorm2d(C)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(C, generator=torch.Generator().manual_seed(1)) + 0.5)
            bn.bias.copy_(torch.randn(C, generator=torch.Generator().manual_seed(2)) * 0.3)
        return bn.to(DEV)

    bn_a, bn_b = mk(), mk()
    out_a, out_b = torch.empty_like(raw), torch.empty_like(raw)
    st_a = _lib.bn_finalize(_lib.bn_stats(raw, 0, C), M, bn_a)
    _lib.bn_apply(raw, 0, C, st_a[0], st_a[1], relu, out_a, 0, drop_p=p, seed=7, offset=3, res=res)
    for rep in range(3):  # the scratch must come back zeroed after every launch
        st_b = _lib.bn_fwd_fused(raw, 0, C, bn_b, relu, out_b, 0, drop_p=p, seed=7, offset=3, res=res,
                                 update_running=(rep == 0))
        assert torch.allclose(st_a, st_b, rtol=1e-5, atol=1e-6), rep
        # the fp32 partial sums are combined by atomics in a different order: scale/shift may differ in the last ulp, so
        # a few outputs may round to the neighbouring bf16 value
        d = (out_a.float() - out_b.float()).abs()
        assert float(d.max()) <= 2 ** -7 * float(out_a.float().abs().max()) and float((d > 0).float().mean()) < 1e-2, rep
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, atol=1e-6)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn_b.num_batches_tracked) == 1
    assert float(bn_b._pp_scratch.abs().max()) == 0.0
    ref = _lib.bn_bwd(dy, 0, raw, 0, C, st_a[0], st_a[1], st_a[2], st_a[3], relu, drop_p=p, seed=7, offset=3, res=res)
    sc = _lib.bn_scratch(bn_b, C, DEV)
    for rep in range(2):
        got = _lib.bn_bwd(dy, 0, raw, 0, C, st_a[0], st_a[1], st_a[2], st_a[3], relu, drop_p=p, seed=7, offset=3, res=res,
                          scratch=sc)
        assert torch.allclose(ref[1], got[1], rtol=2e-4, atol=2e-4 * float(ref[1].abs().max())), rep  # atomics order
        assert (ref[0].float() - got[0].float()).abs().max().item() <= 2e-2 * ref[0].float().abs().max().item()
        if with_res:
            assert torch.equal(ref[2], got[2])
    assert float(sc.abs().max()) == 0.0
    # CUDA-graph capture of the cooperative launches
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    out_g = torch.empty_like(raw)
    with torch.cuda.stream(stream):
        _lib.bn_fwd_fused(raw, 0, C, bn_b, relu, out_g, 0, drop_p=p, seed=7, offset=3, res=res, update_running=False)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            _lib.bn_fwd_fused(raw, 0, C, bn_b, relu, out_g, 0, drop_p=p, seed=7, offset=3, res=res, update_running=False)
    out_g.zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    dg = (out_a.float() - out_g.float()).abs()
    assert float(dg.max()) <= 2 ** -7 * float(out_a.float().abs().max()) and float((dg > 0).float().mean()) < 1e-2


@pytest.mark.parametrize("N,H,W,C", [(2, 128, 256, 64), (1, 33, 47, 16), (3, 8, 8, 8)])
def test_maxpool_stem_matches_torch(N, H, W, C):
    """pp_maxpool3x3s2_fwd / _bwd vs nn.MaxPool2d(3, 2, 1) (resnet_models.py:116): forward exact, backward = the gradient routed
    to the first maximum of each window (odd sizes, ties and NaN included)."""
    from pixelpick_b200 import _lib
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((N, C, H, W), generator=g) * 2).to(torch.bfloat16)
    x[0, 0, 0, 0] = float("nan")
    x[0, 1, 2:4, 2:4] = 1.5  # a tie: the first maximum in row-major window order wins
    dev = torch.device("cuda:0")
    xr = x.to(dev).float().requires_grad_(True)
    want = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    dy = torch.randn(want.shape, generator=g).to(torch.bfloat16)
    want.backward(dy.to(dev).float())
    xn = x.permute(0, 2, 3, 1).contiguous().to(dev)
    y, code = _lib.maxpool3x3s2_fwd(xn)
    got = y.permute(0, 3, 1, 2).float()
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want.detach(), nan=-7.0))
    dx = _lib.maxpool3x3s2_bwd(dy.permute(0, 2, 3, 1).contiguous().to(dev), code, (H, W))
    ref = xr.grad
    d = (dx.permute(0, 3, 1, 2).float() - ref).abs()
    assert d.max().item() <= 2 ** -6 * ref.abs().max().item()  # up to 4 bf16 gradients summed in fp32, rounded once

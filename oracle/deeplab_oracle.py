"""ORACLE (test infrastructure, NOT product code) — CPU fp32 restatement of the reference DeepLabv3+ forward
and of one training step's loss, as plain functional PyTorch over a reference-named state_dict.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.

Follows NoelShin/PixelPick @ 43c2981: networks/mobilenet_v2.py:15-137, networks/aspp.py:49-79,
networks/decoders.py:107-123, networks/deeplab.py:43-61, networks/backbones/resnet_models.py:58-171 +
resnet_backbone.py:42-104, model.py:108-116.  The arithmetic is torch's CPU kernels, exactly what the reference
executes.  Pinned against the imported reference modules by tests/golden/make_golden_model.py ->
tests/golden/model_golden.npz (checked in tests/test_oracle_model_golden.py).
"""
import zlib

import torch
import torch.nn.functional as F

MNV2_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]


def synthetic_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Deterministic weights for a {name: shape} map, independent of module construction order: every tensor is
    drawn from its own generator seeded by crc32(name) ^ seed.  Conv weights ~ N(0, 2/fan_in) (kaiming, as the
    reference initialises them), BN weight ~ U(0.5, 1.5), BN bias ~ N(0, 0.1), running_mean ~ N(0, 0.1),
    running_var ~ U(0.5, 1.5)."""
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        # the reference registers MobileNetV2.features[0:4] / [4:] a second time as low_level_features /
        # high_level_features (mobilenet_v2.py:125-126): aliases must carry identical values
        canon = name
        for alias, shift in (("backbone.low_level_features.", 0), ("backbone.high_level_features.", 0)):  # slices keep indices
            if name.startswith(alias):
                idx, rest = name[len(alias):].split(".", 1)
                canon = f"backbone.features.{int(idx) + shift}.{rest}"
        g = torch.Generator().manual_seed((zlib.crc32(canon.encode()) ^ seed) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=torch.long)
        elif name.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif name.endswith("weight"):
            t = torch.rand(shape, generator=g) + 0.5
        else:
            t = torch.randn(shape, generator=g) * 0.1
        sd[name] = t
    return sd


class _RoundStorage(torch.autograd.Function):
    """x -> x rounded to `dtype` and back to fp32, in BOTH directions (the activation and its gradient are stored in
    `dtype` between layers).  Used only by the storage-precision mode below."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).float(), None


class _RoundWeight(torch.autograd.Function):
    """forward: the bf16 operand copy of an fp32 master weight; backward: identity (weight gradients are fp32)."""

    @staticmethod
    def forward(ctx, w, dtype):
        return w.to(dtype).float()

    @staticmethod
    def backward(ctx, g):
        return g, None


class _Ctx:
    """storage_dtype = None: the reference's fp32 arithmetic, exactly.  storage_dtype = torch.bfloat16: the SAME graph with
    every activation (and, in training, its gradient) rounded to bf16 where a layer hands it to the next one - after each
    BatchNorm (+ activation: ReLU / ReLU6 commute with the rounding) - and dense conv weights rounded to bf16, everything
    else (accumulation, statistics, normalisation) in fp32.  That is the error a bf16-STORAGE implementation of the reference
    carries by construction; the GPU parity tests bound the CUDA path by it (error budget)."""

    def __init__(self, sd, training, dropout_p_scale=1.0, storage_dtype=None):
        self.sd, self.training = sd, training
        self.new_stats = {}
        self.storage_dtype = storage_dtype
        if storage_dtype is not None:
            self.sd = {k: (_RoundWeight.apply(v, storage_dtype) if v.dim() == 4 and v.shape[1] > 1 and v.is_floating_point() else v)
                       for k, v in sd.items()}  # dense conv weights; depthwise (in/groups = 1) weights stay fp32 as in the kernels

    def store(self, x):
        return x if self.storage_dtype is None else _RoundStorage.apply(x, self.storage_dtype)

    def bn(self, x, prefix, eps=1e-5, momentum=0.1):
        sd = self.sd
        if self.training:
            x = self.store(x)  # training keeps the raw conv output (BatchNorm's input) for the backward pass: stored too
        rm, rv = sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone()
        y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], self.training, momentum, eps)
        if self.training:
            self.new_stats[prefix] = (rm, rv)
        return self.store(y)


def _fixed_padding(x, dilation):  # mobilenet_v2.py:15-21
    pad_total = 2 * dilation
    beg = pad_total // 2
    return F.pad(x, (beg, pad_total - beg, beg, pad_total - beg))


def mobilenet_v2(c: _Ctx, x, output_stride=16):
    """mobilenet_v2.py:69-137 -> (high [320, 1/16], low [24, 1/4])."""
    sd = c.sd
    p = "backbone.features."
    x = F.relu6(c.bn(F.conv2d(x, sd[p + "0.0.weight"], None, 2, 1), p + "0.1"))
    inp, cur, rate, idx = 32, 2, 1, 1
    low = None
    for t, ch, n, s in MNV2_SETTING:
        if cur == output_stride:
            stride, dil = 1, rate
            rate *= s
        else:
            stride, dil = s, 1
            cur *= s
        for i in range(n):
            st = stride if i == 0 else 1
            q = f"{p}{idx}.conv."
            hidden = round(inp * t)
            y = _fixed_padding(x, dil)  # padding BEFORE the expansion conv (mobilenet_v2.py:60-61)
            j = 0
            if t != 1:
                y = F.relu6(c.bn(F.conv2d(y, sd[q + "0.weight"]), q + "1"))
                j = 3
            y = F.relu6(c.bn(F.conv2d(y, sd[q + f"{j}.weight"], None, st, 0, dil, hidden), q + f"{j + 1}"))
            y = c.bn(F.conv2d(y, sd[q + f"{j + 3}.weight"]), q + f"{j + 4}")
            x = c.store(x + y) if (st == 1 and inp == ch) else y
            inp = ch
            if idx == 3:
                low = x  # features[0:4]
            idx += 1
    return x, low


def resnet50_dilated8(c: _Ctx, x, prefix="backbone."):
    """resnet_models.py + resnet_backbone.py (dilate_scale=8, multi_grid=None) -> (c5 [2048, 1/8], c2 [256, 1/4])."""
    sd = c.sd
    x = F.relu(c.bn(F.conv2d(x, sd[prefix + "prefix.conv1.weight"], None, 2, 3), prefix + "prefix.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    c2 = None
    for li, (blocks, stride0, dil_rest) in enumerate([(3, 1, 1), (4, 2, 1), (6, 2, 2), (3, 2, 4)], start=1):
        for b in range(blocks):
            q = f"{prefix}layer{li}.{b}."
            stride = stride0 if b == 0 else 1
            dil = 1
            if li >= 3:  # _nostride_dilate: stride removed, first block dilate//2, rest dilate
                dil = dil_rest // 2 if b == 0 else dil_rest
                stride = 1
            idn = x
            y = F.relu(c.bn(F.conv2d(x, sd[q + "conv1.weight"]), q + "bn1"))
            y = F.relu(c.bn(F.conv2d(y, sd[q + "conv2.weight"], None, stride, dil, dil), q + "bn2"))
            y = c.bn(F.conv2d(y, sd[q + "conv3.weight"]), q + "bn3")
            if q + "downsample.0.weight" in sd:
                idn = c.bn(F.conv2d(x, sd[q + "downsample.0.weight"], None, stride), q + "downsample.1")
            x = c.store(F.relu(y + idn))
        if li == 1:
            c2 = x
    return x, c2


def head(c: _Ctx, high, low, dilations, drop=(0.0, 0.0, 0.0)):
    """aspp.py:64-79 + deeplab.py:48-53 + decoders.py:120-123 -> 1/4-resolution logits and embedding.
    drop = dropout probabilities (ASPP, head #1, head #2); parity runs use 0 (GPU Philox != CPU MT)."""
    sd = c.sd
    xs = []
    for i, d in enumerate(dilations, start=1):
        w = sd[f"aspp.aspp{i}.atrous_conv.weight"]
        pad = 0 if w.shape[2] == 1 else d
        xs.append(F.relu(c.bn(F.conv2d(high, w, None, 1, pad, d), f"aspp.aspp{i}.bn")))
    g = F.adaptive_avg_pool2d(high, 1)
    g = F.relu(c.bn(F.conv2d(g, sd["aspp.global_avg_pool.1.weight"]), "aspp.global_avg_pool.2"))
    xs.append(F.interpolate(g, size=high.shape[2:], mode="bilinear", align_corners=True))
    x = F.relu(c.bn(F.conv2d(torch.cat(xs, dim=1), sd["aspp.conv1.weight"]), "aspp.bn1"))
    x = F.dropout(x, drop[0], c.training)
    ll = F.relu(c.bn(F.conv2d(low, sd["low_level_conv.0.weight"]), "low_level_conv.1"))
    x = F.interpolate(x, size=ll.shape[2:], mode="bilinear", align_corners=True)
    x = torch.cat((x, ll), dim=1)
    x = F.relu(c.bn(F.conv2d(x, sd["seg_head.segment_head.0.weight"], None, 1, 1), "seg_head.segment_head.1"))
    x = F.dropout(x, drop[1], c.training)
    x = F.relu(c.bn(F.conv2d(x, sd["seg_head.segment_head.4.weight"], None, 1, 1), "seg_head.segment_head.5"))
    emb = F.dropout(x, drop[2], c.training)
    pred = F.conv2d(emb, sd["seg_head.classifier.weight"], sd["seg_head.classifier.bias"])
    return pred, emb


def deeplab_forward(sd, x, backbone="mobilenet", training=False, drop=(0.0, 0.0, 0.0), return_ctx=False, storage_dtype=None):
    """deeplab.py:43-61 -> dict(pred=[B,C,H,W] full-res logits, lowres=[B,C,H/4,W/4] head logits).
    storage_dtype: see _Ctx (None = the reference's fp32 path)."""
    c = _Ctx(sd, training, storage_dtype=storage_dtype)
    x = c.store(x)
    if backbone == "mobilenet":
        high, low = mobilenet_v2(c, x)
        dil = [1, 6, 12, 18]
    else:
        high, low = resnet50_dilated8(c, x)
        dil = [1, 12, 24, 36]
    lowres, emb = head(c, high, low, dil, drop)
    pred = F.interpolate(lowres, size=x.shape[2:], mode="bilinear", align_corners=True)
    out = {"pred": pred, "lowres": lowres, "emb_lowres": emb, "high": high, "low": low}
    if return_ctx:
        out["ctx"] = c
    return out


def sparse_ce_loss(pred, y, queries, ignore_index):
    """model.py:108-116: unlabelled targets -> ignore_index, then mean CE over the labelled pixels."""
    y = y.clone()
    y.flatten()[~queries.flatten().bool()] = ignore_index
    return F.cross_entropy(pred, y, ignore_index=ignore_index)


def reference_init_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Host-independent weights with the DISTRIBUTIONS of the reference's own initialisation (torch's seeded CPU normal
    stream depends on the host's vector width, NumPy's RandomState does not): MobileNetV2 / ASPP / SegmentHead conv
    weights kaiming-normal, std = sqrt(2 / fan_in) (mobilenet_v2.py:149-155, aspp.py:81-88, decoders.py:125-132);
    ResNet conv weights N(0, sqrt(2 / (k*k*out))) (resnet_models.py:131-137); low_level_conv and every conv bias keep
    nn.Conv2d's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (deeplab.py:23-26); BatchNorm weight 1 / bias 0, running
    statistics 0 / 1.  This is the well-conditioned network a real run starts from (unlike synthetic_state_dict, whose
    random BatchNorm scales are a stress case)."""
    import numpy as np
    sd = {}
    canon_cache = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        canon = name
        for alias in ("backbone.low_level_features.", "backbone.high_level_features."):
            if name.startswith(alias):
                idx, rest = name[len(alias):].split(".", 1)
                canon = f"backbone.features.{int(idx)}.{rest}"
        if canon in canon_cache:
            sd[name] = canon_cache[canon].clone()
            continue
        rs = np.random.RandomState((zlib.crc32(canon.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=torch.long)
        elif name.endswith("running_var"):
            t = torch.ones(shape)
        elif name.endswith("running_mean"):
            t = torch.zeros(shape)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if canon.startswith("low_level_conv."):
                b = 1.0 / fan_in ** 0.5
                t = torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))
            elif canon.startswith("backbone.") and ("layer" in canon or "prefix" in canon):
                std = (2.0 / (shape[2] * shape[3] * shape[0])) ** 0.5
                t = torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))
            else:
                t = torch.from_numpy((rs.standard_normal(size=shape) * (2.0 / fan_in) ** 0.5).astype(np.float32))
        elif canon == "seg_head.classifier.bias":
            b = 1.0 / 256 ** 0.5
            t = torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))
        elif name.endswith("weight"):
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        canon_cache[canon] = t
        sd[name] = t
    return sd


def train_steps(sd, batches, backbone, ignore_index, lr=5e-4, weight_decay=2e-4, drop=(0.0, 0.0, 0.0), storage_dtype=None):
    """k optimisation steps of model.py:103-122 (one per (x, y, queries) batch) with the cs optimiser of
    utils/utils.py:117-141 (Adam, encoder lr/10, torch default betas/eps) on a COPY of `sd`; BatchNorm running statistics
    are updated like nn.BatchNorm2d does.  Returns (new state dict, [loss_k], gradients of step 0 by name)."""
    sd = {k: v.clone() for k, v in sd.items()}
    names = [k for k in sorted(sd) if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
             and not k.startswith(("backbone.low_level_features.", "backbone.high_level_features."))]
    for k in names:
        sd[k].requires_grad_(True)
    enc = [sd[k] for k in names if k.startswith("backbone.")]
    rest = [sd[k] for k in names if not k.startswith("backbone.")]
    opt = torch.optim.Adam([{"params": enc, "lr": lr / 10, "weight_decay": weight_decay},
                            {"params": rest, "lr": lr, "weight_decay": weight_decay}])
    losses, grads0 = [], None
    for i, (x, y, q) in enumerate(batches):
        out = deeplab_forward(sd, x, backbone=backbone, training=True, drop=drop, return_ctx=True, storage_dtype=storage_dtype)
        loss = sparse_ce_loss(out["pred"], y, q, ignore_index)
        opt.zero_grad()
        loss.backward()
        if i == 0:
            grads0 = {k: sd[k].grad.detach().clone() for k in names}
        opt.step()
        losses.append(float(loss.detach()))
        with torch.no_grad():
            for prefix, (rm, rv) in out["ctx"].new_stats.items():
                sd[prefix + ".running_mean"].copy_(rm)
                sd[prefix + ".running_var"].copy_(rv)
                if prefix + ".num_batches_tracked" in sd:
                    sd[prefix + ".num_batches_tracked"] += 1
            # MobileNetV2's aliased registrations follow their canonical tensors
            for k in sd:
                for alias in ("backbone.low_level_features.", "backbone.high_level_features."):
                    if k.startswith(alias):
                        idx, rest_k = k[len(alias):].split(".", 1)
                        sd[k] = sd[f"backbone.features.{int(idx)}.{rest_k}"]
    return {k: v.detach() for k, v in sd.items()}, losses, grads0

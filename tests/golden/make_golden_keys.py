"""state_dict layout of the reference networks (build container only; needs /root/reference):

    python tests/golden/make_golden_keys.py  ->  tests/golden/state_dict_keys.json

{"mobilenet": [[key, shape, dtype], ...], "resnet": [...]} in state_dict order, for DeepLab(args) (networks/deeplab.py) and the
RN50-DeepLabv3+ composition of reference modules defined in make_golden_model.py.  A checkpoint written by the reference
(`torch.save({"model": model.state_dict()})`, model.py:207-212) must load into the drop-in with `strict=True`, and back.

"init_seed0": for DeepLab(args) built right after torch.manual_seed(0), an exact checksum (sum of the raw bit patterns) of every
tensor - the drop-in consumes the RNG in the same order (each ASPP branch draws twice, low_level_conv keeps the default
initialisation), so a seeded run starts from the reference's weights bit for bit.  "canary" is the same checksum of
torch.randn(4096) / kaiming_normal_ on this host: torch's CPU normal stream depends on the host's vector width, the test skips
the comparison where the canary differs."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_model as mg  # noqa: E402  (imports the reference modules, stubs the pretrained-weight download)


def layout(m):
    return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]


def bits_checksum(t):
    import numpy as np
    a = t.detach().contiguous().cpu().numpy()
    a = a.view(np.uint32) if a.dtype == np.float32 else a.astype(np.int64).view(np.uint64)
    return int(a.astype(np.uint64).sum() % (1 << 63))


def canary():
    import torch
    torch.manual_seed(123)
    w = torch.empty(64, 32, 3, 3)
    torch.nn.init.kaiming_normal_(w)
    return [bits_checksum(torch.randn(4096)), bits_checksum(w)]


def seeded_init(ctor):
    import torch
    torch.manual_seed(0)
    return {k: bits_checksum(v) for k, v in ctor().state_dict().items()}


if __name__ == "__main__":
    out = {"init_seed0": seeded_init(lambda: mg.RefDeepLab(mg.ARGS)), "canary": canary(),
           "mobilenet": layout(mg.RefDeepLab(mg.ARGS)), "resnet": layout(mg.RefRN50DeepLab())}
    p = os.path.join(HERE, "state_dict_keys.json")
    json.dump(out, open(p, "w"))
    print("wrote", p, {k: len(v) for k, v in out.items()}, os.path.getsize(p), "bytes")

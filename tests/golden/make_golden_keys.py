"""state_dict layout of the reference networks (build container only; needs /root/reference):

    python tests/golden/make_golden_keys.py  ->  tests/golden/state_dict_keys.json

{"mobilenet": [[key, shape, dtype], ...], "resnet": [...]} in state_dict order, for DeepLab(args) (networks/deeplab.py) and the
RN50-DeepLabv3+ composition of reference modules defined in make_golden_model.py.  A checkpoint written by the reference
(`torch.save({"model": model.state_dict()})`, model.py:207-212) must load into the drop-in with `strict=True`, and back."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_model as mg  # noqa: E402  (imports the reference modules, stubs the pretrained-weight download)


def layout(m):
    return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]


if __name__ == "__main__":
    out = {"mobilenet": layout(mg.RefDeepLab(mg.ARGS)), "resnet": layout(mg.RefRN50DeepLab())}
    p = os.path.join(HERE, "state_dict_keys.json")
    json.dump(out, open(p, "w"))
    print("wrote", p, {k: len(v) for k, v in out.items()}, os.path.getsize(p), "bytes")

"""GPU: run-to-run reproducibility of ONE train step (same weights, same batch): loss and gradient cosines between repetitions."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import test_dp_gpu as T
from pixelpick_b200.loss import sparse_cross_entropy
dev = torch.device("cuda:0")
for backbone in ("mobilenet", "resnet"):
    m = T._model(backbone, dev)
    x, y, q = T._batch(100, 300)
    gs = []
    for rep in range(5):
        m.zero_grad(set_to_none=True)
        loss = sparse_cross_entropy(m.forward_lowres(x.to(dev)), y.to(dev), q.to(dev), T.C)
        loss.backward()
        torch.cuda.synchronize()
        gs.append({n: p.grad.detach().float().clone() for n, p in m.named_parameters()})
        print(backbone, rep, "loss", loss.item())
    for a, b in ((0, 1), (1, 2), (2, 3), (3, 4)):
        cs = []
        for n in gs[a]:
            ga, gb = gs[a][n].flatten(), gs[b][n].flatten()
            if ga.norm() > 0: cs.append((torch.dot(ga, gb) / (ga.norm() * gb.norm())).item())
        print(backbone, "run", a, "vs", b, "cos min", min(cs), "mean", np.mean(cs))

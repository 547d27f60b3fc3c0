"""Per-layer gradient comparison GPU vs oracle for ResNet-20 (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dfxp as O
from lbt_b200 import dfxp as D, models as M
from lbt_b200.trainer import Trainer


def run(fused, exact):
    rng = np.random.default_rng(0)
    om = O.CIFAR10_Resnet20(8, weight_decay=2e-4, noise=O.PhiloxNoise(5), seed=1, exact=exact)
    pm = M.CIFAR10_Resnet20(8, weight_decay=2e-4, seed=5).cuda()
    for m in pm.modules():
        if isinstance(m, D.BatchNorm2d_q):
            m.fused = fused
    for ov, pv in zip(om.variables(), pm.parameters()):
        pv.data.copy_(ov.detach())
    tr = Trainer(pm, lr=1e-2, momentum=0.9)
    X = torch.from_numpy((rng.standard_normal((16, 32, 32, 3)) * 0.5).astype(np.float32))
    y = torch.from_numpy(rng.integers(0, 10, 16))
    om.forward(X)
    # oracle backward, recording the gradient entering each top-level layer
    om.loss, grad = om.loss_and_grad(y)
    ograds = {}
    for i in reversed(range(len(om.layers))):
        ograds[i] = grad
        grad = om.layers[i].backward(grad, True)
    # gpu: same dlogits, hooks on top-level layers.  oracle layer list has separate ReLU_q after conv1-bn.
    pgrads = {}
    def mk(i):
        def hook(mod, gin, gout):
            pgrads[i] = gout[0].detach()
        return hook
    for i, l in enumerate(pm.layers):
        l.register_full_backward_hook(mk(i))
    tr.flat_g.zero_()
    logits = pm(X.permute(0, 3, 1, 2).cuda())
    print('fused=%s exact=%s logits equal: %s' % (fused, exact, torch.equal(logits.detach().cpu(), om.logits)))
    logits.backward(ograds[len(om.layers) - 1].cuda())
    # map gpu layer index -> oracle layer index (oracle has one extra ReLU_q at index 2)
    for pi in reversed(range(len(pm.layers))):
        oi = pi if pi < 2 else pi + 1
        if pi not in pgrads:
            continue
        gp = pgrads[pi]
        gp = gp.permute(0, 2, 3, 1).contiguous().cpu() if gp.dim() == 4 else gp.cpu()
        go = ograds[oi]
        rel = float((gp - go).norm() / (go.norm() + 1e-30))
        nd = int((gp != go).sum())
        print('  grad entering gpu layer %2d (%s): rel %.3e  differing %d / %d  |go| %.3e max|go| %.3e'
              % (pi, type(pm.layers[pi]).__name__, rel, nd, go.numel(), float(go.norm()), float(go.abs().max())))


run(False, True)
run(True, True)

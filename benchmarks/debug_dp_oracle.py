"""Stress aid: first training step of freshly built CIFAR10_Model replicas (dirty allocator) vs the exact oracle, many times,
under variations (side-stream overlap, programmatic dependent launch) — hunts an intermittent first-step mismatch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dfxp as O  # noqa: E402
from lbt_b200 import _lib, models as M  # noqa: E402
from lbt_b200.trainer import Trainer  # noqa: E402

SEED = 13
b = 8
rng = np.random.default_rng(42)
X = torch.from_numpy((rng.standard_normal((b, 32, 32, 3)) * 0.5).astype(np.float32))
y = torch.from_numpy(rng.integers(0, 10, b))
om = O.CIFAR10_Model(8, weight_decay=2e-4, dropout=1.0, noise=O.PhiloxNoise(SEED), seed=1, exact=True)
w0 = [v.detach().clone() for v in om.variables()]
om.forward(X)
om.backward(y)
want = [g.detach().clone() for g, _ in om.grads_and_vars()]
Xd, yd = X.permute(0, 3, 1, 2).cuda(), y.cuda()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for variant in ('default', 'no_overlap', 'no_pdl'):
    _lib.lib().lbt_set_pdl(0 if variant == 'no_pdl' else 1)
    bad = {}
    for it in range(iters):
        pm = M.CIFAR10_Model(8, weight_decay=2e-4, dropout=1.0, seed=SEED).cuda()
        for ov, pv in zip(w0, pm.parameters()):
            pv.data.copy_(ov)
        tr = Trainer(pm, lr=1e-2, momentum=0.9)
        if variant == 'no_overlap':
            pm.runtime.overlap = False
        tr.forward_backward(Xd, yd)
        for (name, _), p, go in zip(pm.named_parameters(), tr.params, want):
            n = int((p.grad.cpu() != go).sum())
            if n:
                bad.setdefault(name, []).append((it, n))
        # dirty the allocator: garbage of assorted sizes on the main stream
        junk = [torch.full((s,), 0x7f, dtype=torch.uint8, device='cuda') for s in (1 << 12, 1 << 16, 1 << 20, 1 << 22, 3 << 20, 205000 * 8)]
        del junk, tr, pm
    print(variant, 'iterations', iters, 'mismatches', {k: v[:6] for k, v in bad.items()})
_lib.lib().lbt_set_pdl(1)

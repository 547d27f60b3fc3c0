"""Max-pool forward / backward on the ImageNet stem shape (256 x 112 x 112 x 64, 3x3 / 2 'SAME'): us per launch and fraction of
the measured HBM peak.  A/B: LBT_POOL_J=1|2|4 (channel groups per thread of the backward kernel)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import dfxp as D  # noqa: E402

N, C, H = 256, 64, 112
x = torch.relu(torch.randn(N, C, H, H, device='cuda')).contiguous(memory_format=torch.channels_last).requires_grad_(True)
pool = D.MaxPool_q(3, 2, 'SAME')
y = pool(x)
g = torch.randn_like(y)
for _ in range(3):
    x.grad = None
    y.backward(g, retain_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    x.grad = None
    y.backward(g, retain_graph=True)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
nbytes = x.numel() * 4 + y.numel() * 5
peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
print(json.dumps({'LBT_POOL_J': os.environ.get('LBT_POOL_J', 'default'), 'bwd_us_incl_autograd': round(us, 1),
                  'gbs': round(nbytes / us / 1e3, 1), 'frac_hbm': round(nbytes / us / 1e3 / peak, 3)}))

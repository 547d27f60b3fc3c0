"""One eager ResNet-20 training step (batch 256) inside cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists; or (--quant LOG2N) a few large quantiser launches for a
`--set full` capture.  Never report a number measured under the profiler."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import models as M, quantizer as Q  # noqa: E402
from lbt_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--quant', type=int, default=0)
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--workload', default='CIFAR10_Resnet20')
a = ap.parse_args()

if a.quant:
    n = 1 << a.quant
    x = torch.randn(256, n // 256, device='cuda') * 1.3
    ib = torch.tensor(2, dtype=torch.int32, device='cuda')
    cnt = Q.new_counters('cuda')
    out = torch.empty_like(x)
    om = torch.empty_like(x, dtype=torch.int8)
    for _ in range(3):
        Q.quantize(x, 8, ib, mode=Q.ROUND_PHILOX, seed=1, offset=2, want_fp32=False, mant_kind=Q.MANT_S8, counters=cnt,
                   update_range=False, out_mant=om)
        Q.quantize(x, 8, ib, mode=Q.ROUND_PHILOX, seed=1, offset=2, want_fp32=True, counters=cnt, update_range=False, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    Q.quantize(x, 8, ib, mode=Q.ROUND_PHILOX, seed=1, offset=2, want_fp32=False, mant_kind=Q.MANT_S8, counters=cnt,
               update_range=False, out_mant=om)
    Q.quantize(x, 8, ib, mode=Q.ROUND_PHILOX, seed=1, offset=2, want_fp32=True, counters=cnt, update_range=False, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    torch.manual_seed(0)
    kw = dict(weight_decay=2e-4, seed=0)
    img = 32
    if a.workload.startswith('Resnet'):
        img = 224
        kw.update(image=224, num_classes=1000)
    model = getattr(M, a.workload)(8, **kw).cuda()
    tr = Trainer(model, lr=1e-2, momentum=0.9)
    X = (torch.randn(a.batch, img, img, 3, device='cuda') * 0.5).permute(0, 3, 1, 2)
    y = torch.randint(0, 10, (a.batch,), device='cuda')
    for _ in range(3):
        tr.step(X, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = tr.step(X, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('loss', float(loss))

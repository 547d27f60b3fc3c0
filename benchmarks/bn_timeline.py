"""Phase timeline of CTA 0 of lbt_bn_bwd_apply on one ResNet-20 stage-1 tensor — a debugging aid."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib  # noqa: E402

N, H, W, C = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 32, 32, 16))]
n_inner = H * W * C
kg1 = torch.randint(-100, 100, (N, n_inner), dtype=torch.int8, device='cuda')
k1 = torch.randint(-100, 100, (N, n_inner), dtype=torch.int8, device='cuda')
ib = torch.tensor(2, dtype=torch.int32, device='cuda')
fs = torch.randint(1, 1000, (2 * C,), dtype=torch.int64, device='cuda')
fs[C:] += 10 ** 6
bs = torch.randint(-1000, 1000, (4 * C,), dtype=torch.int64, device='cuda')
gm = torch.empty_like(kg1)
cnt = torch.zeros(4, dtype=torch.int64, device='cuda')
qs = _lib.QSiteStruct(bits=8, stats_minmax=1, ib=ib.data_ptr(), noise=0, seed=1, offset=5, dev_step=0, counters=cnt.data_ptr())
dbg = torch.zeros(8, dtype=torch.int64, device='cuda')
run = lambda: _lib.call('lbt_bn_bwd_apply', _lib.ptr(kg1), _lib.ptr(k1), N, n_inner, C, 8, _lib.ptr(ib), _lib.ptr(fs), 1e-5, 8,
                        _lib.ptr(ib), _lib.ptr(bs), None, ctypes.addressof(qs), _lib.ptr(gm), _lib.stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
_lib.lib().lbt_bn_set_debug(dbg.data_ptr())
run()
torch.cuda.synchronize()
_lib.lib().lbt_bn_set_debug(None)
t = dbg.cpu().tolist()
for i, name in enumerate(['start', 'fp64 prologue done (thread 0)', 'after syncthreads', 'rows done', 'counters published']):
    print('%-32s +%7.2f us' % (name, (t[i] - t[0]) / 1e3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.graph(g, stream=s):
    for _ in range(10):
        run()
torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
print('back-to-back in a graph: %.2f us per launch' % (a.elapsed_time(b) * 100))

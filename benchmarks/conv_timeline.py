"""Phase timeline (globaltimer) of CTA 0 of the cp.async-gather conv kernel on one small layer — a debugging aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib, dfxp, quantizer as Q  # noqa: E402

N, H, W, Cin, Cout, k, s = [int(v) for v in (sys.argv[1:8] if len(sys.argv) >= 8 else (256, 8, 8, 64, 64, 3, 1))]
OH, pt, _ = dfxp.same_pad(H, k, s)
OW, pl, _ = dfxp.same_pad(W, k, s)
x = torch.randint(0, 256, (N, H, W, Cin), dtype=torch.uint8, device='cuda')
Kf = k * k * Cin
wt = torch.zeros(Cout, dfxp._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
wt.copy_(torch.randint(-128, 128, (Cout, Kf), dtype=torch.int8, device='cuda'))
ib = torch.tensor(2, dtype=torch.int32, device='cuda')
y = torch.empty(N * OH * OW, Cout, dtype=torch.float32, device='cuda')
dbg = torch.zeros(64, dtype=torch.int64, device='cuda')
run = lambda: dfxp._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, s, s, pt, pl, OH, OW, ib, ib, -15, None, y)
for _ in range(3):
    run()
torch.cuda.synchronize()
_lib.lib().lbt_conv_ldg_set_debug(dbg.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
run()
b.record()
torch.cuda.synchronize()
_lib.lib().lbt_conv_ldg_set_debug(None)
t = dbg.cpu().tolist()
names = {0: 'kernel start', 1: 'prologue done', 2: 'load st0 issued', 3: 'load st1', 4: 'load st2', 5: 'load st3', 6: 'load st4',
         7: 'load st5+', 8: 'mma st0 full', 9: 'mma st1', 10: 'mma st2', 11: 'mma st3', 12: 'mma st4', 13: 'mma st5+',
         14: 'mma tile committed', 15: 'epi got acc', 16: 'epi tile done', 17: 'all warps done', 18: 'tmem freed'}
print('event time %.1f us' % (a.elapsed_time(b) * 1e3))
t[0] = t[0] or t[1]
for base, nm in ((44, 'loader issued tile'), (32, 'mma full tile'), (20, 'epilogue done tile')):
    print(nm, ' '.join('%.2f' % ((t[base + i] - t[0]) / 1e3) for i in range(10) if t[base + i]))
for i in sorted(names):
    if t[i]:
        print('%-22s +%7.2f us' % (names[i], (t[i] - t[0]) / 1e3))

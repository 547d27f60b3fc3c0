"""k_splits sweep of the TMA wgrad kernel on the ResNet-20 shapes (graph-captured launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib, dfxp, quantizer as Q  # noqa: E402
from benchmarks.gemm_bench import timeit, rand_i8  # noqa: E402

_lib.lib().lbt_conv_set_path(0)
for (N, H, W, Ci, Co, k, s) in [(256, 32, 32, 16, 16, 3, 1), (256, 16, 16, 32, 32, 3, 1), (256, 8, 8, 64, 64, 3, 1), (256, 32, 32, 16, 32, 3, 2)]:
    OH, pt, _ = dfxp.same_pad(H, k, s)
    OW, pl, _ = dfxp.same_pad(W, k, s)
    x = rand_i8(N, H, W, Ci, unsigned=True)
    g = rand_i8(N, OH, OW, Co)
    acc = torch.zeros(k * k * Ci, Co, dtype=torch.int64, device='cuda')
    res = []
    for sp in (0, 8, 16, 32, 64, 128, 256):
        t = timeit(lambda: _lib.call('lbt_conv_i8_wgrad', _lib.ptr(x), Q.MANT_U8, N, H, W, Ci, _lib.ptr(g), Q.MANT_S8, Co, k, k, s, s,
                                     pt, pl, OH, OW, _lib.ptr(acc), 1, sp, _lib.stream()), iters=10)
        res.append('%d:%.1f' % (sp, t * 1e6))
    print('%dx%d %d->%d /%d  splits:us  %s' % (H, W, Ci, Co, s, '  '.join(res)))

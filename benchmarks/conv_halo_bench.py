"""A/B of the stride-1 3x3 convolution kernels on ResNet-18's 64- and 128-channel shapes: conv_halo.cu (halo patches through
the TMA engine, resident filter) against the im2col-mode TMA kernel of conv_i8.cu, fp32 and fused re-quantising epilogues.
Graph-captured back-to-back launches, CUDA-event timing.

    python benchmarks/conv_halo_bench.py [--iters 20] [--only l1] [--profile]      (--profile: one launch of each inside
    cudaProfilerStart/Stop for `ncu --profile-from-start off`; numbers printed under a profiler are not bench values)
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib, dfxp as D, quantizer as Q  # noqa: E402

SHAPES = [('l1 3x3 64->64 56x56', 256, 56, 56, 64, 64, 3), ('l2 3x3 128->128 28x28', 256, 28, 28, 128, 128, 3)]


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--only', default='')
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    rows = []
    for name, N, H, W, Cin, Cout, k in SHAPES:
        if a.only and a.only not in name:
            continue
        OH, pt, _ = D.same_pad(H, k, 1)
        OW, pl, _ = D.same_pad(W, k, 1)
        x = torch.randint(0, 256, (N, H, W, Cin), dtype=torch.uint8, device='cuda')
        Kf = k * k * Cin
        wt = torch.zeros(Cout, D._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
        wt.copy_(torch.randint(-128, 128, (Cout, Kf), dtype=torch.int8, device='cuda'))
        ib = torch.tensor(2, dtype=torch.int32, device='cuda')
        y = torch.empty(N * OH * OW, Cout, dtype=torch.float32, device='cuda')
        k_out = torch.zeros(N * OH * OW, Cout, dtype=torch.int8, device='cuda')
        sums = torch.zeros(2 * Cout, dtype=torch.int64, device='cuda')
        rt = D.Runtime(seed=5)
        site = D.QuantSite(rt, 'q', 8, 2).cuda()
        rt.finalize('cuda')
        qs = site.abi(OH * OW * Cout, 'cuda')
        ops = 2 * N * OH * OW * Cout * Kf

        def f32():
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pt, pl, OH, OW, ib, ib, -15, None, y)

        def fused():
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pt, pl, OH, OW, ib, ib, -15, None, None, bnq=(qs, k_out, sums))

        for kern, mask in (('halo', 1), ('im2col', 8)):
            _lib.lib().lbt_conv_set_halo(mask)
            for epi, fn in (('fp32', f32), ('fused', fused)):
                if a.profile:
                    fn()
                    torch.cuda.synchronize()
                    torch.cuda.profiler.start()
                    fn()
                    torch.cuda.synchronize()
                    torch.cuda.profiler.stop()
                    continue
                t = timeit(fn, a.iters)
                rows.append(dict(layer=name, kernel=kern, epilogue=epi, us=t * 1e6, tops=ops / t / 1e12))
                print('%-24s %-7s %-6s %8.1f us  %7.1f TOPS' % (name, kern, epi, t * 1e6, ops / t / 1e12), flush=True)
        _lib.lib().lbt_conv_set_halo(1)
    if a.out:
        json.dump(rows, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()

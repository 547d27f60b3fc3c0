"""Config-3 quantiser microbench: GB/s of lbt_quantize vs the HBM roofline (run on the GPU box).

Algorithmic bytes per element: 4 read + s written (s = 1 for <=8-bit packed mantissas,
2 for <=16-bit, 4 for the fp32 fake-quant output).  L2 is flushed between timed launches.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib, quantizer as Q  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


def time_launch(fn, flush, iters=7, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)                      # 512 MB write: evicts L2 (126 MB)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='20,22,24,26,28')
    ap.add_argument('--bits', default='4,6,8,16')
    ap.add_argument('--tune', action='store_true', help='sweep blocks/SM and rows/tile at 2^26')
    ap.add_argument('--out', default='gpurun_out/quantize_bench.json')
    a = ap.parse_args()
    peak, which = peak_gbs()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    rows = []
    h = _lib.lib()

    def run(log2n, bits, mode, outk):
        n = 1 << log2n
        x = torch.randn(256, n // 256, device='cuda') * 1.3
        ib = torch.tensor(2, dtype=torch.int32, device='cuda')
        cnt = Q.new_counters('cuda')
        noise = torch.rand(n // 256, device='cuda') if mode == Q.ROUND_NOISE else None
        want_fp32 = outk in ('fp32', 'both')
        mk = Q.MANT_NONE if outk == 'fp32' else (Q.MANT_S8 if bits <= 8 else Q.MANT_S16)
        out = torch.empty_like(x) if want_fp32 else None
        om = torch.empty_like(x, dtype=torch.int8 if mk == Q.MANT_S8 else torch.int16) if mk else None
        fn = lambda: Q.quantize(x, bits, ib, mode=mode, noise=noise, seed=1, offset=2, want_fp32=want_fp32,
                                mant_kind=mk, counters=cnt, update_range=True, out=out, out_mant=om)
        t = time_launch(fn, flush)
        s = (4 if want_fp32 else 0) + (0 if not mk else (1 if mk == Q.MANT_S8 else 2))
        gbs = n * (4 + s) / t / 1e9
        return dict(log2n=log2n, bits=bits, mode=['nearest', 'noise', 'philox'][mode], out=outk, us=t * 1e6,
                    bytes_per_elem=4 + s, gbs=gbs, frac=gbs / peak)

    if a.tune:
        for bps in (2, 3, 4, 6, 8, 12, 16):
            for rpg in (4, 8, 16, 32):
                h.lbt_quantize_tune(bps, rpg)
                for outk in ('mant', 'fp32'):
                    r = run(26, 8, Q.ROUND_PHILOX, outk)
                    r.update(blocks_per_sm=bps, rows_per_group=rpg)
                    rows.append(r)
                    print(json.dumps(r), flush=True)
        best = max((r for r in rows if r['out'] == 'mant'), key=lambda r: r['gbs'])
        print('BEST', json.dumps(best))
        h.lbt_quantize_tune(best['blocks_per_sm'], best['rows_per_group'])
    for log2n in [int(s) for s in a.sizes.split(',')]:
        for bits in [int(s) for s in a.bits.split(',')]:
            for mode in (Q.ROUND_NEAREST, Q.ROUND_NOISE, Q.ROUND_PHILOX):
                for outk in ('mant', 'fp32'):
                    r = run(log2n, bits, mode, outk)
                    rows.append(r)
                    print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(dict(peak_gbs=peak, peak_kind=which, rows=rows), open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()

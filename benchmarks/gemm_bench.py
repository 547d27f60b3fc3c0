"""Config-3 GEMM / convolution microbench: TOPS of lbt_gemm_i8 and the implicit-GEMM convolution kernels vs the int8
dense tensor-core peak (run on the GPU box).

    python benchmarks/gemm_bench.py [--square 4096,8192,16384] [--layers resnet20|resnet18|none] [--out FILE]

Ops = 2*M*N*K per launch.  Every timed launch is preceded by an L2 flush when --flush is given; by
default launches run back to back (the square GEMMs are compute bound, operands >> L2 at 16384).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import _lib, dfxp, gemm as G, quantizer as Q  # noqa: E402

INT8_PEAK_TOPS = 4500.0     # B200 dense int8 nominal (B200_PROFILING.md); no measured int8 figure in MEASURED_PEAKS.json


ITERS = [10]


def timeit(fn, iters=None, warmup=3, flush=None):
    iters = iters or ITERS[0]
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        # the launches are captured in a CUDA graph: small kernels are otherwise bound by the ~15 us Python/ctypes call
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / iters
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def rand_i8(*shape, unsigned=False):
    if unsigned:
        return torch.randint(0, 256, shape, dtype=torch.uint8, device='cuda')
    return torch.randint(-128, 128, shape, dtype=torch.int8, device='cuda')


def sustained(fn, seconds=2.0):
    """Run fn back to back for ~`seconds`, sampling SM clock / power / throttle reasons with nvidia-smi meanwhile."""
    import subprocess
    import time
    q = 'clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown'
    mon = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '100'],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    a.record()
    while time.time() - t0 < seconds:
        for _ in range(8):
            fn()
        n += 8
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / n
    time.sleep(0.15)
    mon.terminate()
    rows = [l.split(',') for l in mon.stdout.read().strip().splitlines() if l.count(',') >= 4]
    mid = rows[len(rows) // 3:] or rows
    clk = sorted(float(r[0]) for r in mid)
    return dt, dict(sm_mhz_median=clk[len(clk) // 2] if clk else None, power_w_max=max(float(r[1]) for r in mid) if mid else None,
                    sw_power_cap=any('Active' in r[2] and 'Not' not in r[2] for r in mid),
                    hw_slowdown=any('Active' in r[3] and 'Not' not in r[3] for r in mid),
                    sw_thermal=any('Active' in r[4] and 'Not' not in r[4] for r in mid))


def rand_bits(n, bits, unsigned=False):
    """Mantissas uniform over the full range of a `bits`-wide DFXP quantiser (activations: bits+1 unsigned)."""
    if unsigned:
        return torch.randint(0, 1 << bits, (n, n), dtype=torch.int32, device='cuda').to(torch.uint8)
    return torch.randint(-(1 << (bits - 1)), 1 << (bits - 1), (n, n), dtype=torch.int32, device='cuda').to(torch.int8)


def bench_square16(n, flush):
    """16-bit mantissas on both operands: four 8-bit passes into the int64 accumulator (ops counted once: 2*n^3)."""
    A = torch.randint(-32768, 32768, (n, n), dtype=torch.int32, device='cuda').to(torch.int16)
    B = torch.randint(-32768, 32768, (n, n), dtype=torch.int32, device='cuda').to(torch.int16)
    acc = torch.zeros(n, n, dtype=torch.int64, device='cuda')
    ah, al = G.split_s16(A)
    bh, bl = G.split_s16(B)

    def run():
        for a, b, alpha in ((ah, bh, 65536), (ah, bl, 256), (al, bh, 256), (al, bl, 1)):
            G.gemm_i8_acc64(a, b, acc, alpha=alpha, k_splits=1)
    t = timeit(run, flush=flush)
    return dict(kind='gemm', M=n, N=n, K=n, bits=16, us=t * 1e6, tops=2 * n ** 3 / t / 1e12,
                note='16-bit x 16-bit mantissas = four 8-bit tensor-core passes + int64 atomics epilogue; operand split not timed')


def bench_square(n, flush, bits=8):
    if bits > 9:
        return bench_square16(n, flush)
    A, B = rand_bits(n, bits, unsigned=True), rand_bits(n, bits)
    out = torch.empty(n, n, dtype=torch.float32, device='cuda')
    t = timeit(lambda: G.gemm_i8(A, B, exp_const=-14, out=out), flush=flush)
    row = dict(kind='gemm', M=n, N=n, K=n, bits=bits, us=t * 1e6, tops=2 * n ** 3 / t / 1e12)
    if n >= 8192:
        ts, clocks = sustained(lambda: G.gemm_i8(A, B, exp_const=-14, out=out))
        row.update(sustained_tops=2 * n ** 3 / ts / 1e12, clocks=clocks)
        print('   sustained %.0f TOPS, clocks %s' % (row['sustained_tops'], clocks))
    return row


def conv_case(N, H, W, Cin, Cout, k, s, flush, name):
    """fprop / dgrad (stride 1) / wgrad of one Conv2d_q shape through the C ABI."""
    OH, pt, _ = dfxp.same_pad(H, k, s)
    OW, pl, _ = dfxp.same_pad(W, k, s)
    x = rand_i8(N, H, W, Cin, unsigned=True)
    g = rand_i8(N, OH, OW, Cout)
    Kf = k * k * Cin
    wt = torch.zeros(Cout, dfxp._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
    wt.copy_(rand_i8(Cout, Kf))
    ib = torch.tensor(2, dtype=torch.int32, device='cuda')
    y = torch.empty(N * OH * OW, Cout, dtype=torch.float32, device='cuda')
    rows = []
    ops = 2 * N * OH * OW * Cout * Kf
    if dfxp._implicit_ok(Cin, k, k):
        t = timeit(lambda: dfxp._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, s, s, pt, pl, OH, OW, ib, ib, -15, None, y),
                   flush=flush)
        byts = x.numel() + y.numel() * 4
        rows.append(dict(kind='fprop', layer=name, M=N * OH * OW, N=Cout, K=Kf, us=t * 1e6, tops=ops / t / 1e12,
                         gbs=byts / t / 1e9))
    if dfxp._implicit_ok(Cin, k, k) and dfxp._implicit_ok(Cout, 1, 1):
        acc = torch.zeros(Kf, Cout, dtype=torch.int64, device='cuda')
        t = timeit(lambda: _lib.call('lbt_conv_i8_wgrad', _lib.ptr(x), Q.MANT_U8, N, H, W, Cin, _lib.ptr(g), Q.MANT_S8, Cout,
                                     k, k, s, s, pt, pl, OH, OW, _lib.ptr(acc), 1, 0, _lib.stream()), flush=flush)
        rows.append(dict(kind='wgrad', layer=name, M=Kf, N=Cout, K=N * OH * OW, us=t * 1e6, tops=ops / t / 1e12,
                         gbs=(x.numel() + g.numel()) / t / 1e9))
    return rows


LAYERS = {
    'resnet20': [(256, 32, 32, 16, 16, 3, 1, 's1 3x3 16->16'), (256, 32, 32, 16, 32, 3, 2, 's2 3x3 16->32 /2'),
                 (256, 16, 16, 32, 32, 3, 1, 's2 3x3 32->32'), (256, 16, 16, 32, 64, 3, 2, 's3 3x3 32->64 /2'),
                 (256, 8, 8, 64, 64, 3, 1, 's3 3x3 64->64')],
    'resnet18': [(256, 56, 56, 64, 64, 3, 1, 'l1 3x3 64->64'), (256, 56, 56, 64, 128, 3, 2, 'l2 3x3 64->128 /2'),
                 (256, 28, 28, 128, 128, 3, 1, 'l2 3x3 128->128'), (256, 28, 28, 128, 256, 3, 2, 'l3 3x3 128->256 /2'),
                 (256, 14, 14, 256, 256, 3, 1, 'l3 3x3 256->256'), (256, 14, 14, 256, 512, 3, 2, 'l4 3x3 256->512 /2'),
                 (256, 7, 7, 512, 512, 3, 1, 'l4 3x3 512->512')],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--square', default='4096,8192,16384')
    ap.add_argument('--layers', default='resnet20,resnet18')
    ap.add_argument('--flush', action='store_true')
    ap.add_argument('--bits', default='8', help='operand widths of the square GEMM sweep (config 3: 4,6,8)')
    ap.add_argument('--only', default='', help='substring filter on the layer name')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--path', type=int, default=1, help='0: TMA kernels only, 1: gather kernels for narrow channels')
    ap.add_argument('--pair', type=int, default=1, help='0: single-CTA GEMM kernel only, 1: CTA-pair kernel for large GEMMs')
    ap.add_argument('--out', default='gpurun_out/gemm_bench.json')
    a = ap.parse_args()
    ITERS[0] = a.iters
    _lib.lib().lbt_conv_set_path(a.path)
    _lib.lib().lbt_gemm_set_pair(a.pair)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda') if a.flush else None
    rows = []
    for n in [int(s) for s in a.square.split(',') if s]:
        for bits in [int(b) for b in a.bits.split(',')]:
            rows.append(bench_square(n, flush, bits))
    for fam in [f for f in a.layers.split(',') if f and f != 'none']:
        for (N, H, W, Ci, Co, k, s, name) in LAYERS[fam]:
            if a.only and a.only not in name:
                continue
            rows += conv_case(N, H, W, Ci, Co, k, s, flush, fam + ' ' + name)
    for r in rows:
        r['frac_int8_peak'] = r['tops'] / INT8_PEAK_TOPS
        print('%-6s %-32s M=%-8d N=%-6d K=%-8d %9.1f us %8.1f TOPS (%.3f of %.0f)%s' % (
            r['kind'], r.get('layer', 'square %d-bit' % r.get('bits', 8)), r['M'], r['N'], r['K'], r['us'], r['tops'], r['frac_int8_peak'],
            INT8_PEAK_TOPS, '  %.0f GB/s' % r['gbs'] if 'gbs' in r else ''))
    assert G.debug_error() == 0
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(dict(int8_peak_tops=INT8_PEAK_TOPS, flush=a.flush, rows=rows), open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()

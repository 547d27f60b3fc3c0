"""The measured int8 tensor-core peak of this box ("to be measured"): stock torch._int_mm (cuBLASLt s8 x s8 -> s32)
at 8192^3, best-of-10 burst and 4 s sustained — the method of MEASURED_PEAKS.json's bf16 entry — with lbt_gemm_i8 timed the
same way beside it.

    python benchmarks/int8_peak.py [--n 8192] [--out gpurun_out/int8_peak.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import gemm as G  # noqa: E402


def burst(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3
        best = t if best is None else min(best, t)
    return best


def sustained(fn, seconds=4.0):
    q = 'clocks.sm,power.draw'
    mon = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '200'],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    a.record()
    while time.time() - t0 < seconds:
        for _ in range(16):
            fn()
        n += 16
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    mon.terminate()
    rows = [l.split(',') for l in mon.stdout.read().strip().splitlines() if l.count(',') >= 1]
    tail = rows[len(rows) // 2:] or rows
    clk = sorted(float(r[0]) for r in tail) if tail else [0.0]
    return a.elapsed_time(b) * 1e-3 / n, clk[len(clk) // 2], max(float(r[1]) for r in tail) if tail else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=8192)
    ap.add_argument('--out', default='gpurun_out/int8_peak.json')
    a = ap.parse_args()
    n = a.n
    ops = 2.0 * n ** 3
    A = torch.randint(-128, 128, (n, n), dtype=torch.int8, device='cuda')
    Bt = torch.randint(-128, 128, (n, n), dtype=torch.int8, device='cuda')        # [N, K]: B^T, K-major like lbt_gemm_i8's operand
    Brow = Bt.t().contiguous()                                                    # [K, N] row-major
    out32 = torch.empty(n, n, dtype=torch.float32, device='cuda')
    cases = {
        'torch._int_mm (cuBLASLt s8*s8->s32), B column-major': lambda: torch._int_mm(A, Bt.t()),
        'torch._int_mm (cuBLASLt s8*s8->s32), B row-major': lambda: torch._int_mm(A, Brow),
        'lbt_gemm_i8 (s8*s8->s32 in TMEM, fp32 epilogue)': lambda: G.gemm_i8(A, Bt, exp_const=-14, out=out32),
    }
    res = {}
    for name, fn in cases.items():
        try:
            tb = burst(fn)
            ts, clk, pw = sustained(fn)
            res[name] = dict(burst_us=tb * 1e6, burst_tops=ops / tb / 1e12, sustained_us=ts * 1e6, sustained_tops=ops / ts / 1e12,
                             sm_mhz_sustained=clk, power_w_max=pw)
        except Exception as e:  # noqa: BLE001
            res[name] = dict(error='%s: %s' % (type(e).__name__, e))
        print(name, res[name], flush=True)
        time.sleep(2.0)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(dict(n=n, operands='uniform s8 over the full range', results=res), open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()

import sys, torch
sys.path.insert(0, '.')
from lbt_b200 import quantizer as Q
n = 1 << 28
xs = [torch.randn(256, n // 256, device='cuda') for _ in range(2)]
ib = torch.tensor(2, dtype=torch.int32, device='cuda')
cnt = Q.new_counters('cuda')
def t(fn, reps=5):
    fn(0); fn(1); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(0); fn(1); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 2)
    return best
ro = t(lambda i: Q.quantize(xs[i], 8, ib, mode=Q.ROUND_NEAREST, want_fp32=False, mant_kind=Q.MANT_NONE, counters=cnt, update_range=False))
print('read-only (statistics only): %.1f us, %.0f GB/s' % (ro * 1e3, n * 4 / ro / 1e6))
s = t(lambda i: xs[i].sum())
print('torch sum: %.1f us, %.0f GB/s' % (s * 1e3, n * 4 / s / 1e6))
om = [torch.empty(256, n // 256, dtype=torch.int8, device='cuda') for _ in range(2)]
m = t(lambda i: Q.quantize(xs[i], 8, ib, mode=Q.ROUND_NEAREST, want_fp32=False, mant_kind=Q.MANT_S8, counters=cnt, update_range=False, out_mant=om[i]))
print('s8 out: %.1f us, %.0f GB/s' % (m * 1e3, n * 5 / m / 1e6))
mm = t(lambda i: Q.quantize(xs[i], 8, ib, mode=Q.ROUND_NEAREST | Q.STATS_MINMAX, want_fp32=False, mant_kind=Q.MANT_S8, counters=cnt, update_range=False, out_mant=om[i]))
print('s8 out, minmax stats: %.1f us, %.0f GB/s' % (mm * 1e3, n * 5 / mm / 1e6))

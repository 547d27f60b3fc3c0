"""First-light check of the tcgen05 GEMM on a B200: tiny problems, prints mismatches instead of asserting."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbt_b200 import gemm as G  # noqa: E402


def run(M, N, K, sa=False, sb=True):
    rng = np.random.default_rng(0)
    ld = -(-K // 16) * 16
    a = torch.zeros(M, ld, dtype=torch.int8 if sa else torch.uint8)
    b = torch.zeros(N, ld, dtype=torch.int8 if sb else torch.uint8)
    a[:, :K] = torch.from_numpy(rng.integers(-128 if sa else 0, 128 if sa else 256, (M, K)).astype(np.int8 if sa else np.uint8))
    b[:, :K] = torch.from_numpy(rng.integers(-128 if sb else 0, 128 if sb else 256, (N, K)).astype(np.int8 if sb else np.uint8))
    A, B = a.cuda()[:, :K], b.cuda()[:, :K]
    out = G.gemm_i8(A, B)
    torch.cuda.synchronize()
    err = G.debug_error()
    ref = (A.double() @ B.double().T).float()
    bad = (out != ref)
    print('M=%d N=%d K=%d sa=%d sb=%d  watchdog=%d  mismatches=%d/%d  maxdiff=%g' % (
        M, N, K, sa, sb, err, int(bad.sum()), bad.numel(), float((out - ref).abs().max())), flush=True)
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        for i, j in idx:
            print('   [%d,%d] got %g want %g' % (i, j, float(out[i, j]), float(ref[i, j])))
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print('   bad rows %d..%d (%d)  bad cols %d..%d (%d)' % (int(rows.min()), int(rows.max()), len(rows),
                                                               int(cols.min()), int(cols.max()), len(cols)))
    return int(bad.sum()) == 0 and err == 0


if __name__ == '__main__':
    ok = True
    for shp in [(128, 16, 32), (128, 16, 128), (128, 16, 256), (128, 128, 512), (256, 256, 1024), (300, 40, 200),
                (4096, 256, 4096)]:
        ok &= run(*shp)
        ok &= run(*shp, sa=True, sb=True)
    print('ALL OK' if ok else 'FAILURES')

"""DRAM traffic per launch of each kernel class of a bench.py step, from an `ncu --set full` report of that very command:

    python benchmarks/ncu_traffic.py <workload> gpurun_out/prof.ncu-rep profiles/r2_<name>_ncu_full_summary.csv

writes the per-launch summary CSV (benchmarks/ncu_summary.py's columns) and merges
{workload: {C-ABI entry: {bytes_per_launch, launches, source}}} into profiles/traffic.json, which bench.py reads at run time
for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of the class).
Runs where ncu is installed; no GPU needed.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'benchmarks'))
import ncu_summary  # noqa: E402

# kernel-name fragment -> C-ABI entry whose launches bench.py groups under that name
CLASSES = [
    ('conv_wgrad', 'lbt_conv_i8_wgrad'), ('conv_ldg_wgrad', 'lbt_conv_i8_wgrad'), ('conv_fprop', 'lbt_conv_i8_fprop'),
    ('conv_ldg_kernel', 'lbt_conv_i8_fprop'), ('conv_stem', 'lbt_conv_i8_fprop'), ('gemm_i8', 'lbt_gemm_i8'),
    ('bn_fwd2', 'lbt_bn_fwd_apply'), ('bn_fwd1', 'lbt_bn_fwd_quant_stats'), ('bn_bwd1', 'lbt_bn_bwd_quant_stats'),
    ('bn_bwd2', 'lbt_bn_bwd_apply'), ('quantize_', 'lbt_quantize'), ('maxpool_fwd', 'lbt_maxpool_fwd'),
    ('maxpool_bwd', 'lbt_maxpool_bwd'),
]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main(workload, rep, out_csv):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    kn, rd, wr = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    acc = {}
    for r in rows[2:]:
        for frag, entry in CLASSES:
            if frag in r[kn]:
                d = acc.setdefault(entry, [0, 0.0])
                d[0] += 1
                d[1] += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
                break
    with open(out_csv, 'w') as f:
        stdout, sys.stdout = sys.stdout, f
        try:
            ncu_summary.main(rep)
        finally:
            sys.stdout = stdout
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[workload] = {e: {'bytes_per_launch': b / n, 'launches': n, 'source': os.path.relpath(out_csv, ROOT)} for e, (n, b) in acc.items()}
    with open(path, 'w') as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(table[workload], indent=1))


if __name__ == '__main__':
    main(*sys.argv[1:4])

"""Compact per-launch summary of an `ncu --set full` report (run where ncu is installed; no GPU needed):

    python benchmarks/ncu_summary.py gpurun_out/foo.ncu-rep > profiles/r1_foo_ncu_summary.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__waves_per_multiprocessor', 'waves'),
    ('smsp__inst_executed.sum', 'warp_insts'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    kn = h.index('Kernel Name')
    cols = [(h.index(m), n) for m, n in METRICS if m in h]
    out = csv.writer(sys.stdout)
    out.writerow(['kernel'] + ['%s [%s]' % (n, units[i]) if units[i] else n for i, n in cols])
    for r in rows[2:]:
        name = r[kn].split('(')[0].replace('lbt::<unnamed>::', '').replace('void ', '').replace('unnamed>::', '')
        out.writerow([name] + [r[i] for i, _ in cols])


if __name__ == '__main__':
    main(sys.argv[1])

"""Turn the CSV pages the evidence run (benchmarks/gpu_evidence.sh) leaves in gpurun_out/ into the committed evidence under profiles/:

    python benchmarks/collect_profiles.py <tag> [workload]

  * profiles/<tag>_ncu_<group>_summary.csv  — per-launch summary of every `ncu --set full` capture (time, DRAM bytes, pipe and
    issue utilisation, registers, grid) from gpurun_out/<tag>_ncu_<group>_raw.csv;
  * profiles/<tag>_launches_<workload>.csv  — the ncu launch list of one step (gpu__time_duration.sum) as a per-kernel table;
  * profiles/traffic.json                    — measured dram__bytes_read.sum + dram__bytes_write.sum per launch of each C-ABI entry
    (averaged over the captured launches), which bench.py reads at run time for `roofline.traffic`.
No GPU needed."""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'benchmarks'))
import summarize_launches  # noqa: E402
from ncu_summary import METRICS  # noqa: E402

# kernel-name fragment -> C-ABI entry whose launches bench.py groups under that name
CLASSES = [
    ('stem_wgrad', 'lbt_conv_i8_wgrad_c3'), ('conv_wgrad', 'lbt_conv_i8_wgrad'), ('conv_ldg_wgrad', 'lbt_conv_i8_wgrad'),
    ('conv_fprop', 'lbt_conv_i8_fprop'), ('conv_ldg_kernel', 'lbt_conv_i8_fprop'), ('conv_halo', 'lbt_conv_i8_fprop'),
    ('gemm_i8', 'lbt_gemm_i8'), ('bn_fwd2_pool', 'lbt_bn_fwd_apply_pooled'), ('bn_fwd2_kernel<(bool)1, (bool)1>', 'lbt_bn_fwd_apply2'),
    ('bn_fwd2_kernel<(bool)0, (bool)1>', 'lbt_bn_fwd_apply2'), ('bn_fwd2', 'lbt_bn_fwd_apply'), ('bn_fwd1', 'lbt_bn_fwd_quant_stats'),
    ('bn_bwd1', 'lbt_bn_bwd_quant_stats'), ('bn_bwd2', 'lbt_bn_bwd_apply'), ('quantize_', 'lbt_quantize'),
    ('maxpool_fwd', 'lbt_maxpool_fwd'), ('maxpool_bwd', 'lbt_maxpool_bwd'), ('param_prep', 'lbt_param_prep'),
    ('finalize_multi', 'lbt_finalize_multi'), ('dp_step', 'lbt_dp_step'),
]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main(tag, workload='resnet18'):
    acc = {}
    for raw in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', tag + '_ncu_*_raw.csv'))):
        rows = list(csv.reader(open(raw)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        kn, rd, wr = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
        cols = [(h.index(m), n) for m, n in METRICS if m in h]
        group = os.path.basename(raw)[len(tag) + 5:-8]
        out = os.path.join(ROOT, 'profiles', '%s_ncu_%s_summary.csv' % (tag, group))
        with open(out, 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel'] + ['%s [%s]' % (n, units[i]) if units[i] else n for i, n in cols])
            for r in rows[2:]:
                name = r[kn].split('(')[0].replace('lbt::<unnamed>::', '').replace('void ', '').replace('unnamed>::', '')
                w.writerow([name] + [r[i] for i, _ in cols])
                for frag, entry in CLASSES:
                    if frag in r[kn]:
                        d = acc.setdefault(entry, [0, 0.0, os.path.relpath(out, ROOT)])
                        d[0] += 1
                        d[1] += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
                        break
        print('wrote', out)
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    if acc:
        table[workload] = {e: {'bytes_per_launch': b / n, 'launches': n, 'source': src} for e, (n, b, src) in sorted(acc.items())}
        with open(path, 'w') as f:
            json.dump(table, f, indent=1, sort_keys=True)
        print('wrote', path)
    ll = os.path.join(ROOT, 'gpurun_out', '%s_launches_%s.csv' % (tag, workload))
    if os.path.exists(ll):
        out = os.path.join(ROOT, 'profiles', '%s_launches_%s.csv' % (tag, workload))
        with open(out, 'w') as f:
            f.write('# ncu --metrics gpu__time_duration.sum --clock-control none over ONE step of `python bench.py --steps 1` (%s); '
                    'cold-cache, serialised launches: shares, not absolute times\n' % workload)
            f.write(summarize_launches.summarize(ll) + '\n')
        print('wrote', out)


if __name__ == '__main__':
    main(*sys.argv[1:3])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import collections
import csv
import sys


def summarize(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        v = v / 1000 if r[mu] == 'ns' else (v * 1000 if r[mu] == 'ms' else v)
        name = r[kn].split('(')[0].replace('lbt::<unnamed>::', '').replace('void ', '')[:64]
        d[name][0] += 1
        d[name][1] += v
    tot = sum(v[1] for v in d.values())
    out = ['# total %.1f us over %d launches' % (tot, sum(v[0] for v in d.values())), 'kernel,launches,total_us,share,avg_us']
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append('%s,%d,%.1f,%.3f,%.1f' % (k, v[0], v[1], v[1] / tot, v[1] / v[0]))
    return '\n'.join(out)


if __name__ == '__main__':
    print(summarize(sys.argv[1]))

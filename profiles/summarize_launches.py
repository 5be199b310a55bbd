"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> markdown table.

    python profiles/summarize_launches.py profiles/<launches>.csv > profiles/<launches>.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        name = d['Kernel Name'].split('(')[0].replace('void ', '')
        t = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        t = t / 1e3 if unit == 'ns' else (t * 1e3 if unit == 'ms' else t)
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f%% |' % (k[:70], v[0], v[1], 100 * v[1] / tot))
    print('\ntotal: %.1f us over %d launches' % (tot, len(data)))


if __name__ == '__main__':
    main(sys.argv[1])

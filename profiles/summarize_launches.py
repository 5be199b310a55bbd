"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> markdown table.

    python profiles/summarize_launches.py profiles/<launches>.csv [--step K] > profiles/<launches>.md
--step K: only the launches after the K-th `adam_kernel` launch up to and including the next one (= one training step).
The epilogue variants of `gemm_tc_kernel<bf16, pair, epilogue>` are listed separately and as one total.
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import collections
import csv
import sys


def kernel_name(full):
    """the name with its template arguments, without the parameter list (the first '(' outside <...>)"""
    depth = 0
    for i, c in enumerate(full):
        if c == '<':
            depth += 1
        elif c == '>':
            depth -= 1
        elif c == '(' and depth == 0:
            return full[:i].replace('void ', '').replace('(int)', '')
    return full.replace('void ', '')


def main(path, step=None):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    if step is not None:
        adam = [i for i, d in enumerate(data) if 'adam_kernel' in d['Kernel Name']]
        data = data[adam[step] + 1:adam[step + 1] + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    gemm = [0, 0.0]
    for d in data:
        name = kernel_name(d['Kernel Name'])
        t = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        t = t / 1e3 if unit == 'ns' else (t * 1e3 if unit == 'ms' else t)
        agg[name][0] += 1
        agg[name][1] += t
        if name.startswith('gemm_tc_kernel'):
            gemm[0] += 1
            gemm[1] += t
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f%% |' % (k[:70], v[0], v[1], 100 * v[1] / tot))
    print('\ntotal: %.1f us over %d launches' % (tot, len(data)))
    if gemm[0]:
        print('\n`gemm_tc_kernel`, all variants: %d launches, %.1f us, %.1f%% of the serialised step' % (gemm[0], gemm[1], 100 * gemm[1] / tot))


if __name__ == '__main__':
    a = sys.argv[1:]
    st = int(a[a.index('--step') + 1]) if '--step' in a else None
    main(a[0], st)

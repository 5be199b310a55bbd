"""Pick the metrics B200_PROFILING.md names out of an `ncu -i X.ncu-rep --page raw --csv` dump -> markdown.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/prof.csv
    python profiles/summarize_ncu_raw.py /tmp/prof.csv > profiles/<name>.md
"""
import csv
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
]


def main(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(r for r in rows if r[0] == 'ID')
    units = rows[rows.index(hdr) + 1]
    for r in rows[rows.index(hdr) + 2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('### `%s`  grid %s block %s\n' % (d.get('Kernel Name', '?')[:90], d.get('Grid Size'), d.get('Block Size')))
        print('| metric | value | unit |\n|---|---:|---|')
        for k in KEYS:
            if k in d and d[k] != '':
                print('| %s | %s | %s |' % (k, d[k], u.get(k, '')))
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                         if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and v not in ('', 'n/a')),
                        reverse=True)[:6]
        for v, k in stalls:
            print('| %s | %.2f | warps/issue |' % (k.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''), v))
        print()


if __name__ == '__main__':
    main(sys.argv[1])

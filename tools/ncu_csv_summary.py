#!/usr/bin/env python
"""ncu raw page (ncu -i X.ncu-rep --page raw --csv) -> per-launch summary JSON, launches in capture order.
usage: tools/ncu_csv_summary.py raw.csv out.json [label ...]   (labels: path names, attached by tools/prof_paths.py order)"""
import csv
import json
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 'msecond': 1e-3,
        'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}

rows = list(csv.reader(open(sys.argv[1], errors='replace')))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, units = rows[start], rows[start + 1]
out = []
for r in rows[start + 2:]:
    if len(r) != len(hdr):
        continue
    d = {'kernel': r[hdr.index('Kernel Name')].split('(')[0]}
    for k in KEYS:
        if k in hdr:
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            try:
                v = float(v.replace(',', ''))
            except ValueError:
                continue
            if u in UNIT and ('bytes' in k or 'duration' in k):
                v *= UNIT[u]
                u = 'byte' if 'bytes' in k else 's'
            d[k] = v
            if u:
                d[k + '.unit'] = u
    d['dram_bytes_total'] = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    if d.get('gpu__time_duration.sum'):
        d['dram_gbs'] = d['dram_bytes_total'] / d['gpu__time_duration.sum'] / 1e9
        d['us'] = d['gpu__time_duration.sum'] * 1e6
    out.append(d)
json.dump({'source': sys.argv[1], 'note': 'ncu --set full --clock-control none, one launch per row, capture order; '
           'times under ncu are cold-cache and serialised', 'launches': out}, open(sys.argv[2], 'w'), indent=1)
for d in out:
    print('%-44s %9.1f us  dram %8.1f MB r %8.1f MB w  %6.0f GB/s  issue %5.1f%%  warps %5.1f%%  inst %8.2f M  regs %3d' % (
        d['kernel'][:44], d.get('us', 0), d.get('dram__bytes_read.sum', 0) / 1e6, d.get('dram__bytes_write.sum', 0) / 1e6,
        d.get('dram_gbs', 0), d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0),
        d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0), d.get('smsp__inst_executed.sum', 0) / 1e6,
        int(d.get('launch__registers_per_thread', 0))))

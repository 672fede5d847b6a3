#!/usr/bin/env python
"""Turn ncu reports in gpurun_out/ into the committed summaries under profiles/.
usage: tools/profile_summary.py <tag>   (reads gpurun_out/prof_scan.ncu-rep, prof_emit.ncu-rep, launches.csv)"""
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'sm__inst_executed.sum',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 'msecond': 1e-3,
        'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}
out = {}
for name in ('scan', 'emit'):
    rep = 'gpurun_out/prof_%s.ncu-rep' % name
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ks = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    v = float(v.replace(',', ''))
                except ValueError:
                    continue
                d[k] = v
                if u:
                    d[k + '.unit'] = u
        b = 0.0
        for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            if k in d:
                b += d[k] * UNIT.get(d.get(k + '.unit', 'byte'), 1.0)
        d['dram_bytes_total'] = b
        ks.append(d)
    out[name] = ks
if os.path.exists('gpurun_out/launches.csv'):
    rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 5]
    hdr = next((r for r in rows if 'Kernel Name' in r), None)
    if hdr:
        agg = {}
        for r in rows:
            if r is hdr or len(r) != len(hdr):
                continue
            if r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
                continue
            k = r[hdr.index('Kernel Name')].split('(')[0][:60]
            v = float(r[hdr.index('Metric Value')].replace(',', ''))
            u = r[hdr.index('Metric Unit')]
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1.0)
            a = agg.setdefault(k, [0, 0.0, []])
            a[0] += 1
            a[1] += v
            a[2].append(v)
        tot = sum(a[1] for a in agg.values())
        out['launch_list'] = [{'kernel': k, 'launches': a[0], 'total_us': round(a[1], 2), 'avg_us': round(a[1] / a[0], 2),
                               'share': round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda x: -x[1][1])]
        # the capture window holds launches of the timed steps (1 GiB per launch) followed by the 64 MiB chunks of the
        # end-to-end leg: the steps are the launches of at least half the longest duration of their kernel
        steps = {k: [v for v in a[2] if v >= 0.5 * max(a[2])] for k, a in agg.items()}
        stot = sum(sum(v) for v in steps.values())
        out['launch_list_timed_steps'] = [{'kernel': k, 'launches': len(v), 'avg_us': round(sum(v) / len(v), 2),
                                           'share': round(sum(v) / stot, 4)}
                                          for k, v in sorted(steps.items(), key=lambda x: -sum(x[1]))]
os.makedirs('profiles', exist_ok=True)
json.dump(out, open('profiles/%s_ncu_summary.json' % tag, 'w'), indent=1)
print(json.dumps(out, indent=1)[:3000])

#!/bin/bash
# round 2, call E: baseline evidence of the restored tree -- parity tests, default bench + reference arm, launch list
# and full ncu metrics of every device path (tools/prof_paths.py)
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
S=$(date +%s)
timeout -s KILL 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
tail -c 1200 gpurun_out/bench.err
S=$(date +%s)
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $? in $(( $(date +%s) - S )) s"
cut -c1-300 gpurun_out/bench_ref.log
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'pipe', round(d['roofline']['pipeline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    print('cpu', json.dumps(d.get('cpu_baseline'))[:1200])
    print('dropin', json.dumps(d['e2e'].get('dropin_per_record'))[:400])
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:700])
except Exception as e:
    print('bench parse failed', e)
PY
S=$(date +%s)
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_paths.csv python tools/prof_paths.py > gpurun_out/prof_paths_launch.log 2>&1; echo "launch list exit $? in $(( $(date +%s) - S )) s"
grep -E "^ran|records" gpurun_out/prof_paths_launch.log | tr '\n' ';'; echo
S=$(date +%s)
timeout -s KILL 1500 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/prof_paths python tools/prof_paths.py > gpurun_out/prof_paths_full.log 2>&1; echo "ncu full exit $? in $(( $(date +%s) - S )) s"
ncu -i gpurun_out/prof_paths.ncu-rep --page raw --csv > gpurun_out/prof_paths_raw.csv 2> /dev/null
ls -la gpurun_out/ | head -30
python tools/ncu_csv_summary.py gpurun_out/prof_paths_raw.csv gpurun_out/prof_paths_summary.json | tail -70
# keep the merge-back small: the raw CSV and the summary carry what the report holds
[ $(stat -c %s gpurun_out/prof_paths.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ] && rm -f gpurun_out/prof_paths.ncu-rep
true

#!/bin/bash
# round 2, call M: scan regression hunt (library variants), selective Phred mirror diagnostics + ncu
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "selective or fused_decode or arrayadd" > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
python tools/ab_paths.py fast dec multiline 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for v in variants/*.so; do
  FQB200_LIB=$PWD/$v python tools/ab_paths.py fast dec 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
done
python tools/ab_paths.py fast 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none -k regex:fq_scan_kernel -o gpurun_out/prof_dec python tools/prof_paths.py dec > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec exit $?"
ncu -i gpurun_out/prof_dec.ncu-rep --page raw --csv > gpurun_out/prof_dec_raw.csv 2>/dev/null
python tools/ncu_csv_summary.py gpurun_out/prof_dec_raw.csv gpurun_out/prof_dec_summary.json | tail -3
rm -f gpurun_out/prof_dec.ncu-rep

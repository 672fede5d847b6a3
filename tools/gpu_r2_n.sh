#!/bin/bash
# round 2, call N: parity + fuzz, A/B step times incl. Phred mirror (selective / whole), ncu of the mirror-writing scan
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
python tools/ab_paths.py ${AB_PATHS:-fast dec ont multiline} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for v in ${VARIANTS}; do
  FQB200_LIB=$PWD/$v python tools/ab_paths.py ${AB_PATHS:-fast dec} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
done
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none -k regex:fq_scan_kernel -o gpurun_out/prof_dec python tools/prof_paths.py dec > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec exit $?"
ncu -i gpurun_out/prof_dec.ncu-rep --page raw --csv > gpurun_out/prof_dec_raw.csv 2>/dev/null
python tools/ncu_csv_summary.py gpurun_out/prof_dec_raw.csv gpurun_out/prof_dec_summary.json | tail -3
rm -f gpurun_out/prof_dec.ncu-rep

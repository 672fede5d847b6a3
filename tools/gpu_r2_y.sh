#!/bin/bash
# round 2, call Y: warp-per-chunk speculative pass -- parity tests, A/B, ncu (metrics + SASS source page) of fq_gspec2_kernel
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log | cut -c1-300
for v in ${VARIANTS:-2}; do
  FQB200_SPEC=$v timeout -s KILL 300 python tools/ab_paths.py ${AB_PATHS:-multiline} 2>&1 | grep -v Warning | sed "s/^/spec v$v: /" | tee -a gpurun_out/ab_y.log
done
if [ -n "$FUZZ_SECONDS" ]; then timeout -s KILL 400 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log | cut -c1-600; fi
if [ -z "$NO_NCU" ]; then
k=fq_gspec2_kernel; pth=multiline_spec
timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/src_$k python tools/prof_paths.py $pth > gpurun_out/ncu_src_$k.log 2>&1; echo "ncu $k exit $?"
ncu -i gpurun_out/src_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
ncu -i gpurun_out/src_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
rm -f gpurun_out/src_$k.ncu-rep
python tools/ncu_csv_summary.py gpurun_out/raw_$k.csv gpurun_out/sum_$k.json
fi

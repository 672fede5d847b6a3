#!/bin/bash
# round 2, call G: parity + fuzz, A/B step times, scan configurations, source-level ncu of scan / emit / gspec
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for v in ${VARIANTS}; do
  FQB200_LIB=$PWD/$v python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
done
if [ -n "$CFGS" ]; then python tools/scan_skeleton.py $CFGS 2>&1 | grep -v Warning | tee gpurun_out/cfgs.log; fi
for spec in ${NCU_SRC}; do   # kernelregex:path
  k=${spec%%:*}; pth=${spec##*:}
  timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/src_${k}_${pth} python tools/prof_paths.py $pth > gpurun_out/ncu_src_${k}.log 2>&1; echo "ncu $k $pth exit $?"
  ncu -i gpurun_out/src_${k}_${pth}.ncu-rep --page source --csv > gpurun_out/src_${k}_${pth}.csv 2>/dev/null
  ncu -i gpurun_out/src_${k}_${pth}.ncu-rep --page raw --csv > gpurun_out/raw_${k}_${pth}.csv 2>/dev/null
done
ls -la gpurun_out | head -40

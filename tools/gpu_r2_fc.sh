#!/bin/bash
# round 2: parity tests of the build, FASTA with scan geometry 0 / 2
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log | cut -c1-200
for rep in 1 2; do for c in 0 2; do
  FASTA_CFG=$c timeout -s KILL 300 python tools/ab_paths.py fasta 2>&1 | grep -v Warning | tee -a gpurun_out/ab_fc.log
done; done

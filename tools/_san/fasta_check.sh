timeout -s KILL 300 python -m pytest tests/test_fasta.py -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 600 python bench.py --steps 50 --no-cpu --e2e-steps 1 2>/dev/null > gpurun_out/bench_fa.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fa.log').read().strip().splitlines()[-1]); print(round(d["value"],1)); print(d["extras"]["fasta_1g"])
PY

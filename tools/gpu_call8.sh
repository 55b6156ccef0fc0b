#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c8_pytest.txt
tail -12 gpurun_out/r2c8_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2c8_bench_c4.json 2> gpurun_out/r2c8_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c8_bench_c4.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY

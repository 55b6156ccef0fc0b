#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c7_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c7_pytest.txt
tail -15 gpurun_out/r2c7_pytest.txt
ASGART_B200_DEBUG_PHASES=1 timeout 600 python tools/quick_bench.py 4 0 2 > gpurun_out/r2c7_phases_c4.out 2> gpurun_out/r2c7_phases_c4.err; tail -12 gpurun_out/r2c7_phases_c4.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest > gpurun_out/r2c7_bench_c4.json 2> gpurun_out/r2c7_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c7_bench_c4.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'], d['roofline']['frac'])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c15_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c15_pytest.txt
ASGART_B200_DEBUG_PHASES=1 timeout 600 python tools/quick_bench.py 3 0 3 > gpurun_out/r2c15_phases_c3.out 2> gpurun_out/r2c15_phases_c3.err; grep "sa_build" gpurun_out/r2c15_phases_c3.err | tail -7
timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2c15_bench_c4.json 2> gpurun_out/r2c15_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c15_bench_c4.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY

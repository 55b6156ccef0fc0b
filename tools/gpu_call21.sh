#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2c21_bench_c4.json 2> gpurun_out/r2c21_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2c21_bench_c4.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step']); print({k:(round(v['ms_per_step'],2), round(v['frac_of_hbm_peak'],3)) for k,v in d['kernel_families'].items()})
PY

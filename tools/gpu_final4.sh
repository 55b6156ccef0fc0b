#!/bin/bash
# single-GPU validation after the one-pass deep-table fill: gpu tests, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2i_bench_c4.json',):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), d.get('families_match_oracle_golden'), d['phases_ms_per_step'], d['gpu_launches'])
PY

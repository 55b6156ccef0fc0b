#!/bin/bash
# single-GPU validation after the 21-symbol keys and the byte-permute packing: gpu tests, bench, ncu launch list of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2h_bench_c4.json',):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), d.get('families_match_oracle_golden'), d['phases_ms_per_step'], d['initial_sort'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()}, (d.get('roofline') or {}).get('kernel','')[:30], (d.get('roofline') or {}).get('frac'), d.get('ingest'))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2h_launches_c4.csv python tools/quick_bench.py 4 0 2 > gpurun_out/r2h_launches_c4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2h_launches_c4.log | cut -c1-400

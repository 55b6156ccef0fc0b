#!/bin/bash
# final single-GPU validation of the round: gpu tests, memcheck with the MSD sort forced (lazy ranks, single and sharded), bench,
# and the A/B of the initial key length (ASGART_B200_P0=formula = the LSD rule, 18 symbols at C4)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.txt
ASGART_B200_MSD_MIN=0 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2g_sanitizer_memcheck_msd.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2g_sanitizer_memcheck_msd.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; echo "bench rc=$?"
ASGART_B200_P0=formula timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench_c4_p0formula.json 2> gpurun_out/r2g_bench_c4_p0formula.err; echo "bench(formula) rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2g_bench_c4.json','gpurun_out/r2g_bench_c4_p0formula.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), d.get('families_match_oracle_golden'), d['phases_ms_per_step'], d['initial_sort'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()}, (d.get('roofline') or {}).get('kernel','')[:30], (d.get('roofline') or {}).get('frac'))
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2c19_bench_c4.json 2> gpurun_out/r2c19_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2c19_bench_c4.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY
export ASGART_B200_MSD_MIN=0
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msd_local_kernel' -s 2 -c 1 -f -o gpurun_out/r2_msd_prof_local_v3 tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof6.log 2>&1
echo "ncu rc=$?"

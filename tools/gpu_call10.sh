#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c10_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c10_pytest.txt
ASGART_B200_DEBUG_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c10_bench_n2_phases.json 2> gpurun_out/r2c10_bench_n2_phases.err; echo "rc=$?"
grep "sa_build r0" gpurun_out/r2c10_bench_n2_phases.err | tail -7
grep "sa_build r1" gpurun_out/r2c10_bench_n2_phases.err | tail -7
python - <<'PY'
import json
for line in open('gpurun_out/r2c10_bench_n2_phases.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY

#!/bin/bash
mkdir -p gpurun_out
N=$1
ASGART_B200_DEBUG_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2c11_bench_n${N}_phases.json 2> gpurun_out/r2c11_bench_n${N}_phases.err; echo "rc=$?"
grep "sa_build r0" gpurun_out/r2c11_bench_n${N}_phases.err | tail -7
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c11_bench_n${N}.json 2> gpurun_out/r2c11_bench_n${N}.err; echo "rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for line in open(f'gpurun_out/r2c11_bench_n{N}.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY

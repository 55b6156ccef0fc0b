#!/bin/bash
# lean scaling run at N GPUs: (N = 2: the two-process tests first) one bench run, then three steps with the phase log
mkdir -p gpurun_out
N=$1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2t_pytest_dist_n2.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest_dist_n2.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2t_bench_n${N}.json 2> gpurun_out/r2t_bench_n${N}.err; echo "rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for line in open(f'gpurun_out/r2t_bench_n{N}.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['families_sha256'][:16], d['phases_ms_per_step'])
PY
ASGART_B200_DEBUG_PHASES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2t_bench_n${N}_phases.json 2> gpurun_out/r2t_bench_n${N}_phases.err; echo "rc=$?"
grep "sa_build r0" gpurun_out/r2t_bench_n${N}_phases.err | tail -7

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2c9_pytest_dist.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c9_pytest_dist.txt
ASGART_B200_DEBUG_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c9_bench_n2_phases.json 2> gpurun_out/r2c9_bench_n2_phases.err; echo "rc=$?"
grep "sa_build r0" gpurun_out/r2c9_bench_n2_phases.err | tail -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench_n2.json'))
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step'])
PY

#!/bin/bash
# 64-bit indices: C3 forced to u64 on one GPU, then (8 GPUs) a 4.5 Gbp strand that needs them
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --config 3 --index-bits 64 --steps 5 --warmup 3 --no-ingest --no-cpu-baseline --check-sa > gpurun_out/r2c16_bench_c3_u64.json 2> gpurun_out/r2c16_bench_c3_u64.err; echo "rc=$?"
  timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2c16_bench_c3_u32.json 2> gpurun_out/r2c16_bench_c3_u32.err; echo "rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:msd_|probe_search|gather_rank' -c 60 --csv --log-file gpurun_out/r2c16_traffic_c4.csv python tools/quick_bench.py 4 0 1 > gpurun_out/r2c16_traffic_c4.log 2>&1; echo "ncu rc=$?"
  for f in c3_u64 c3_u32; do python - $f <<'PY'
import json,sys
for line in open(f'gpurun_out/r2c16_bench_{sys.argv[1]}.json'):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d['dtype'], round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['families_match_oracle_golden'], d['device_bytes_peak'], d['sa_check_violations'], d['roofline']['kernel'][:30], round(d['roofline']['frac'],3))
PY
  done
else
  timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --config 4 --scale-n 4500000000 --steps 3 --warmup 3 --check-sa > gpurun_out/r2c16_bench_4p5g_n$N.json 2> gpurun_out/r2c16_bench_4p5g_n$N.err; echo "rc=$?"
  tail -3 gpurun_out/r2c16_bench_4p5g_n$N.err
  python - $N <<'PY'
import json,sys
for line in open(f'gpurun_out/r2c16_bench_4p5g_n{sys.argv[1]}.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['dtype'], d['config']['strand_bp'], round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['device_bytes_peak'], d['sa_check_violations'], d['counters_per_step'])
PY
fi

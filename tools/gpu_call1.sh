#!/bin/bash
# round 2, call 1: parity at full size (goldens), sanitizer logs, a short C4 bench with the digest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2c1_env.txt; nproc >> gpurun_out/r2c1_env.txt; free -g >> gpurun_out/r2c1_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.txt
tail -25 gpurun_out/r2c1_pytest.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2c1_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r2c1_sanitizer_$tool.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench_c4.json 2> gpurun_out/r2c1_bench_c4.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2c1_bench_c4.json

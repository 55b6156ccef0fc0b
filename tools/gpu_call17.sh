#!/bin/bash
mkdir -p gpurun_out
export ASGART_B200_MSD_MIN=0
L=gpurun_out/r2_msd_bench3.log; : > $L
run() { echo "== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run compute-sanitizer --tool memcheck --print-limit 10 tools/msd_bench 300000 18 1
run compute-sanitizer --tool racecheck --print-limit 10 tools/msd_bench 100000 18 1
run tools/msd_bench 1000001 18 0
run tools/msd_bench 57227416 16 1
run tools/msd_bench 3000000000 18 0
grep -E "^==|rc=|OK|FAIL|best|per rep|ERROR SUMMARY|RACECHECK SUMMARY" $L
unset ASGART_B200_MSD_MIN
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c17_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c17_pytest.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2c17_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2c17_sanitizer_$tool.log
done
ASGART_B200_MSD_MIN=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2c17_sanitizer_memcheck_msd.log 2>&1; echo "memcheck(msd forced) rc=$?"; tail -3 gpurun_out/r2c17_sanitizer_memcheck_msd.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c17_bench_c4.json 2> gpurun_out/r2c17_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2c17_bench_c4.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['families_match_oracle_golden'], d['phases_ms_per_step']); print(d['roofline']); print({k:(round(v['ms_per_step'],2), round(v['frac_of_hbm_peak'],3)) for k,v in d['kernel_families'].items()})
PY

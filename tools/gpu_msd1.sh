#!/bin/bash
# msd sort: correctness (sanitizer on small inputs, self-checks on growing ones) and speed against the LSD path
mkdir -p gpurun_out
L=gpurun_out/r2_msd_bench.log
: > $L
run() { echo "== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
export ASGART_B200_MSD_MIN=0
run compute-sanitizer --tool memcheck --print-limit 10 tools/msd_bench 300000 18 1
run compute-sanitizer --tool racecheck --print-limit 10 tools/msd_bench 200000 18 1
run tools/msd_bench 1000000 18 0
run tools/msd_bench 3000001 16 2
run tools/msd_bench 5000003 21 1
run tools/msd_bench 57227416 16 1
run tools/msd_bench 250000000 17 1
run tools/msd_bench 3000000000 18 1
grep -E "^==|rc=|OK|FAIL|best|per rep|ERROR SUMMARY|RACECHECK" $L

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_c4.csv python tools/quick_bench.py 4 0 2 > gpurun_out/r2_launches_c4.log 2>&1; echo rc=$?

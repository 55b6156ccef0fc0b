#!/bin/bash
# DRAM bytes + durations of the sort / search / gather kernels of one C4 step (profiles/traffic.json is made from this csv)
mkdir -p gpurun_out
timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:msd_|probe_search|gather_rank' --csv --log-file gpurun_out/r2j_traffic_c4.csv python tools/quick_bench.py 4 0 1 > gpurun_out/r2j_traffic_c4.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/r2j_traffic_c4.log | cut -c1-200; grep -c "" gpurun_out/r2j_traffic_c4.csv

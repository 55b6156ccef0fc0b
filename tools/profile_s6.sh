#!/bin/bash
# Session profile pass (run under gpurun, one GPU): C4 phase times, ncu launch list of the default bench command,
# ncu --set full captures of the hot kernels at C3 size (250 Mbp: beyond L2, small enough for ncu's replay save/restore).
tag=${1:-s6}
mkdir -p gpurun_out
ASGART_B200_DEBUG_PHASES=1 timeout 300 python tools/quick_bench.py 4 0 2 > gpurun_out/${tag}_phases_c4.log 2>&1
tail -3 gpurun_out/${tag}_phases_c4.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${tag}_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_c4.out 2>&1
tail -c 300 gpurun_out/${tag}_launches_c4.out
i=0
for k in rs_scatter_kernel gather_rank_kernel probe_search_kernel rs_hist_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 2 -f \
      -o "gpurun_out/${tag}_full_${i}_${k}" python tools/quick_bench.py 3 0 1 > "gpurun_out/${tag}_full_${i}.log" 2>&1
  tail -1 "gpurun_out/${tag}_full_${i}.log" | cut -c1-200
  i=$((i+1))
done
ls -la gpurun_out | tail -20

#!/bin/bash
# One `ncu --set full` capture per hot kernel of a C2 pass (run under gpurun, one GPU).
#   tools/profile_full.sh <tag> [kernel-regex ...]
# Reports land in gpurun_out/<tag>_<n>.ncu-rep; summarise here with tools/ncu_summary.py.
tag=${1:-prof}; shift
kernels=("$@")
if [ ${#kernels[@]} -eq 0 ]; then
  kernels=(rs_scatter_kernel probe_search_kernel automaton_segment_kernel rank_scan init_keys lut_build_kernel)
fi
mkdir -p gpurun_out
i=0
for k in "${kernels[@]}"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 2 -f \
      -o "gpurun_out/${tag}_${i}_${k//[^a-zA-Z0-9_]/}" python tools/quick_bench.py 2 0 1 > "gpurun_out/${tag}_${i}.log" 2>&1
  tail -2 "gpurun_out/${tag}_${i}.log" | cut -c1-200
  i=$((i+1))
done

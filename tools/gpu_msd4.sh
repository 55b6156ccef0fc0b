#!/bin/bash
mkdir -p gpurun_out
export ASGART_B200_MSD_MIN=0
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msd_local_kernel' -s 2 -c 1 -f -o gpurun_out/r2_msd_prof_local_fast tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof2.log 2>&1
echo "ncu rc=$?"
ASGART_B200_MSD_LOCAL=slow timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msd_local_kernel' -s 2 -c 1 -f -o gpurun_out/r2_msd_prof_local_slow tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof3.log 2>&1
echo "ncu rc=$?"
ASGART_B200_MSD_LOCAL=slow tools/msd_bench 3000000000 18 0 | grep -E "per rep|best|OK|FAIL"

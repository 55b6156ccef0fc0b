#!/bin/bash
mkdir -p gpurun_out
export ASGART_B200_MSD_MIN=0
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msd_local_kernel' -s 2 -c 1 -f -o gpurun_out/r2_msd_prof_local_rank tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof4.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msd_scatter_kernel' -s 2 -c 1 -f -o gpurun_out/r2_msd_prof_scatter tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof5.log 2>&1
echo "ncu rc=$?"

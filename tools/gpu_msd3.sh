#!/bin/bash
mkdir -p gpurun_out
export ASGART_B200_MSD_MIN=0
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:msd_(local|scatter|hist)_kernel' -c 6 -f -o gpurun_out/r2_msd_prof1 tools/msd_bench 3000000000 18 0 > gpurun_out/r2_msd_prof1.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2_msd_prof1.log; ls -la gpurun_out/r2_msd_prof1.ncu-rep

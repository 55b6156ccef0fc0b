#!/bin/bash
mkdir -p gpurun_out
export ASGART_B200_MSD_MIN=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_msd_launches_250m.csv tools/msd_bench 250000000 17 1 > gpurun_out/r2_msd_launches_250m.log 2>&1
echo rc=$?

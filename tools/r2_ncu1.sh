#!/bin/bash
# ncu: launch list of a 40 MB one-stream encode + one full capture (with source counters) of stage A v7
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv \
    python tools/v7_probe.py 40000000 2 > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orz_parse7 -s 2 -c 1 -o gpurun_out/r2_parse7_full -f \
    python tools/v7_probe.py 3000000 2 > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
tail -3 gpurun_out/r2_ncu_full.log

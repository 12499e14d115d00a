#!/bin/bash
# ncu: one full capture (with source counters) each of stage A v7 (3 MB one-stream encode: one launch) and of the decoder (2 MB)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:orz_parse7 -s 1 -c 1 -o gpurun_out/r2_parse7_full -f \
    python tools/v7_probe.py 3000000 2 > gpurun_out/r2_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:orz_decode_kernel -s 1 -c 1 -o gpurun_out/r2_decode_full -f \
    python tools/dec_probe.py 2000000 2 > gpurun_out/r2_ncu_dec.log 2>&1
ls -la gpurun_out | tail -6
tail -3 gpurun_out/r2_ncu_full.log gpurun_out/r2_ncu_dec.log

#!/bin/bash
# event trace of the position where the 100 MB stream sometimes differs (window position 21087536 of block 4)
set -u
mkdir -p gpurun_out
timeout 600 python tools/trace_probe.py 72000000 2 12 ORZ_B200_STATS=1 ORZ_B200_DBG_POS=21087536 ORZ_B200_DBG_INPOS=72000000 2>&1 | grep -E "trace|== rep" > gpurun_out/p13_trace.log
grep -c trace gpurun_out/p13_trace.log; grep "== rep" gpurun_out/p13_trace.log

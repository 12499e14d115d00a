#!/bin/bash
# production-timing build with the event trace compiled in: the position where the 100 MB stream sometimes differs
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_rep*.bin
ORZ_B200_LIB=$PWD/build_alt/liborz_trace.so timeout 900 python tools/trace_probe.py 100000000 2 26 ORZ_B200_DBG_POS=21087536 ORZ_B200_DBG_INPOS=83886080 2>&1 | grep -E "trace|== rep" > gpurun_out/p14_trace.log
grep "== rep" gpurun_out/p14_trace.log | grep -c True; grep "== rep" gpurun_out/p14_trace.log | grep False

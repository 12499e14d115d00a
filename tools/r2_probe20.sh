#!/bin/bash
# record checksums: the 17 MB / lead 160 reproducer (7 of 120 before), torn-read counters, a few 100 MB runs
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin
{
echo "== lead160 17 MB x120"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 400 python tools/stress_probe.py 17000000 2 120 2>&1 | tail -1
echo "== lead160 STATS 17 MB"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 120 python tools/v7_probe.py 17000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default STATS 6 MB"; timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default 6 MB"; timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
echo "== lead160 6 MB"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
echo "== lead160 100 MB x6"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 400 python tools/stress_probe.py 100000000 2 6 2>&1 | tail -1
} > gpurun_out/p20.log 2>&1
cat gpurun_out/p20.log

#!/bin/bash
# stress: how often does the 100 MB stream differ, with and without the refresh protocol
set -u
mkdir -p gpurun_out
{
echo "== default x12"; timeout 400 python tools/stress_probe.py 100000000 2 12 2>&1 | tail -1
echo "== norefresh x12"; ORZ_B200_LIB=$PWD/build_alt/liborz_norefresh.so timeout 400 python tools/stress_probe.py 100000000 2 12 2>&1 | tail -1
echo "== noedge x8"; ORZ_B200_LIB=$PWD/build_alt/liborz_noedge.so timeout 400 python tools/stress_probe.py 100000000 2 8 2>&1 | tail -1
} > gpurun_out/p11_stress.log 2>&1
cat gpurun_out/p11_stress.log

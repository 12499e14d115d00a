#!/bin/bash
# table-based verification + longer leads; stress of the 100 MB identity with the failing stream kept for tools/diff_parse.py
set -u
mkdir -p gpurun_out
{
echo "== default STATS"; timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default"; timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
for f in build_alt/liborz_*.so; do
  echo "== $f"
  ORZ_B200_LIB=$PWD/$f timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
done
echo "== default 100 MB"; timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
echo "== lead160 100 MB"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
} > gpurun_out/p12_probe.log 2>&1
{
echo "== default x10"; timeout 400 python tools/stress_probe.py 100000000 2 10 2>&1 | tail -1
} > gpurun_out/p12_stress.log 2>&1
cat gpurun_out/p12_probe.log gpurun_out/p12_stress.log

#!/bin/bash
# bisect of the 100 MB -l2 identity failure of probe 9: edge-exact and refresh switched off separately and together
set -u
mkdir -p gpurun_out
{
for f in build_alt/liborz_noedge_norefresh.so build_alt/liborz_norefresh.so build_alt/liborz_noedge.so; do
  echo "== $f 100 MB"
  ORZ_B200_LIB=$PWD/$f timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
done
echo "== default 100 MB again"; timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
echo "== default 40 MB"; timeout 200 python tools/v7_probe.py 40000000 2 2>&1 | tail -1
} > gpurun_out/p10_probe.log 2>&1
cat gpurun_out/p10_probe.log

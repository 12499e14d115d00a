#!/bin/bash
# after the builder fence fix: stress of the 100 MB identity (default and lead 160), lead A/B, ncu capture of stage A
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin
{
echo "== default STATS"; timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default"; timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
for f in build_alt/liborz_*.so; do
  echo "== $f"
  ORZ_B200_LIB=$PWD/$f timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
done
echo "== default 100 MB"; timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
echo "== lead160 100 MB"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
} > gpurun_out/p16_probe.log 2>&1
{
echo "== default x12"; timeout 400 python tools/stress_probe.py 100000000 2 12 2>&1 | tail -1
echo "== lead160 x14"; ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 400 python tools/stress_probe.py 100000000 2 14 2>&1 | tail -1
} > gpurun_out/p16_stress.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:orz_parse7 -s 1 -c 1 -o gpurun_out/r2_parse7_full2 -f \
    python tools/v7_probe.py 3000000 2 > gpurun_out/p16_ncu.log 2>&1
cat gpurun_out/p16_probe.log gpurun_out/p16_stress.log; tail -2 gpurun_out/p16_ncu.log

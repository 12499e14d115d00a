#!/bin/bash
# A/B builds (tools/build_alt.sh) of stage A on 6 MB of enwik8: STATS line of the default build, then the parse time of every variant
set -u
mkdir -p gpurun_out
{
echo "== default STATS"; ORZ_B200_STATS=1 timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default"; timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
for f in build_alt/liborz_*.so; do
  echo "== $f"
  ORZ_B200_LIB=$PWD/$f timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
done
echo "== muttight 100 MB (a wrong ring-edge check must not stay bit-identical)"
ORZ_B200_LIB=$PWD/build_alt/liborz_muttight.so timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
echo "== default 100 MB"; timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
} > gpurun_out/p8_probe.log 2>&1
cat gpurun_out/p8_probe.log
{
echo "== decode 20 MB STATS"; timeout 120 python tools/dec_probe.py 20000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== decode 20 MB no rank cache"; timeout 120 python tools/dec_probe.py 20000000 2 ORZ_B200_DEC_CACHE=0 2>&1 | tail -1
} > gpurun_out/p8_dec.log 2>&1
cat gpurun_out/p8_dec.log

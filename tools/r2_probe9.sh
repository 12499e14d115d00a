#!/bin/bash
# stage A with refresh requests + A/B builds; decoder in on-chip mode; ncu captures of both kernels
set -u
mkdir -p gpurun_out
{
echo "== default STATS"; timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== default"; timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
for f in build_alt/liborz_*.so; do
  echo "== $f"
  ORZ_B200_LIB=$PWD/$f timeout 120 python tools/v7_probe.py 6000000 2 2>&1 | tail -1
done
echo "== norefresh STATS"; ORZ_B200_LIB=$PWD/build_alt/liborz_norefresh.so timeout 120 python tools/v7_probe.py 6000000 2 ORZ_B200_STATS=1 2>&1 | tail -2 | head -1
echo "== default 100 MB"; timeout 200 python tools/v7_probe.py 100000000 2 2>&1 | tail -1
echo "== default 100 MB -l0"; timeout 200 python tools/v7_probe.py 100000000 0 2>&1 | tail -1
} > gpurun_out/p9_probe.log 2>&1
{
echo "== decode 20 MB STATS"; timeout 120 python tools/dec_probe.py 20000000 2 ORZ_B200_STATS=1 2>&1 | tail -2
echo "== decode 100 MB"; timeout 200 python tools/dec_probe.py 100000000 2 2>&1 | tail -1
} > gpurun_out/p9_dec.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "decode or roundtrip or stream or smoke" > gpurun_out/p9_pytest.log 2>&1
bash tools/r2_ncu2.sh > gpurun_out/p9_ncu.log 2>&1
cat gpurun_out/p9_probe.log gpurun_out/p9_dec.log; tail -3 gpurun_out/p9_pytest.log; tail -8 gpurun_out/p9_ncu.log

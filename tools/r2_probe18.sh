#!/bin/bash
# lead 160: find the differing step (17 MB first, else 100 MB), then trace the three records behind it in the same configuration
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin
N=17000000; INPOS=16777216
ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 300 python tools/find_diff_pos.py $N 2 14 > gpurun_out/p18_find.log 2>&1
X=$(grep -m1 "^POS" gpurun_out/p18_find.log | awk '{print $2}')
if [ "$X" = "none" ] || [ -z "$X" ]; then
  N=100000000
  ORZ_B200_LIB=$PWD/build_alt/liborz_lead160.so timeout 400 python tools/find_diff_pos.py $N 2 10 >> gpurun_out/p18_find.log 2>&1
  X=$(grep "^POS" gpurun_out/p18_find.log | tail -1 | awk '{print $2}')
fi
echo "N $N X $X" >> gpurun_out/p18_find.log
if [ "$X" != "none" ] && [ -n "$X" ]; then
  BLK=$((X / 16777216)); W=$((X - BLK * 16777216 + 16777215))
  if [ $N -gt $(( (BLK + 1) * 16777216 )) ]; then INPOS=$(( (BLK + 1) * 16777216 )); else INPOS=$N; fi
  echo "window position $W in_pos $INPOS" >> gpurun_out/p18_find.log
  ORZ_B200_LIB=$PWD/build_alt/liborz_trace160.so timeout 400 python tools/trace_probe.py $N 2 12 ORZ_B200_DBG_POS=$W ORZ_B200_DBG_INPOS=$INPOS 2>&1 | grep -E "trace|== rep" > gpurun_out/p18_trace.log
fi
cat gpurun_out/p18_find.log; grep "== rep" gpurun_out/p18_trace.log

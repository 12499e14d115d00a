#!/bin/bash
# lead 160 reproduces a difference in chunk 3 of the corpus four times in ten: provenance of the differing step, then the event trace
# of the three records behind it (with the ext words the repair saw)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin
ORZ_B200_LIB=$PWD/build_alt/liborz_prov160.so timeout 300 python tools/prov_probe.py 17000000 2 10 > gpurun_out/p17_prov.log 2>&1
X=$(grep -m1 -o "> item [0-9]* pos [0-9]*" gpurun_out/p17_prov.log | awk '{print $5}')
echo "differing step at input position: $X" >> gpurun_out/p17_prov.log
if [ -n "$X" ]; then
  W=$((X + 16777215))
  ORZ_B200_LIB=$PWD/build_alt/liborz_trace160.so timeout 300 python tools/trace_probe.py 17000000 2 12 ORZ_B200_DBG_POS=$W ORZ_B200_DBG_INPOS=16777216 2>&1 | grep -E "trace|== rep" > gpurun_out/p17_trace.log
fi
cat gpurun_out/p17_prov.log; grep "== rep" gpurun_out/p17_trace.log

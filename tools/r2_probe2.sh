#!/bin/bash
# round 2, GPU call 2: stage A v7 with in-place repairs, split committers, private record slots
set -u
mkdir -p gpurun_out
{
for args in "300000 2" "6000000 2 ORZ_B200_STATS=1" "6000000 2" "6000000 2 ORZ_B200_V7_CLUSTER=16 ORZ_B200_STATS=1" "6000000 2 ORZ_B200_V7_CLUSTER=1 ORZ_B200_STATS=1" "6000000 0 ORZ_B200_STATS=1" "40000000 2 ORZ_B200_STATS=1" "100000000 2" "100000000 0"; do
  echo "== $args"
  timeout 300 python tools/v7_probe.py $args 2>&1 | tail -3
done
} > gpurun_out/p2_probe.log 2>&1
tail -30 gpurun_out/p2_probe.log

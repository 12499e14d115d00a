#!/bin/bash
# round 2, first GPU call: stage A v7 correctness + counters on the B200
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/p1_smi.txt 2>&1
{
for args in "300000 0" "300000 2" "6000000 2 ORZ_B200_STATS=1" "6000000 2" "6000000 2 ORZ_B200_V7_CLUSTER=16" "6000000 2 ORZ_B200_V7_CLUSTER=1" "6000000 2 ORZ_B200_LEGACY_STAGE_A=1" "40000000 2 ORZ_B200_STATS=1" "100000000 2" "100000000 0"; do
  echo "== $args"
  timeout 600 python tools/v7_probe.py $args 2>&1 | tail -4
done
} > gpurun_out/p1_probe.log 2>&1
tail -40 gpurun_out/p1_probe.log

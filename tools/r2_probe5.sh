#!/bin/bash
set -u
mkdir -p gpurun_out
{
for args in "6000000 2 ORZ_B200_STATS=1" "6000000 2" "6000000 0 ORZ_B200_STATS=1"; do
  echo "== $args"
  timeout 60 python tools/v7_probe.py $args 2>&1 | tail -2
done
for args in "100000000 2" "100000000 1"; do
  echo "== $args"
  timeout 150 python tools/v7_probe.py $args 2>&1 | tail -2
done
} > gpurun_out/p5_probe.log 2>&1
tail -30 gpurun_out/p5_probe.log

#!/bin/bash
set -u
mkdir -p gpurun_out
{
for args in "6000000 2 ORZ_B200_STATS=1" "6000000 2"; do
  echo "== $args"
  timeout 60 python tools/v7_probe.py $args 2>&1 | tail -2
done
} > gpurun_out/p6_probe.log 2>&1
tail -30 gpurun_out/p6_probe.log

#!/bin/bash
# per-step provenance: which path produced the record of the step that differs from the oracle's parse
set -u
mkdir -p gpurun_out
ORZ_B200_LIB=$PWD/build_alt/liborz_prov.so timeout 1100 python tools/prov_probe.py 100000000 2 28 > gpurun_out/p15_prov.log 2>&1
cat gpurun_out/p15_prov.log

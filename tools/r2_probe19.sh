#!/bin/bash
# lead 160 + builder/repair provenance on the 17 MB reproducer, many repetitions
set -u
mkdir -p gpurun_out
ORZ_B200_LIB=$PWD/build_alt/liborz_prov160.so timeout 500 python tools/prov_probe.py 17000000 2 120 > gpurun_out/p19_prov.log 2>&1
cat gpurun_out/p19_prov.log | tail -40

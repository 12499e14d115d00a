#!/bin/bash
# final candidate (lead 160, record checksums): the whole -m gpu suite, then the 17 MB reproducer with the incomplete-record counters
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p21_pytest.log 2>&1
tail -3 gpurun_out/p21_pytest.log
{
echo "== 17 MB x100"; ORZ_B200_TORN=1 timeout 400 python tools/stress_probe.py 17000000 2 100 2>&1 | sort | uniq -c | sort -rn | head -8
} > gpurun_out/p21_stress.log 2>&1
cat gpurun_out/p21_stress.log | cut -c1-400

#!/bin/bash
# round-2 GPU visit: stage-A probes, then the whole -m gpu suite, smoke() and both bench arms
set -u
mkdir -p gpurun_out
{
for args in "6000000 2 ORZ_B200_STATS=1" "6000000 2" "100000000 2"; do
  echo "== $args"
  timeout 120 python tools/v7_probe.py $args 2>&1 | tail -2
done
} > gpurun_out/p7_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p7_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/p7_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/p7_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/p7_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/p7_bench.log
timeout 400 python bench.py --impl reference > gpurun_out/p7_bench_ref.log 2>&1
tail -8 gpurun_out/p7_probe.log; tail -5 gpurun_out/p7_pytest.log; tail -2 gpurun_out/p7_smoke.log; tail -3 gpurun_out/p7_bench.log; tail -2 gpurun_out/p7_bench_ref.log

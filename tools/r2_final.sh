#!/bin/bash
# end of round 2: bench line (both arms), ncu launch list of the bench command, one full capture each of stage A and the decoder
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bad_stream_*.bin gpurun_out/*.ncu-rep
timeout 600 python bench.py --gpus 1 --steps 1 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 1 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_final.csv \
    python tools/v7_probe.py 40000000 2 > gpurun_out/r2_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:orz_parse7 -s 1 -c 1 -o gpurun_out/r2_parse7_final -f \
    python tools/v7_probe.py 3000000 2 > gpurun_out/r2_ncu_full.log 2>&1
cat gpurun_out/r2_bench_n1.json | cut -c1-3000; tail -2 gpurun_out/r2_ncu_full.log; ls -la gpurun_out | tail -8

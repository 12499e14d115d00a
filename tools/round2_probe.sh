#!/bin/bash
# First gpurun call of round 2 (one GPU, ~6 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/round2_probe.sh'
# 1. every stage A variant on 6 MB of enwik8 with counters (ORZ_B200_STATS) -> gpurun_out/r2_variants.log
# 2. ncu launch list of a short bench run of the kernel chosen in $ORZ_R2_ENV (default: library default)
# 3. one `--set full` capture of the stage A kernel (traffic, stall mix, source page)
set -u
mkdir -p gpurun_out
python - > gpurun_out/r2_variants.log 2>&1 <<'PY'
import json, subprocess, sys, os
sys.path.insert(0, os.getcwd())
import bench
for env in bench.VARIANTS:
    e = dict(os.environ, ORZ_B200_STATS="1", **env)
    p = subprocess.run([sys.executable, "tools/variant_probe.py", "6000000", "1"] + [f"{k}={v}" for k, v in env.items()],
                       capture_output=True, text=True, timeout=120, env=e)
    print(json.dumps(env), p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "", flush=True)
    print("   ", " | ".join(l for l in p.stderr.splitlines() if "stage A" in l)[-400:], flush=True)
PY
export ${ORZ_R2_ENV:-ORZ_B200_PARSE=4}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python tools/variant_probe.py 12000000 1 ${ORZ_R2_ENV:-ORZ_B200_PARSE=4} > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orz_parse -s 1 -c 1 -o gpurun_out/r2_parse_full -f \
    python tools/variant_probe.py 4000000 1 ${ORZ_R2_ENV:-ORZ_B200_PARSE=4} > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_parse_full.ncu-rep --page raw --csv 2>/dev/null | \
    grep -E 'dram__bytes_(read|write)\.sum|gpu__time_duration|sm__warps_active|launch__registers_per_thread|smsp__issue_active|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate' \
    > gpurun_out/r2_parse_full_summary.csv
tail -3 gpurun_out/r2_variants.log

#!/bin/bash
# Randomized emulator soak (TEST INFRASTRUCTURE): stage A variants x seeds x levels on 300 KB slices of a corpus file,
# each compared with the oracle.  usage: tests/simt/soak.sh CORPUS_FILE [WORKER_ID] [LOG]   (run several workers in parallel)
# A seed that produces a DIFFERENT stream is a deterministic reproducer: rerun emu_cli with the same SIMT_SEED and switches.
set -u
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
cd "$ROOT"
corpus=$1
worker=${2:-0}
log=${3:-/tmp/orz_soak.log}
export ORZ_B200_TEST_EMULATOR=1 SIMT_SWITCH_LIMIT=6000000000000
size=$(stat -c %s "$corpus")
seed=$((1000 + worker * 100000))
# (the library reads its switches once per process: one emu_cli run per configuration)
configs=("ORZ_B200_V7_CLUSTER=1" "ORZ_B200_V7_CLUSTER=8" "ORZ_B200_V7_CLUSTER=1 ORZ_B200_STATS=1" "ORZ_B200_LEGACY_STAGE_A=v4" "ORZ_B200_LEGACY_STAGE_A=v3"
         "ORZ_B200_LEGACY_STAGE_A=v3s")
i=0
while true; do
  cfg=${configs[$(( (i + worker) % ${#configs[@]} ))]}
  lvl=$(( (i + worker) % 3 ))
  off=$(( ( (i * 7 + worker * 13) * 1000000 ) % (size > 400000 ? size - 400000 : 1) ))
  seed=$((seed + 1))
  in=/tmp/orz_soak_in.$worker
  dd if="$corpus" of=$in bs=1000 skip=$((off / 1000)) count=300 2>/dev/null
  ./oracle/orz_oracle encode $lvl $in /tmp/orz_soak_ref.$worker.orz >/dev/null 2>&1
  env SIMT_SEED=$seed $cfg ./tests/simt/emu_cli $lvl $in /tmp/orz_soak_out.$worker.orz >/dev/null 2>&1
  if cmp -s /tmp/orz_soak_out.$worker.orz /tmp/orz_soak_ref.$worker.orz; then r=identical; else r=DIFFERENT; cp $in /tmp/orz_soak_fail.$worker.$i.in; fi
  echo "$(date +%H:%M:%S) w$worker seed $seed l$lvl off $off [$cfg]: $r" >> "$log"
  i=$((i + 1))
done

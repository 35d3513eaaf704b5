#!/bin/bash
# Development aid: run the throughput probe under several REBOP_B200_CODEGEN settings.
# usage: scripts/sweep.sh model n tmax nb arith cfg1 cfg2 ...
model=$1; n=$2; tmax=$3; nb=$4; arith=$5; shift 5
for cfg in "$@"; do
  echo "== $cfg"
  REBOP_B200_CODEGEN="$cfg" timeout 120 python scripts/perf_probe.py $model $n 2 $tmax $nb $arith noprobe 2>&1 | tail -1
done

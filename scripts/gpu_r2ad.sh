#!/bin/bash
# round 2, GPU session AD: A/B of the NaN-uniform "no event" form (first-match networks) on Vilar and Dimers
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
for rep in 1 2; do
for knobs in "" "defs=RB_X_NO_NANPICK"; do
  probe "$knobs" vilar 1250000 2 200 200 1
  probe "$knobs" vilar 1326080 2 200 200 1
  probe "$knobs" dimers 1000000 2 1 1 1
done
done
} 2>&1 | tee $OUT/r2ad_sweep.log

#!/bin/bash
# round 2, GPU session C: occupancy / state-form sweep of the register-resident kernel (NVRTC path, knobs through REBOP_B200_CODEGEN)
OUT=gpurun_out
mkdir -p $OUT
N=${N:-284160}
probe() {  # $1 = knobs, rest = perf_probe args
  local knobs="$1"; shift
  echo "-- [$knobs] $*"
  REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1
}
{
for knobs in "" "conv=1" "conv=1,minctas=6" "minctas=6" "conv=1,minctas=7" "block=64,minctas=10" "conv=1,block=64,minctas=12" "conv=1,block=256,minctas=3" "tick=32" "conv=1,minctas=6,tick=32"; do
  REBOP_B200_SCHEDULE=sparse probe "$knobs" vilar $N 2 200 200 1
done
for knobs in "" "conv=1,minctas=6" "conv=1,minctas=8" "minctas=8"; do
  REBOP_B200_SCHEDULE=dense probe "$knobs" sir 1000000 2 250 250 0
  REBOP_B200_SCHEDULE=sparse probe "$knobs" dimers 1000000 2 1 1 1
done
} 2>&1 | tee $OUT/r2c_sweep.log

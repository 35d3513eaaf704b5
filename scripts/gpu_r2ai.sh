#!/bin/bash
# round 2, GPU session AI: absorbing branches of the crossing block merged; every model under both claiming schedules
OUT=gpurun_out
mkdir -p $OUT
probe() { echo "-- [$REBOP_B200_SCHEDULE] $*"; timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
for sched in sparse dense; do
  export REBOP_B200_SCHEDULE=$sched
  probe vilar 1250000 3 200 200 1
  probe dimers 1000000 3 1 1 1
  probe sir 1000000 3 250 250 1
  probe mm_lma 1000000 2 100 100 0
done
unset REBOP_B200_SCHEDULE
} 2>&1 | tee $OUT/r2ai_sweep.log
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_system.py tests/test_frontend.py -q -m gpu -x 2>&1 | tail -3 | tee -a $OUT/r2ai_sweep.log

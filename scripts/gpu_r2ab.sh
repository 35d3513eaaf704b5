#!/bin/bash
# round 2, GPU session AB: running record pointer in the dense schedule; totals outside the short divide's range
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
probe "" sir 1000000 3 250 250 1
probe "" sir 1000000 2 250 250 0
probe "" mm_lma 1000000 2 100 100 0
REBOP_B200_SCHEDULE=dense probe "" mm_lma 1000000 2 100 100 0
probe "" vilar 1250000 3 200 200 1
} 2>&1 | tee $OUT/r2ab_sweep.log
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 | tee -a $OUT/r2ab_sweep.log

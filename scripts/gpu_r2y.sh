#!/bin/bash
# round 2, GPU session Y: reversed stoichiometry rows / masked uniform (166-instruction pass), unroll and tick knobs
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
probe "" vilar 1250000 3 200 200 1
for knobs in "" "unroll=2" "tick=64" "tick=16" "minctas=4" "block=64,minctas=10"; do
  probe "$knobs" vilar 1250000 2 200 200 1
done
probe "" dimers 1000000 3 1 1 1
probe "" sir 1000000 3 250 250 1
probe "" mm_lma 1000000 2 100 100 0
} 2>&1 | tee $OUT/r2y_sweep.log
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_system.py -q -m gpu -x 2>&1 | tail -3 | tee -a $OUT/r2y_sweep.log

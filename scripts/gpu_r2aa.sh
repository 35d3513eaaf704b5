#!/bin/bash
# round 2, GPU session AA: wide ziggurat tables (slow path bounds as one fma each): parity + throughput
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
probe "" vilar 1250000 3 200 200 1
probe "" vilar 1250000 2 200 200 1
probe "" dimers 1000000 3 1 1 1
probe "" sir 1000000 3 250 250 1
probe "" mm_lma 1000000 2 100 100 0
probe "" synthetic 300000 2 1 1 0
} 2>&1 | tee $OUT/r2aa_sweep.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/r2aa_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2aa_pytest_gpu.log

#!/bin/bash
# round 2, GPU session X: the shorter pass as the default (all kernels): full GPU test suite, throughput of every config
OUT=gpurun_out
mkdir -p $OUT
probe() { echo "-- $*"; timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
probe vilar 1250000 3 200 200 1
probe dimers 1000000 3 1 1 1
probe sir 1000000 3 250 250 1
probe sir 1000000 2 250 250 0
probe mm_lma 1000000 2 100 100 0
probe synthetic 300000 2 1 1 0
probe vilar 300000 1 20 20 1
} 2>&1 | tee $OUT/r2x_sweep.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/r2x_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2x_pytest_gpu.log

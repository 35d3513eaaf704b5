#!/bin/bash
# round 2, GPU session AC: the shorter pass as the default (all kernels): full GPU test suite, throughput of every config
TAG=${1:-r2ac}
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
} 2>&1 | tee $OUT/${TAG}_sweep.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest_gpu.log

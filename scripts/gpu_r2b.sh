#!/bin/bash
# round 2, GPU session B: full parity suite after the resume fix, bench line of the headline config
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/r2b_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu > $OUT/r2b_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -15 $OUT/r2b_pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r2b_bench.json 2> $OUT/r2b_bench.err ; echo "rc=$?"; cat $OUT/r2b_bench.json; tail -5 $OUT/r2b_bench.err

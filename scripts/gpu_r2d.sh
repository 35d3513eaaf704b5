#!/bin/bash
# round 2, GPU session D: warp-wise claiming at the end of the ensemble (sweep of the threshold), parity suite, new bench line
OUT=gpurun_out
mkdir -p $OUT
{
for eg in 0 25 50 75 100; do
  for cfg in "vilar 1250000 3 200 200 1" "vilar 284160 3 200 200 1" "sir 1000000 0 250 250 0" "dimers 1000000 0 1 1 1" "mm_lma 1000000 0 100 100 0"; do
    echo "-- endgame=$eg $cfg"
    REBOP_B200_ENDGAME=$eg timeout 300 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
  done
done
} 2>&1 | tee $OUT/r2d_endgame.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu > $OUT/r2d_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -8 $OUT/r2d_pytest_gpu.log
echo "== bench" ; timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/r2d_bench.json 2> $OUT/r2d_bench.err ; echo "rc=$?"; cat $OUT/r2d_bench.json; tail -5 $OUT/r2d_bench.err

#!/bin/bash
# round 2, GPU session I: branch-free dense crossing + finish kernel; parity suite; probes
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu > $OUT/r2i_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -8 $OUT/r2i_pytest_gpu.log
{
for cfg in "sir 1000000 0 250 250 0" "mm_lma 1000000 0 100 100 0" "dimers 1000000 0 1 1 1"; do
  for sched in sparse dense; do
    echo "-- $sched $cfg"
    REBOP_B200_SCHEDULE=$sched timeout 200 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
  done
done
echo "-- auto vilar"; timeout 300 python scripts/perf_probe.py vilar 1250000 3 200 200 1 noprobe 2>&1 | tail -1
} 2>&1 | tee $OUT/r2i_probes.log

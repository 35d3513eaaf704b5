#!/bin/bash
# round 2, GPU session A: parity after the schedule/sample-path rewrite, then schedule and store-path probes
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/r2a_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2a_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -15 $OUT/r2a_pytest_gpu.log
echo "== probes"
for sched in static sparse dense; do
  for cfg in "sir 1000000 0 250 250 0" "mm_lma 1000000 0 100 100 0" "dimers 1000000 0 1 1 1"; do
    echo "-- $sched $cfg"
    REBOP_B200_SCHEDULE=$sched timeout 200 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
  done
done 2>&1 | tee $OUT/r2a_probes.log
echo "-- bulk=0 sir dense" | tee -a $OUT/r2a_probes.log
REBOP_B200_BULK_STORE=0 REBOP_B200_SCHEDULE=dense timeout 200 python scripts/perf_probe.py sir 1000000 0 250 250 0 noprobe 2>&1 | tail -1 | tee -a $OUT/r2a_probes.log
for sched in sparse dense; do
  echo "-- $sched vilar" | tee -a $OUT/r2a_probes.log
  REBOP_B200_SCHEDULE=$sched timeout 300 python scripts/perf_probe.py vilar 1250000 0 200 200 1 noprobe 2>&1 | tail -1 | tee -a $OUT/r2a_probes.log
done
echo "-- auto synthetic" | tee -a $OUT/r2a_probes.log
timeout 300 python scripts/perf_probe.py synthetic 300000 0 0.2 100 0 noprobe 2>&1 | tail -1 | tee -a $OUT/r2a_probes.log

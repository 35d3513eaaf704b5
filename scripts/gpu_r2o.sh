#!/bin/bash
# round 2, GPU session O: finishing kernel with 64 x (512 B) tiles
OUT=gpurun_out
mkdir -p $OUT
{
for cfg in "sir 1000000 0 250 250 0" "mm_lma 1000000 0 100 100 0" "vilar 1250000 3 200 200 1"; do
  echo "-- $cfg"; timeout 300 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
done
echo "-- bulk=0 sir"; REBOP_B200_BULK_STORE=0 timeout 300 python scripts/perf_probe.py sir 1000000 0 250 250 0 noprobe 2>&1 | tail -1
} 2>&1 | tee $OUT/r2o_probes.log
echo "== pytest sample tests"; timeout 900 python -m pytest tests -q -m gpu -k "sample or golden or host or int16 or 65535 or ragged or absorbing or nan" 2>&1 | tail -3

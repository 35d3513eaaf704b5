#!/bin/bash
# round 2, GPU session N: finishing kernel with four reads in flight per warp; host threads of the staged copy
OUT=gpurun_out
mkdir -p $OUT
{
for cfg in "sir 1000000 0 250 250 0" "mm_lma 1000000 0 100 100 0" "vilar 1250000 3 200 200 1"; do
  echo "-- $cfg"; timeout 300 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
done
for th in 4 8 16 32; do
  echo "== REBOP_B200_COPY_THREADS=$th python api probe 1e7"
  REBOP_B200_COPY_THREADS=$th timeout 600 python scripts/python_api_probe.py 1e7 2>&1 | tail -2 | head -1
done
} 2>&1 | tee $OUT/r2n_probes.log
echo "== pytest sample tests"; timeout 900 python -m pytest tests -q -m gpu -k "sample or golden or host or int16 or 65535" 2>&1 | tail -3

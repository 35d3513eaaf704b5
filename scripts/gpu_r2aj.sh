#!/bin/bash
# round 2, GPU session AJ: large form with the products issued a batch ahead of the running sum (knob lbatch=)
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 200 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
for lb in 0 4 8 16; do
  probe "lbatch=$lb" synthetic 300000 2 0.5 1 0
done
} 2>&1 | tee $OUT/r2aj_sweep.log
echo "== parity, large networks, lbatch=8"; REBOP_B200_CODEGEN="lbatch=8" timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "large_networks" -x 2>&1 | tail -2 | tee -a $OUT/r2aj_sweep.log
echo "== parity, large networks, lbatch=4"; REBOP_B200_CODEGEN="lbatch=4" timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "large_networks" -x 2>&1 | tail -2 | tee -a $OUT/r2aj_sweep.log

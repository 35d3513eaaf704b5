#!/bin/bash
# round 2, GPU session AF: K7 (partial propensities) with more checkpoints of the running sum (shorter block walks)
OUT=gpurun_out
mkdir -p $OUT
{
for ck in 32 50 100; do
  echo "-- REBOP_B200_PDM_CK=$ck"
  REBOP_B200_PDM_CK=$ck timeout 300 python scripts/pdm_probe.py synthetic 300000 0.2 1 2>&1 | grep -v "^mean"
done
} 2>&1 | tee $OUT/r2af_pdm.log

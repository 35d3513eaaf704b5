#!/bin/bash
# round 2, GPU session G: partial-propensity kernel with the flat block walk
OUT=gpurun_out
mkdir -p $OUT
{
echo "== pdm synthetic"; timeout 600 python scripts/pdm_probe.py synthetic 300000 2>&1 | tail -8
echo "== pdm vilar"; timeout 600 python scripts/pdm_probe.py vilar 50000 20 20 2>&1 | tail -8
echo "== pytest pdm"; timeout 900 python -m pytest tests/test_pdm.py -q -m gpu 2>&1 | tail -8
} 2>&1 | tee $OUT/r2g_probes.log

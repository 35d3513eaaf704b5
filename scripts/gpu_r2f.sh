#!/bin/bash
# round 2, GPU session F: partial-propensity kernel (codegen form), whole-wave grid sizing
OUT=gpurun_out
mkdir -p $OUT
{
echo "== pdm synthetic"; timeout 600 python scripts/pdm_probe.py synthetic 300000 2>&1 | tail -8
echo "== pdm vilar"; timeout 600 python scripts/pdm_probe.py vilar 50000 20 20 2>&1 | tail -8
echo "== pdm sir"; timeout 600 python scripts/pdm_probe.py sir 100000 2>&1 | tail -8
for w in 0 1; do
  echo "== waves=$w vilar 1.25e6"
  REBOP_B200_WAVES=$w timeout 300 python scripts/perf_probe.py vilar 1250000 3 200 200 1 noprobe 2>&1 | tail -1
  echo "== waves=$w dimers 1e6"
  REBOP_B200_WAVES=$w timeout 300 python scripts/perf_probe.py dimers 1000000 0 1 1 1 noprobe 2>&1 | tail -1
  echo "== waves=$w sir 1e6"
  REBOP_B200_WAVES=$w timeout 300 python scripts/perf_probe.py sir 1000000 0 250 250 0 noprobe 2>&1 | tail -1
  echo "== waves=$w mm 1e6"
  REBOP_B200_WAVES=$w timeout 300 python scripts/perf_probe.py mm_lma 1000000 0 100 100 0 noprobe 2>&1 | tail -1
done
} 2>&1 | tee $OUT/r2f_probes.log

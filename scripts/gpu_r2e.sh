#!/bin/bash
# round 2, GPU session E: latency probe, one warp per scheduler, first run of the dependency-driven kernel
OUT=gpurun_out
mkdir -p $OUT
{
echo "== latency probe"; timeout 120 rebop_b200/csrc/build/lat_probe
echo "== vilar, one CTA per SM (one warp per scheduler), static"
REBOP_B200_SCHEDULE=static timeout 300 python scripts/perf_probe.py vilar 18944 3 200 200 1 noprobe 2>&1 | tail -1
echo "== vilar, two CTAs per SM"
REBOP_B200_SCHEDULE=static timeout 300 python scripts/perf_probe.py vilar 37888 3 200 200 1 noprobe 2>&1 | tail -1
echo "== vilar, full"
REBOP_B200_SCHEDULE=static timeout 300 python scripts/perf_probe.py vilar 94720 3 200 200 1 noprobe 2>&1 | tail -1
echo "== pdm synthetic"; timeout 600 python scripts/pdm_probe.py synthetic 100000 2>&1 | tail -8
echo "== pdm vilar"; timeout 600 python scripts/pdm_probe.py vilar 50000 20 20 2>&1 | tail -8
echo "== pdm sir"; timeout 600 python scripts/pdm_probe.py sir 100000 2>&1 | tail -8
} 2>&1 | tee $OUT/r2e_probes.log

#!/bin/bash
# round 2, GPU session V: the shorter pass (RB_X_PASS2: no per-pass liveness test, short divide behind a range test of
# the total, grid time in shared memory, ziggurat index/exponent tricks) against the current one, NVRTC kernels
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for knobs in "" "defs=RB_X_PASS2" "defs=RB_X_PASS2,tick=32"; do
  probe "$knobs" vilar 1250000 2 200 200 1
  probe "$knobs" dimers 1000000 2 1 1 1
  probe "$knobs" sir 1000000 2 250 250 0
  REBOP_B200_SCHEDULE=sparse probe "$knobs" sir 1000000 2 250 250 0
  probe "$knobs" mm_lma 1000000 2 100 100 0
done
} 2>&1 | tee $OUT/r2v_sweep.log
echo "== parity with RB_X_PASS2 (NVRTC kernels)"; REBOP_B200_CODEGEN="defs=RB_X_PASS2" timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_frontend.py -q -m gpu -k "nvrtc or frontend or Frontend or run" -x 2>&1 | tail -5 | tee -a $OUT/r2v_sweep.log

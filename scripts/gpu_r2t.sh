#!/bin/bash
# round 2, GPU session T: micro-variants of the sparse pass through NVRTC knobs (no liveness test, prescaled pick, tick)
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
for knobs in "" "defs=RB_X_NOLIVE" "prescale=1" "prescale=1,defs=RB_X_NOLIVE" "prescale=1,tick=32,defs=RB_X_NOLIVE" "tick=32"; do
  probe "$knobs" vilar 1250000 2 200 200 1
done
for knobs in "" "prescale=1,defs=RB_X_NOLIVE"; do
  probe "$knobs" dimers 1000000 2 1 1 1
  REBOP_B200_SCHEDULE=sparse probe "$knobs" mm_lma 1000000 2 100 100 0
done
} 2>&1 | tee $OUT/r2t_sweep.log
echo "== pytest -m gpu (default source changed: fire/cross predicates)"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== pytest parity with the experimental knobs"; REBOP_B200_CODEGEN="prescale=1,defs=RB_X_NOLIVE" timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nvrtc" 2>&1 | tail -3

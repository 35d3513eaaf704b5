#!/bin/bash
# round 2, GPU session AE: the last wave in well-filled warps (RB_TAIL_MIN lanes of a warp free before it claims)
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
for knobs in "" "defs=RB_TAIL_MIN:32" "defs=RB_TAIL_MIN:24" "defs=RB_TAIL_MIN:16" "defs=RB_TAIL_MIN:8"; do
  probe "$knobs" vilar 1250000 2 200 200 1
  probe "$knobs" dimers 1000000 2 1 1 1
  probe "$knobs" sir 1000000 2 250 250 0
  probe "$knobs" mm_lma 1000000 2 100 100 0
done
} 2>&1 | tee $OUT/r2ae_sweep.log
echo "== parity with RB_TAIL_MIN:24"; REBOP_B200_CODEGEN="defs=RB_TAIL_MIN:24" timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nvrtc" -x 2>&1 | tail -3 | tee -a $OUT/r2ae_sweep.log

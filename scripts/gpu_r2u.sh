#!/bin/bash
# round 2, GPU session U: pipe-balance variants of the sparse pass (NVRTC knobs): reaction choice through predicated
# IMADs, int32 stoichiometry rows updated by IMAD, divide without the range test / with the reciprocal issued early
OUT=gpurun_out
mkdir -p $OUT
probe() { local knobs="$1"; shift; echo "-- [$knobs] $*"; REBOP_B200_CODEGEN="$knobs" timeout 300 python scripts/perf_probe.py "$@" noprobe 2>&1 | tail -1; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for knobs in "" "imadsel=1" "widerows=1" "imadsel=1,widerows=1" "defs=RB_X_DIV:1" "defs=RB_X_DIV:2" "imadsel=1,widerows=1,defs=RB_X_DIV:1" "imadsel=1,widerows=1,defs=RB_X_DIV:2"; do
  probe "$knobs" vilar 1250000 2 200 200 1
done
for knobs in "" "imadsel=1,widerows=1" "imadsel=1,widerows=1,defs=RB_X_DIV:2"; do
  probe "$knobs" dimers 1000000 2 1 1 1
  probe "$knobs" sir 1000000 2 250 250 0
  REBOP_B200_SCHEDULE=sparse probe "$knobs" mm_lma 1000000 2 100 100 0
done
} 2>&1 | tee $OUT/r2u_sweep.log
echo "== parity with the experimental knobs"; REBOP_B200_CODEGEN="imadsel=1,widerows=1,defs=RB_X_DIV:2" timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nvrtc" -x 2>&1 | tail -3 | tee -a $OUT/r2u_sweep.log

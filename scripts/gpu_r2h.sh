#!/bin/bash
# round 2, GPU session H: ncu full captures of the three kernels (Vilar sparse prebuilt, SIR dense, synthetic PDM)
OUT=gpurun_out
mkdir -p $OUT
export REBOP_B200_JIT_DUMP=$PWD/$OUT
cap() {  # tag, kernel regex, probe args...
  local tag=$1; shift
  local rx=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $OUT/${tag} \
      python scripts/perf_probe.py "$@" noprobe > $OUT/${tag}.log 2>&1 ; echo "$tag rc=$?" ; tail -1 $OUT/${tag}.log
  ncu -i $OUT/${tag}.ncu-rep --page raw --csv > $OUT/${tag}_raw.csv 2>/dev/null
  ncu -i $OUT/${tag}.ncu-rep --page source --csv > $OUT/${tag}_src.csv 2>/dev/null
}
cap r2h_vilar_sparse rb_ssa_sys_Vilar_dyn vilar 606208 3 20 20 1
REBOP_B200_SCHEDULE=dense cap r2h_sir_dense "rb_ssa.*_dns" sir 1000000 0 250 250 0
cap r2h_synthetic_pdm "rb_ssa_jit_dyn" synthetic 100000 4 0.05 25 0
cap r2h_synthetic_exact "rb_ssa_jit_dyn" synthetic 100000 2 0.05 25 0
ls -la $OUT | grep r2h

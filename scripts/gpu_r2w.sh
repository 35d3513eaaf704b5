#!/bin/bash
# round 2, GPU session W: issue-throughput probe (which instruction kinds are half rate, which pipes overlap) and an
# ncu capture of the shorter pass
OUT=gpurun_out
mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/tput_probe scripts/probes/tput_probe.cu && /tmp/tput_probe 2>&1 | tee $OUT/r2w_tput_probe.log
REBOP_B200_CODEGEN="defs=RB_X_PASS2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_ssa_jit_dyn -c 1 -f -o $OUT/r2w_vilar_pass2 \
    python scripts/perf_probe.py vilar 606208 2 20 20 1 noprobe > $OUT/r2w_ncu.log 2>&1 ; echo "rc=$?" ; tail -1 $OUT/r2w_ncu.log | cut -c1-200
ncu -i $OUT/r2w_vilar_pass2.ncu-rep --page raw --csv > $OUT/r2w_vilar_pass2_raw.csv 2>/dev/null
ncu -i $OUT/r2w_vilar_pass2.ncu-rep --page source --csv > $OUT/r2w_vilar_pass2_src.csv 2>/dev/null
rm -f $OUT/r2w_vilar_pass2.ncu-rep

#!/bin/bash
# compute-sanitizer passes over small ensembles of every kernel family (run under gpurun).
OUT=${1:-gpurun_out/sanitizer.log}
: > $OUT
for tool in memcheck racecheck initcheck; do
  for cfg in "sir 3000 1 60 12 0" "sir 3000 2 60 12 0" "vilar 2000 0 2 4 1" "synthetic 1500 2 0.01 3 0" "dimers 2500 2 0.05 2 1"; do
    echo "== $tool: $cfg" >> $OUT
    timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/perf_probe.py $cfg noprobe >> $OUT 2>&1
    echo "rc=$?" >> $OUT
  done
done
grep -E "^== |rc=|ERROR SUMMARY|RACECHECK SUMMARY" $OUT

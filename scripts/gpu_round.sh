#!/bin/bash
# One GPU-box session: parity tests, smoke, the bench (both arms), launch list and a full ncu capture.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/${TAG}_smoke.log
echo "== bench reference" ; timeout 400 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ; echo "rc=$?" ; cat $OUT/${TAG}_bench_reference.json
echo "== bench" ; timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "rc=$?" ; cat $OUT/${TAG}_bench.json ; tail -3 $OUT/${TAG}_bench.err
echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1 ; echo "rc=$?"
echo "== ncu full capture of the SSA kernel (Vilar, dynamic schedule, 606208 trajectories, t=0..20)"
export REBOP_B200_JIT_DUMP=$PWD/$OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_ssa -c 1 -f -o $OUT/${TAG}_vilar_full \
    python scripts/perf_probe.py vilar 606208 0 20 20 1 noprobe > $OUT/${TAG}_ncu_full.log 2>&1 ; echo "rc=$?" ; tail -2 $OUT/${TAG}_ncu_full.log
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page raw --csv > $OUT/${TAG}_vilar_full_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page source --csv > $OUT/${TAG}_vilar_full_src.csv 2>/dev/null
unset REBOP_B200_JIT_DUMP
echo "== probes (auto kernel, auto schedule)"
for cfg in "sir 1000000 0 250 250 0" "dimers 1000000 0 1 1 1" "mm_lma 1000000 0 100 100 0" "vilar 1000000 0 200 200 1" "synthetic 300000 0 0.2 100 0"; do
  timeout 200 python scripts/perf_probe.py $cfg noprobe 2>&1 | tail -1
done | tee $OUT/${TAG}_probes.log

#!/bin/bash
# One GPU-box session: parity tests, smoke, the bench (both arms), launch list and a full ncu capture.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/${TAG}_smoke.log
echo "== bench reference" ; timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ; echo "rc=$?" ; cat $OUT/${TAG}_bench_reference.json
echo "== bench" ; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "rc=$?" ; cat $OUT/${TAG}_bench.json ; tail -3 $OUT/${TAG}_bench.err
echo "== ncu launch list (same command as the bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1 ; echo "rc=$?"
echo "== ncu full capture of the SSA kernel (Vilar, 303104 trajectories, t=0..20)"
export REBOP_B200_JIT_DUMP=$PWD
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rb_ssa -c 1 -f -o $OUT/${TAG}_vilar_full \
    python scripts/perf_probe.py vilar 303104 2 20 20 1 noprobe > $OUT/${TAG}_ncu_full.log 2>&1 ; echo "rc=$?" ; tail -3 $OUT/${TAG}_ncu_full.log
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page raw --csv > $OUT/${TAG}_vilar_full_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page source --csv > $OUT/${TAG}_vilar_full_src.csv 2>/dev/null
echo "== probes" ; for m in sir dimers vilar; do timeout 300 python scripts/perf_probe.py $m 1000000 2 ; done 2>&1 | tee $OUT/${TAG}_probes.log

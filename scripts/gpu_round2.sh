#!/bin/bash
# One GPU-box session of round 2: parity tests, smoke, the bench (both arms), the ncu launch list of the bench command
# and a full ncu capture of the headline kernel.   usage (under gpurun): bash scripts/gpu_round2.sh <tag>
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 ; echo "rc=$?" ; tail -4 $OUT/${TAG}_smoke.log
echo "== bench reference" ; timeout 400 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ; echo "rc=$?" ; cut -c1-300 $OUT/${TAG}_bench_reference.json
echo "== bench" ; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "rc=$?" ; cut -c1-600 $OUT/${TAG}_bench.json ; tail -3 $OUT/${TAG}_bench.err
echo "== ncu launch list (same command as the bench, fewer steps, headline part only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > $OUT/${TAG}_bench_under_ncu.log 2>&1 ; echo "rc=$?"
echo "== ncu full capture of the headline kernel (Vilar, prebuilt, sparse, 606208 trajectories, t=0..20)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_ssa_sys_Vilar_dyn -c 1 -f -o $OUT/${TAG}_vilar_full \
    python scripts/perf_probe.py vilar 606208 3 20 20 1 noprobe > $OUT/${TAG}_ncu_full.log 2>&1 ; echo "rc=$?" ; tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-200
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page raw --csv > $OUT/${TAG}_vilar_full_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_vilar_full.ncu-rep --page source --csv > $OUT/${TAG}_vilar_full_src.csv 2>/dev/null
echo "== python api probe (C3)"; timeout 300 python scripts/python_api_probe.py 1e7 2>&1 | tail -2 | tee $OUT/${TAG}_python_api.log

#!/bin/bash
# round 2, 8-GPU session (gpurun --gpus 8): one process, eight devices through rebop_ensemble_*, NCCL INFO log kept
N=${1:-8}
OUT=gpurun_out
TAG=r2l_${N}gpu
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
free -g | head -2
echo "== bench --single-process --gpus $N, full size, device-resident (NCCL_DEBUG=INFO)"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/${TAG}_nccl_%h_%p.log timeout 600 python bench.py --single-process --gpus $N --steps 2 --warmup 2 --no-e2e \
    > $OUT/${TAG}_bench_single_process.json 2> $OUT/${TAG}_bench_single_process.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench_single_process.json; tail -3 $OUT/${TAG}_bench_single_process.err
grep -h "Init COMPLETE\|NVLS\|Connected all\|Channel 00/" $OUT/${TAG}_nccl_*.log 2>/dev/null | head -40 > $OUT/${TAG}_nccl_summary.log; cut -c1-200 $OUT/${TAG}_nccl_summary.log | head -12
echo "== bench --single-process --gpus $N, 10^5 trajectories per GPU, host buffers + oracle check"
timeout 600 python bench.py --single-process --gpus $N --steps 2 --warmup 1 --traj-per-gpu 100000 \
    > $OUT/${TAG}_bench_single_process_e2e.json 2> $OUT/${TAG}_bench_single_process_e2e.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_single_process_e2e.json; tail -3 $OUT/${TAG}_bench_single_process_e2e.err
echo "== multi-device tests"
timeout 600 python -m pytest tests -q -m gpu -k "ensemble or devices or sharded" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log

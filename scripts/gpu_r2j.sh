#!/bin/bash
# round 2, multi-GPU session (run under gpurun --gpus N): the ensemble handle over N physical devices, NCCL all-reduce
# inside the library (NCCL_DEBUG=INFO log kept), bench.py in both multi-GPU modes
N=${1:-2}
OUT=gpurun_out
TAG=r2j_${N}gpu
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest (multi-device tests)"; timeout 600 python -m pytest tests -q -m gpu -k "ensemble or devices or sharded" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== bench --single-process --gpus $N (NCCL_DEBUG=INFO)"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/${TAG}_nccl_%h_%p.log timeout 900 python bench.py --single-process --gpus $N --steps 2 --warmup 2 \
    > $OUT/${TAG}_bench_single_process.json 2> $OUT/${TAG}_bench_single_process.err; echo "rc=$?"; cut -c1-700 $OUT/${TAG}_bench_single_process.json; tail -3 $OUT/${TAG}_bench_single_process.err
grep -h "Init COMPLETE\|nranks\|NVLS\|Connected all" $OUT/${TAG}_nccl_*.log 2>/dev/null | head -24 > $OUT/${TAG}_nccl_summary.log; cat $OUT/${TAG}_nccl_summary.log | cut -c1-220 | head -12
echo "== bench --gpus $N (torchrun, one process per GPU)"
timeout 900 python bench.py --gpus $N --steps 2 --warmup 3 --no-configs > $OUT/${TAG}_bench_torchrun.json 2> $OUT/${TAG}_bench_torchrun.err; echo "rc=$?"; cut -c1-500 $OUT/${TAG}_bench_torchrun.json; tail -3 $OUT/${TAG}_bench_torchrun.err

#!/bin/bash
# round 2, final multi-GPU sanity of the final kernels (run under gpurun --gpus 2): multi-device tests, torchrun bench
N=${1:-2}
OUT=gpurun_out
TAG=r2ah_${N}gpu
mkdir -p $OUT
echo "== pytest (multi-device tests)"; timeout 300 python -m pytest tests -q -m gpu -k "ensemble or devices or sharded" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
echo "== bench --gpus $N (torchrun, one process per GPU)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 --no-configs > $OUT/${TAG}_bench_torchrun.json 2> $OUT/${TAG}_bench_torchrun.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench_torchrun.json; tail -3 $OUT/${TAG}_bench_torchrun.err
echo "== bench --single-process --gpus $N"
timeout 600 python bench.py --single-process --gpus $N --steps 2 --warmup 3 --no-configs > $OUT/${TAG}_bench_single_process.json 2> $OUT/${TAG}_bench_single_process.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench_single_process.json; tail -3 $OUT/${TAG}_bench_single_process.err

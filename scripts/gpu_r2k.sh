#!/bin/bash
# multi-device tests only (run under gpurun --gpus 2)
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu -k "ensemble or devices or sharded" > $OUT/r2k_2gpu_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2k_2gpu_pytest.log

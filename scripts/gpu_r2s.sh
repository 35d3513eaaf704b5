#!/bin/bash
# round 2: the driver's multi-GPU bench command with every config (run under gpurun --gpus 2)
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 \
   > $OUT/r2s_${N}gpu_bench.json 2> $OUT/r2s_${N}gpu_bench.err; echo "rc=$?"; tail -4 $OUT/r2s_${N}gpu_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2s_${N}gpu_bench.json') if l.startswith('{')][-1])
print("N=%d value %.4g e2e %.4g parity %s"%(d['n_gpus'], d['value'], d['e2e']['value'], d['parity_check']))
for k,c in d['configs'].items():
    print(k, "value %.4g"%c['value'], c['unit'], "parity", c['parity_check']['equal'], c.get('fast_kernel',{}).get('value'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | grep '^{' | cut -c1-200

#!/bin/bash
# round 2, GPU session Q: new tests (Hill rate, Python fast kernel), PDM with ordered blocks, bench configs with the fast-kernel leg
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest new"; timeout 900 python -m pytest tests -q -m gpu -k "hill or pdm or fast" 2>&1 | tail -4
echo "== pdm synthetic"; timeout 600 python scripts/pdm_probe.py synthetic 300000 2>&1 | tail -7 | head -3 | tee $OUT/r2q_pdm.log
echo "== bench (2 steps)"; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/r2q_bench.json 2> $OUT/r2q_bench.err; echo "rc=$?"; tail -3 $OUT/r2q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench.json'))
c=d['configs']['C5_synthetic_api']
print("C5 exact %.4g fast %s"%(c['value'], json.dumps(c.get('fast_kernel'))[:500]))
print("headline %.4g e2e %.4g"%(d['value'], d['e2e']['value']))
PY

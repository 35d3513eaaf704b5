#!/bin/bash
# round 2, GPU session M: staged copies into pageable host memory (C3 through the Python binding), host-buffer tests
OUT=gpurun_out
mkdir -p $OUT
{
for st in 0 1; do
  echo "== REBOP_B200_STAGED_COPY=$st python api probe 1e7"
  REBOP_B200_STAGED_COPY=$st timeout 600 python scripts/python_api_probe.py 1e7 2>&1 | tail -3
done
} 2>&1 | tee $OUT/r2m_python_api.log
echo "== pytest host-buffer tests"; timeout 900 python -m pytest tests -q -m gpu -k "host or sample_types or frontend or segmented or ensemble or michaelis" 2>&1 | tail -4

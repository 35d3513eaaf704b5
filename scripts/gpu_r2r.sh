#!/bin/bash
# round 2, GPU session R: ncu full capture of the sample-finishing kernel K6 (SIR, 10^6 trajectories, int32 and int16)
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/k6_probe.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from rebop_b200 import _ffi, models
m = models.sir(); net = models.build_network(m)
b = _ffi.Batch(net, 1000000, m["x0"], seeds=None, seed_base=0)
b.set_sample_dtype(np.int16 if sys.argv[1] == "16" else np.int32)
for i in range(2):
    b.set_species(m["x0"]); b.set_time(0.0); b.seed(None, 0); b.run_grid(250.0, 250)
print("loop", b.last_kernel_ms, "finish", b.last_finish_ms)
PY
for w in 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_samples_finish -s 1 -c 1 -f -o $OUT/r2r_k6_i$w python /tmp/k6_probe.py $w > $OUT/r2r_k6_i$w.log 2>&1; echo "rc=$?"; tail -1 $OUT/r2r_k6_i$w.log
  ncu -i $OUT/r2r_k6_i$w.ncu-rep --page raw --csv > $OUT/r2r_k6_i${w}_raw.csv 2>/dev/null
  ncu -i $OUT/r2r_k6_i$w.ncu-rep --page source --csv > $OUT/r2r_k6_i${w}_src.csv 2>/dev/null
done

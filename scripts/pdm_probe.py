"""Dependency-driven kernel (REBOP_KERNEL_PDM) against the bit-exact kernels: agreement and throughput (development aid)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from rebop_b200 import _ffi, models

name = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100000
m = models.MODELS[name]()
tmax = float(sys.argv[3]) if len(sys.argv) > 3 else m["tmax"]
nb = int(sys.argv[4]) if len(sys.argv) > 4 else m["nb_steps"]
net = models.build_network(m, 0)
save = list(range(min(10, len(m["species"]))))
res = {}
for kernel in (_ffi.KERNEL_AUTO, _ffi.KERNEL_PDM):
    b = _ffi.Batch(net, n, m["x0"], seeds=None, seed_base=0, kernel=kernel)
    for rep in range(2):
        b.set_species(m["x0"]); b.set_time(0.0); b.seed(None, 0)
        t0 = time.time()
        b.run_grid(tmax, nb, save_idx=save)
        wall = time.time() - t0
    ev, ms = b.events()[1], b.last_kernel_ms
    out = b.samples()
    res[kernel] = out
    print(f"{name} kernel={b.kernel_used} sched={b.schedule_used} n={n}: events={ev:.4g} loop={ms:.2f} ms wall={wall*1e3:.1f} -> "
          f"{ev / ms * 1e3:.4g} events/s, lane eff {ev / max(b.lane_slots, 1):.3f}", flush=True)
    b.close()
a, p = res[_ffi.KERNEL_AUTO].astype(np.float64), res[_ffi.KERNEL_PDM].astype(np.float64)
same = (a == p).all(axis=(0, 1)).mean()
print(f"trajectories identical at every sample: {same:.4f}")
ma, mp = a[-1].mean(axis=1), p[-1].mean(axis=1)
se = a[-1].std(axis=1) / np.sqrt(n)
print("mean(last row) exact:", np.round(ma, 3))
print("mean(last row) pdm  :", np.round(mp, 3))
print("max |diff| / stderr :", np.max(np.abs(ma - mp) / np.maximum(se, 1e-12)))

"""Quick throughput probe (development aid; the judged numbers come from bench.py)."""
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


from rebop_b200 import _ffi, models

name = sys.argv[1] if len(sys.argv) > 1 else "vilar"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 148 * 2048
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tmax = float(sys.argv[4]) if len(sys.argv) > 4 else None
nb = int(sys.argv[5]) if len(sys.argv) > 5 else None
arith = int(sys.argv[6]) if len(sys.argv) > 6 else 0
mkw = {}
if ":" in name:  # e.g. ring:n=30
    name, kws = name.split(":", 1)
    mkw = {k: int(v) for k, v in (kv.split("=") for kv in kws.split(","))}
m = models.MODELS[name](**mkw)
tmax = m["tmax"] if tmax is None else tmax
nb = m["nb_steps"] if nb is None else nb
net = models.build_network(m, arith)
for rep in range(2):
    b = _ffi.Batch(net, n, m["x0"], seeds=None, seed_base=0, kernel=kernel)
    t0 = time.time()
    b.run_grid(tmax, nb)
    wall = time.time() - t0
    ev = b.events()[1]
    ms = b.last_kernel_ms
    fin = b.last_finish_ms
    print(f"{name} n={n} kernel={b.kernel_used} sched={b.schedule_used} arith={arith} tmax={tmax} nb={nb}: events={ev:.4g} "
          f"({ev / n:.1f}/traj) loop={ms:.2f} ms finish={fin:.2f} ms wall={wall * 1e3:.1f} ms -> {ev / ms * 1e3:.4g} events/s "
          f"(loop), {ev / (ms + fin) * 1e3:.4g} (loop+finish), {n / (ms + fin) * 1e3:.4g} traj/s, "
          f"lane eff {ev / max(b.lane_slots, 1):.3f}", flush=True)
    b.close()
if len(sys.argv) <= 7:
    ops, mhz = _ffi.measure_fp64_rate(0)
    print(f"fp64 non-fused issue rate {ops:.4g} op/s at {mhz:.0f} MHz")

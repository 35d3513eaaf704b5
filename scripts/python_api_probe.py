"""End-to-end probe of the Python front end (development aid): examples/mm.py of the reference, batched."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rebop_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
mm = rebop_b200.Gillespie()
mm.add_reaction("V * A / (Km + A)", ["A"], ["P"])
for rep in range(2):
    t0 = time.perf_counter()
    ds = mm.run({"A": 100}, tmax=250, nb_steps=100, params={"V": 1, "Km": 20}, rng=0, n_trajectories=n, dtype=np.int16)
    dt = time.perf_counter() - t0
    print(f"mm expr n={n}: {dt * 1e3:.1f} ms wall, kernel {mm.last_kernel_ms:.2f} ms, {mm.last_events} events -> "
          f"{n / dt:.4g} traj/s end to end, {mm.last_events / mm.last_kernel_ms * 1e3:.4g} events/s in the kernel; "
          f"mean P(tmax) = {float(np.mean(ds.P[-1])):.3f}", flush=True)
t0 = time.perf_counter()
ds = mm.run({"A": 100}, tmax=250, nb_steps=100, params={"V": 1, "Km": 20}, rng=0, n_trajectories=n, reduce=True)
print(f"reduce=True: {(time.perf_counter() - t0) * 1e3:.1f} ms wall; P_mean(tmax) = {float(ds.P_mean[-1]):.3f}")

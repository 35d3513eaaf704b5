"""Pinned-memory D2H bandwidth of this box, and the e2e overhead of run_grid(host_out) (development aid)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rebop_b200 import _ffi, models

n_bytes = 4 << 30
dev = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"torch pinned D2H: {n_bytes / dt / 1e9:.1f} GB/s")
del dev, host
m = models.vilar()
n = 1_250_000
net = models.build_network(m, 1)
b = _ffi.Batch(net, n, m["x0"], seeds=None, seed_base=0)
out = _ffi.PinnedBuffer((21, 9, n), np.int32)
for rep in range(3):
    b.set_species(m["x0"]); b.set_time(0.0); b.seed(None, rep * n)
    t0 = time.perf_counter()
    b.run_grid(20.0, 20, host_out=out.array)
    dt = time.perf_counter() - t0
    print(f"run_grid(host_out) t=20: wall {dt * 1e3:.1f} ms, kernel {b.last_kernel_ms:.1f} ms, copy+overhead {dt * 1e3 - b.last_kernel_ms:.1f} ms "
          f"for {out.array.nbytes / 1e9:.2f} GB -> {out.array.nbytes / max(dt - b.last_kernel_ms * 1e-3, 1e-9) / 1e9:.1f} GB/s")
out2 = _ffi.PinnedBuffer((201, 9, n), np.int32)
t0 = time.perf_counter()
b.set_species(m["x0"]); b.set_time(0.0); b.seed(None, 7 * n)
b.run_grid(2.0, 200, host_out=out2.array)
dt = time.perf_counter() - t0
print(f"run_grid(host_out) 201 rows: wall {dt * 1e3:.1f} ms, kernel {b.last_kernel_ms:.1f} ms, {out2.array.nbytes / 1e9:.2f} GB -> "
      f"{out2.array.nbytes / max(dt - b.last_kernel_ms * 1e-3, 1e-9) / 1e9:.1f} GB/s")

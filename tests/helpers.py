"""Shared helpers: the same model dict goes to the oracle and to the product."""
import numpy as np

from rebop_b200 import models


def oracle_network(O, model, arith=0, dense=False):
    rx = [("lma", k, terms, diff) for k, terms, diff in model["reactions"]]
    return O.Network(len(model["species"]), rx, arith=arith, dense=dense)


def run_product(ffi, model, seeds, tmax, nb_steps, kernel, arith=0, save_idx=None, x0=None):
    net = models.build_network(model, arith)
    b = ffi.Batch(net, len(seeds), model["x0"] if x0 is None else x0, seeds=seeds, kernel=kernel)
    b.run_grid(tmax, nb_steps, save_idx=save_idx)
    out = b.samples()
    ev = b.events()[0]
    used = b.kernel_used
    b.close()
    return out, ev, used


def numpy_seeds(n, rng=0):
    """Per-trajectory seeds the way python/rebop/gillespie.py:139-140 derives one per run."""
    return np.random.default_rng(rng).integers(np.iinfo(np.uint64).max, size=n, dtype=np.uint64)

#!/usr/bin/env python3
"""Writes tests/golden/*.json from the CPU oracle.

The Rust reference cannot be built or imported in this image (no cargo/rustc/maturin), so the
fixtures cannot come from the reference itself.  They come from the oracle AFTER it has been pinned
to the reference's own known answers (tests/test_oracle.py: test_fixed_seed golden vector,
rate_lma table, test_eval), and exist to (a) guard the oracle against regressions and (b) give the
GPU tests full-length final states to compare with (tests/test_gpu_parity.py::test_golden_fixtures).
The first fixture embeds the reference's golden vector itself (rng=42 => 0, 227, 773).

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from rebop_b200 import models  # noqa: E402
from tests.helpers import numpy_seeds, oracle_network  # noqa: E402

CASES = [
    # name, model, arith, seeds, tmax, nb_steps
    ("sir_rng42", "sir", 0, [int(np.random.default_rng(42).integers(np.iinfo(np.uint64).max, dtype=np.uint64))], 250.0, 250),
    ("sir_api", "sir", 0, numpy_seeds(64, rng=0).tolist(), 250.0, 250),
    ("dimers_macro", "dimers", 1, list(range(32)), 1.0, 1),
    ("mm_lma_api", "mm_lma", 0, numpy_seeds(48, rng=5).tolist(), 100.0, 100),
    ("vilar_macro_full", "vilar", 1, list(range(16)), 200.0, 200),
    ("vilar_api_full", "vilar", 0, list(range(1000, 1008)), 200.0, 200),
    ("synthetic_api", "synthetic", 0, list(range(24)), 0.05, 10),  # BASELINE config C5: 100 species, 500 reactions
]

for name, mname, arith, seeds, tmax, nb in CASES:
    model = models.MODELS[mname]()
    out, ev, tot = oracle_network(O, model, arith).run_batch(model["x0"], np.array(seeds, dtype=np.uint64), tmax, nb, threads=8)
    doc = dict(model=mname, arith=arith, seeds=[int(s) for s in seeds], tmax=tmax, nb_steps=nb,
               final=out[-1].T.tolist(), events=[int(e) for e in ev], checksum=int(out.astype(np.int64).sum()))
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(doc, fh)
    print(name, "events", tot, "final[0]", doc["final"][0])

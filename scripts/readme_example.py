"""The README's usage example, executable (run on a GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rebop_b200

sir = rebop_b200.Gillespie()
sir.add_reaction(1e-4, ["S", "I"], ["I", "I"])
sir.add_reaction(0.01, ["I"], ["R"])
ds = sir.run({"S": 999, "I": 1}, tmax=250, nb_steps=250, rng=42, n_trajectories=100_000)
print("ds.S", np.asarray(ds.S).shape, "first trajectory final:", int(ds.S[-1][0]), int(ds.I[-1][0]), int(ds.R[-1][0]))
assert (int(ds.S[-1][0]), int(ds.I[-1][0]), int(ds.R[-1][0])) == (0, 227, 773)  # trajectory 0 = the reference's rng=42 run

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Vilar = rebop_b200.define_system(open(os.path.join(root, "rebop_b200", "systems", "vilar.rsys"), encoding="utf-8").read())
v = Vilar.with_parameters(50., 500., .01, 50., 50., 5., 10., .5, 1., .2, 1., 1., 2., 50., 100., n_trajectories=100_000)
v.Da = 1
v.Dr = 1
v.seed(0)
v.advance_until(20.)
print("v.A", v.A.shape, "mean A(20) =", float(v.A.mean()), "events", v.events, "kernel", v.kernel_used)
events = sir.run({"S": 999, "I": 1}, tmax=250, nb_steps=0, rng=42)
print("event log rows:", len(events.time), "last time", float(events.time[-1]))

"""The partial-propensity kernel (REBOP_KERNEL_PDM): opt-in, statistically exact, not stream-exact.

It fills the reference's promise of "heuristics" for large systems (python/rebop/gillespie.py:123-126, absent from
src/pyo3_gillespie.rs:161).  Its floating-point sums are ordered and fused differently from the reference's running
sum (src/gillespie.rs:357-364), so parity is tier 2 of the north star: the GPU ensemble against oracle runs with
INDEPENDENT seeds -- two-sample KS at family-wise alpha = 1e-3 (Bonferroni over sample times x species), means and
variances within 5 standard errors (tests/stats_helpers.py).  The CPU part checks the lowering and that the kernel
cross-compiles for sm_100a within its register and shared-memory budget.
"""
import re
import subprocess

import numpy as np
import pytest

from rebop_b200 import models
from tests.helpers import oracle_network
from tests.stats_helpers import compare_ensembles

ALPHA, Z = 1e-3, 5.0


def test_lowering_factors_the_network_by_owner_species(ffi):
    m = models.synthetic()
    src = models.build_network(m).codegen_pdm()
    head = src.splitlines()[0]
    groups, consts = (int(v) for v in re.search(r"(\d+) owner groups, (\d+) constants", head).groups())
    owners = {terms[0][0] for _, terms, _ in m["reactions"] if terms}
    pairs = {(terms[0][0], terms[-1][0]) for _, terms, _ in m["reactions"] if sum(e for _, e in terms) == 2}
    assert groups == len(owners)
    assert consts == len(owners) + len(pairs)          # one c_i per owner, one K_ij per distinct reactant pair
    assert src.count("fma(") >= consts                  # one fused multiply-add per constant in the unrolled pass
    assert "rb_pp_select" in src and "__dmul_rn" not in src


def test_lowering_rejects_what_the_form_cannot_express(ffi):
    net = ffi.Network(2)
    net.add_reaction_lma_sparse(1.0, [(0, 3)], [-3, 1])                 # third order
    with pytest.raises(ffi.RebopError, match="total order <= 2"):
        net.codegen_pdm()
    net = ffi.Network(2)
    net.add_reaction_lma_sparse(1.0, [(0, 1)], [-2, 1])                 # consumes more than its order: counts could go negative
    with pytest.raises(ffi.RebopError, match="negative"):
        net.codegen_pdm()
    net = ffi.Network(1)
    net.add_reaction_lma_sparse(-1.0, [(0, 1)], [-1])
    with pytest.raises(ffi.RebopError, match="rate constants >= 0"):
        net.codegen_pdm()
    net = ffi.Network(1)
    net.add_reaction_expr([("species", 0, 0.0)], [-1])
    with pytest.raises(ffi.RebopError, match="expression rate"):
        net.codegen_pdm()


def test_kernel_cross_compiles_within_budget(ffi, tmp_path):
    """NVRTC needs no GPU: two CTAs of 128 threads per SM must fit (registers, shared memory), no spills."""
    cubin = tmp_path / "pdm.cubin"
    cubin.write_bytes(models.build_network(models.synthetic()).jit_cubin_pdm())
    usage = subprocess.check_output(["cuobjdump", "-res-usage", str(cubin)], text=True)
    regs = [int(v) for v in re.findall(r"REG:(\d+)", usage)]
    stack = [int(v) for v in re.findall(r"STACK:(\d+)", usage)]
    assert len(regs) == 3 and max(regs) <= 255 and max(stack) == 0, usage
    sass = subprocess.check_output(["cuobjdump", "-sass", str(cubin)], text=True)
    assert sass.count("DFMA") > 240 and "STL" not in sass


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,tmax,nb_steps", [
    ("synthetic", 3000, 0.05, 5),    # BASELINE config C5 (100 species x 500 reactions), 5 400 events per trajectory
    ("vilar", 1500, 20.0, 10),
    ("sir", 20000, 250.0, 10),
    ("dimers", 3000, 1.0, 4),
])
def test_pdm_ensemble_matches_cpu_law(gpu, ffi, oracle, name, n, tmax, nb_steps):
    m = models.MODELS[name]()
    ref, _, _ = oracle_network(oracle, m, 0).run_batch(m["x0"], models.seeds_sequence(n, 5 * 10**8), tmax, nb_steps, threads=16)
    b = ffi.Batch(models.build_network(m, 0), n, m["x0"], seeds=None, seed_base=0, kernel=ffi.KERNEL_PDM)
    b.run_grid(tmax, nb_steps)
    assert b.kernel_used == ffi.KERNEL_PDM
    out = b.samples()
    b.close()
    fails = compare_ensembles(out, ref, ALPHA, Z)
    assert fails == [], "\n".join(fails[:10])


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [1, 2, 3])
def test_pdm_conserves_mass_and_counts_events(gpu, ffi, schedule):
    """Size-independent properties: the synthetic network conserves sum(m_s x_s); counts stay >= 0; results do not
    depend on the schedule (every trajectory owns its state and random stream)."""
    m = models.synthetic()
    n = 5000
    mass = np.array([1 + s % 3 for s in range(len(m["species"]))], dtype=np.int64)
    b = ffi.Batch(models.build_network(m, 0), n, m["x0"], seeds=None, seed_base=7, kernel=ffi.KERNEL_PDM)
    b.set_schedule(schedule)
    b.run_grid(0.02, 4)
    out = b.samples().astype(np.int64)
    ev = b.events()[1]
    b.close()
    assert (out >= 0).all() and ev > 0
    total = (out * mass[None, :, None]).sum(axis=1)
    assert (total == total[0, 0]).all()
    if schedule == 1:
        test_pdm_conserves_mass_and_counts_events.first = out
    else:
        np.testing.assert_array_equal(out, test_pdm_conserves_mass_and_counts_events.first)


@pytest.mark.gpu
def test_pdm_is_opt_in_and_refuses_event_logs(gpu, ffi):
    m = models.sir()
    b = ffi.Batch(models.build_network(m), 64, m["x0"], seeds=None)
    b.run_grid(10.0, 2)
    assert b.kernel_used != ffi.KERNEL_PDM          # AUTO never picks it
    b.close()
    b = ffi.Batch(models.build_network(m), 64, m["x0"], seeds=None, kernel=ffi.KERNEL_PDM)
    with pytest.raises(ffi.RebopError, match="event-log"):
        b.run_events(1.0)
    b.close()


@pytest.mark.gpu
def test_python_front_end_fast_kernel(gpu, oracle):
    """`Gillespie.run(..., kernel="fast")`: the binding's switch for the partial-propensity kernel; same law as the
    default kernel (tier 2), and refused with the reason for a network the form cannot express."""
    import rebop_b200

    sir = rebop_b200.Gillespie()
    sir.add_reaction(1e-4, ["S", "I"], ["I", "I"])
    sir.add_reaction(0.01, ["I"], ["R"])
    n = 20000
    fast = sir.run({"S": 999, "I": 1}, tmax=250, nb_steps=10, rng=1, n_trajectories=n, kernel="fast", dtype=np.int32)
    exact = sir.run({"S": 999, "I": 1}, tmax=250, nb_steps=10, rng=2, n_trajectories=n, dtype=np.int32)
    a = np.stack([np.asarray(fast[v]) for v in ("S", "I", "R")], axis=1)
    b = np.stack([np.asarray(exact[v]) for v in ("S", "I", "R")], axis=1)
    assert compare_ensembles(a, b, ALPHA, Z) == []
    mm = rebop_b200.Gillespie()
    mm.add_reaction("V * A / (Km + A)", ["A"], ["P"])
    with pytest.raises(Exception, match="expression rate"):
        mm.run({"A": 100}, tmax=1, nb_steps=1, params={"V": 1, "Km": 20}, n_trajectories=8, kernel="fast")

"""Tier-1 parity: the CUDA path must reproduce the oracle's integer trajectories bit for bit.

Every test calls through the C ABI (rebop_b200/_ffi.py) and compares with the CPU oracle on
the same seeds.  Sizes are chosen so that the oracle finishes in seconds.
"""
import numpy as np
import pytest

from rebop_b200 import models
from tests.helpers import numpy_seeds, oracle_network, run_product

pytestmark = pytest.mark.gpu

KERNELS = {"table": 1, "nvrtc": 2}


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("name,n,tmax,nb_steps", [
    ("sir", 4096, 250.0, 250),
    ("dimers", 1024, 1.0, 4),
    ("mm_lma", 2048, 100.0, 100),
    ("vilar", 96, 20.0, 20),
])
def test_bit_exact_vs_oracle(gpu, ffi, oracle, kernel, arith, name, n, tmax, nb_steps):
    model = models.MODELS[name]()
    seeds = numpy_seeds(n, rng=7)
    ref, ref_ev, ref_tot = oracle_network(oracle, model, arith).run_batch(model["x0"], seeds, tmax, nb_steps, threads=8)
    out, ev, used = run_product(ffi, model, seeds, tmax, nb_steps, KERNELS[kernel], arith)
    assert used == KERNELS[kernel]
    assert out.shape == ref.shape
    np.testing.assert_array_equal(out, ref)
    assert ev == ref_tot


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("name,kwargs,n,tmax,nb_steps", [
    ("synthetic", {}, 160, 0.02, 5),                 # BASELINE config C5: 100 species, 500 reactions
    ("flocculation", {"n": 50}, 96, 0.02, 4),        # my_benchmark.rs:683-700: 50 species, 625 reactions
    ("ring", {"n": 50}, 96, 2.0, 4),                 # my_benchmark.rs:601-614: 50 species, 50 reactions
    ("ring", {"n": 300, "a0": 200}, 40, 1.0, 2),     # 300 species: 32-thread CTAs
    ("ring", {"n": 40}, 96, 2.0, 4),                 # 40 species + 40 reactions: still register-resident (2 CTAs/SM)
    ("flocculation", {"n": 16}, 96, 0.02, 4),        # 16 species, 64 reactions: register-resident
])
def test_large_networks_bit_exact(gpu, ffi, oracle, arith, name, kwargs, n, tmax, nb_steps):
    """Networks beyond the register-resident limits: the shared-memory specialised form (NVRTC) and the
    table-driven kernel against the oracle."""
    model = models.MODELS[name](**kwargs)
    seeds = numpy_seeds(n, rng=11)
    ref, _, ref_tot = oracle_network(oracle, model, arith).run_batch(model["x0"], seeds, tmax, nb_steps, threads=8)
    assert ref_tot > 50 * n
    for kernel in ("nvrtc", "table"):
        out, ev, used = run_product(ffi, model, seeds, tmax, nb_steps, KERNELS[kernel], arith)
        assert used == KERNELS[kernel]
        np.testing.assert_array_equal(out, ref)
        assert ev == ref_tot
    # a subset of the species, in the middle of the index range
    save = sorted({3, len(model["species"]) // 2 + 1, len(model["species"]) - 1})
    out, _, _ = run_product(ffi, model, seeds, tmax, nb_steps, KERNELS["nvrtc"], arith, save_idx=save)
    np.testing.assert_array_equal(out, ref[:, save, :])


def test_large_form_falls_back_when_counts_can_be_negative(gpu, ffi, oracle):
    """A negative initial count breaks the monotone cumulative sums the large form relies on: AUTO falls back
    to the table-driven kernel, an explicit NVRTC request is refused."""
    model = models.ring(n=50)
    x0 = list(model["x0"])
    x0[7] = -3
    seeds = numpy_seeds(32, rng=3)
    ref, _, tot = oracle_network(oracle, model).run_batch(x0, seeds, 0.5, 2)
    out, ev, used = run_product(ffi, model, seeds, 0.5, 2, 0, x0=x0)
    assert used == KERNELS["table"]
    np.testing.assert_array_equal(out, ref)
    with pytest.raises(ffi.RebopError) as e:
        run_product(ffi, model, seeds, 0.5, 2, KERNELS["nvrtc"], x0=x0)
    assert e.value.status == ffi.ERR_LIMIT


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 1000])
def test_ragged_sizes(gpu, ffi, oracle, kernel, n):
    model = models.sir()
    seeds = models.seeds_sequence(n, first=1000)
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 250.0, 50)
    out, ev, _ = run_product(ffi, model, seeds, 250.0, 50, KERNELS[kernel])
    np.testing.assert_array_equal(out, ref)
    assert ev == tot


def test_reference_golden_vector(gpu, ffi):
    """tests/test_rebop.py:30-36: rng=42 => S=0, I=227, R=773 at t=250."""
    seed = np.random.default_rng(42).integers(np.iinfo(np.uint64).max, dtype=np.uint64)
    for kernel in KERNELS.values():
        out, ev, _ = run_product(ffi, models.sir(), np.array([seed], dtype=np.uint64), 250.0, 250, kernel)
        assert out[-1, :, 0].tolist() == [0, 227, 773]
        assert out[0, :, 0].tolist() == [999, 1, 0]
        assert ev == 1772


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_seed_sequence_matches_explicit_seeds(gpu, ffi, kernel):
    model = models.sir()
    net = models.build_network(model)
    n = 500
    a = ffi.Batch(net, n, model["x0"], seeds=models.seeds_sequence(n, 12345), kernel=KERNELS[kernel])
    b = ffi.Batch(net, n, model["x0"], seeds=None, seed_base=12345, kernel=KERNELS[kernel])
    a.run_grid(100.0, 10)
    b.run_grid(100.0, 10)
    np.testing.assert_array_equal(a.samples(), b.samples())


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_repeated_advance_until_continues_streams(gpu, ffi, oracle, kernel):
    """Gillespie keeps its RNG between advance_until calls (src/lib.rs:129-133)."""
    model = models.dimers()
    n = 256
    seeds = models.seeds_sequence(n)
    net = models.build_network(model)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=KERNELS[kernel])
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 1.0, 4)
    for i in range(5):
        b.advance_until(1.0 * i / 4)
        np.testing.assert_array_equal(b.species().T, ref[i])
        np.testing.assert_array_equal(b.times(), np.full(n, 1.0 * i / 4))
    assert b.events()[0] == tot


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_var_names_subset(gpu, ffi, kernel):
    """tests/test_rebop.py:68-89: saving a subset never changes the dynamics."""
    model = models.sir()
    seeds = models.seeds_sequence(300)
    full, _, _ = run_product(ffi, model, seeds, 250.0, 25, KERNELS[kernel])
    sub, _, _ = run_product(ffi, model, seeds, 250.0, 25, KERNELS[kernel], save_idx=[0, 2])
    np.testing.assert_array_equal(sub, full[:, [0, 2], :])


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_conservation_and_bounds(gpu, ffi, kernel):
    """src/gillespie.rs:504-530, tests/test_rebop.py:15-27."""
    model = models.sir()
    out, _, _ = run_product(ffi, model, models.seeds_sequence(2000), 250.0, 250, KERNELS[kernel])
    assert (out.sum(axis=1) == 1000).all()
    assert (out >= 0).all() and (out[:, 0] <= 999).all()
    model = models.dimers()
    out, _, _ = run_product(ffi, model, models.seeds_sequence(200), 1.0, 1, KERNELS[kernel])
    assert (out[-1, 0] == 1).all() and (out[-1, 2] > 1000).all() and (out[-1, 3] < 10000).all()


@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_nan_rate_freezes(gpu, ffi, kernel, schedule):
    """src/gillespie_macro.rs:224-238: a NaN rate constant => no reaction, t = tmax."""
    net = ffi.Network(1, 1)
    net.add_reaction_lma_sparse(10.0, [], [1])
    net.add_reaction_lma_sparse(float("nan"), [(0, 1)], [-1])
    b = ffi.Batch(net, 64, [0], seeds=models.seeds_sequence(64), kernel=KERNELS[kernel])
    b.set_schedule(schedule)
    b.advance_until(100.0)
    assert (b.species() == 0).all()
    assert (b.times() == 100.0).all()
    assert b.events()[0] == 0


@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_no_reactions(gpu, ffi, kernel, schedule):
    """src/gillespie_macro.rs:239-253."""
    net = ffi.Network(3, 0)
    b = ffi.Batch(net, 40, [42, 1337, 0], seeds=models.seeds_sequence(40), kernel=KERNELS[kernel])
    b.set_schedule(schedule)
    b.run_grid(1e20, 3)
    out = b.samples()
    assert (out[:, 0] == 42).all() and (out[:, 1] == 1337).all() and (out[:, 2] == 0).all()
    assert (b.times() == 1e20).all()


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_sample_sums(gpu, ffi, kernel):
    model = models.sir()
    n = 3001
    net = models.build_network(model)
    b = ffi.Batch(net, n, model["x0"], seeds=models.seeds_sequence(n), kernel=KERNELS[kernel])
    b.run_grid(250.0, 20)
    out = b.samples().astype(np.int64)
    s1, s2 = b.sample_sums()
    np.testing.assert_array_equal(s1.reshape(21, 3), out.sum(axis=2))
    np.testing.assert_array_equal(s2.reshape(21, 3).astype(np.int64), (out * out).sum(axis=2))


@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_iteration_cap_is_reported(gpu, ffi, oracle, kernel, schedule):
    """The watchdog stops a launch between two passes; advance_until can then simply be called again and ends
    where an uninterrupted run ends (state, time and streams are written back)."""
    model = models.dimers()
    net = models.build_network(model)
    n = 64
    seeds = models.seeds_sequence(n)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=KERNELS[kernel])
    b.set_schedule(schedule)
    b.set_max_iters(64)
    calls = 0
    while True:
        calls += 1
        try:
            b.advance_until(0.05)
            break
        except ffi.RebopError as e:
            assert e.status == ffi.ERR_ITER_CAP
            assert calls < 200
    assert calls > 2
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 0.05, 0)
    np.testing.assert_array_equal(b.species().T, ref[-1])
    assert b.events()[0] == tot
    assert (b.times() == 0.05).all()


@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("kernel", ["table", "nvrtc", "prebuilt"])
def test_golden_fixtures(gpu, ffi, kernel, schedule):
    """Committed fixtures (tests/golden/, full-length runs incl. Vilar to t=200 and the reference's
    rng=42 vector): final states, per-run event totals and a checksum over every sample -- for every kernel
    family (the build-time kernels wherever one matches the fixture's network) in both schedules."""
    import json
    import os

    gdir = os.path.join(os.path.dirname(__file__), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".json"))
    assert files
    ran = 0
    for f in files:
        g = json.load(open(os.path.join(gdir, f)))
        model = models.MODELS[g["model"]]()
        net = models.build_network(model, g["arith"])
        if kernel == "prebuilt" and not net.has_prebuilt:
            continue
        seeds = np.array(g["seeds"], dtype=np.uint64)
        b = ffi.Batch(net, len(seeds), model["x0"], seeds=seeds, kernel={"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel])
        b.set_schedule(schedule)
        b.run_grid(g["tmax"], g["nb_steps"])
        out, ev = b.samples(), b.events()[0]
        assert b.schedule_used == schedule and b.kernel_used == {"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel]
        b.close()
        assert out[-1].T.tolist() == g["final"], f
        assert ev == sum(g["events"]), f
        assert int(out.astype(np.int64).sum()) == g["checksum"], f
        ran += 1
    assert ran >= (2 if kernel == "prebuilt" else len(files))  # vilar_macro_full and dimers_macro have build-time kernels


@pytest.mark.parametrize("kernel", ["prebuilt", "nvrtc"])
def test_headline_configuration_bit_exact(gpu, ffi, oracle, kernel):
    """The benchmarked combination itself (bench.py, BASELINE config C4): Vilar in define_system! arithmetic on
    the build-time kernel, dynamic schedule, t = 0..200 with 201 samples, more trajectories than resident lanes
    (148 SMs x 5 CTAs x 128 lanes = 94 720), so lanes claim further trajectories from the work counter.  The
    first 1024 trajectories and a strided subsample (the claimed ones included) are compared with the oracle on
    all 201 x 9 samples; the whole ensemble is checked through the event total of the subsample and through
    invariants of the network (gene copies conserved)."""
    model = models.vilar()
    n = 121_000
    net = models.build_network(model, 1)
    b = ffi.Batch(net, n, model["x0"], seeds=None, seed_base=0, kernel={"prebuilt": 3, "nvrtc": 2}[kernel])
    b.set_schedule(2)
    b.run_grid(200.0, 200)
    assert b.schedule_used == 2
    out = b.samples()
    b.close()
    head = models.seeds_sequence(1024)
    ref, _, _ = oracle.run_batch_macro("vilar", model["params"], model["x0"], head, 200.0, 200, threads=16)
    np.testing.assert_array_equal(out[:, :, :1024], ref)
    stride = 118
    sub = np.arange(1024 + 7, n, stride, dtype=np.uint64)  # 1017 trajectories, most of them claimed dynamically
    ref, _, _ = oracle.run_batch_macro("vilar", model["params"], model["x0"], sub, 200.0, 200, threads=16)
    np.testing.assert_array_equal(out[:, :, 1024 + 7::stride], ref)
    # Da + Dpa = 1 and Dr + Dpr = 1 for every trajectory at every sample
    assert (out[:, 0] + out[:, 2] == 1).all() and (out[:, 1] + out[:, 3] == 1).all()
    assert (out >= 0).all()


def test_full_size_properties(gpu, ffi):
    """BASELINE-size run (SIR, 10^6 trajectories x 251 samples, config C1) checked through properties
    that do not need the oracle: conservation, monotonicity, row 0, exact K4 sums, and equality of
    the two kernels on a strided subsample."""
    model = models.sir()
    n = 1_000_000
    net = models.build_network(model)
    b = ffi.Batch(net, n, model["x0"], seeds=None, seed_base=0, kernel=KERNELS["nvrtc"])
    b.run_grid(250.0, 250)
    out = b.samples()
    s1, s2 = b.sample_sums()
    ev = b.events()[0]
    b.close()
    assert out.shape == (251, 3, n)
    assert (out.sum(axis=1) == 1000).all()
    assert (out[0, 0] == 999).all() and (out[0, 1] == 1).all() and (out[0, 2] == 0).all()
    assert (np.diff(out[:, 0, ::97], axis=0) <= 0).all() and (np.diff(out[:, 2, ::97], axis=0) >= 0).all()
    np.testing.assert_array_equal(s1.reshape(251, 3), out.sum(axis=2, dtype=np.int64))
    assert 1.55e9 < ev < 1.70e9  # 1622 events per trajectory on average
    # every event moves exactly one individual: events = (S0 - S_end) + R_end summed over trajectories
    assert ev == int((999 - out[-1, 0].astype(np.int64)).sum() + out[-1, 2].astype(np.int64).sum())
    sub = np.arange(0, n, 4001, dtype=np.uint64)
    t, _, _ = run_product(ffi, model, sub, 250.0, 250, KERNELS["table"])
    np.testing.assert_array_equal(t, out[:, :, ::4001])


@pytest.mark.parametrize("kernel", ["table", "nvrtc", "prebuilt"])
@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("name,n,tmax,nb_steps", [
    ("sir", 2000, 250.0, 50),        # most trajectories end absorbing before tmax, crossings every few events
    ("dimers", 500, 0.5, 3),
    ("mm_lma", 1000, 100.0, 20),
    ("vilar", 96, 10.0, 10),
])
@pytest.mark.parametrize("schedule", [1, 2, 3])
def test_both_variants_bit_exact_vs_oracle(gpu, ffi, oracle, kernel, arith, name, n, tmax, nb_steps, schedule):
    """Both variants of every kernel, forced: static (ring-staged samples, uniform from a copy of the stream)
    and dynamic (draws made ahead of the propensities, stream stepped back on crossings and absorbing states,
    all-zero row for lanes without an event) against the oracle."""
    model = models.MODELS[name]()
    net = models.build_network(model, arith)
    if kernel == "prebuilt" and not net.has_prebuilt:
        pytest.skip("no build-time kernel for this network")
    seeds = numpy_seeds(n, rng=5)
    ref, _, ref_tot = oracle_network(oracle, model, arith).run_batch(model["x0"], seeds, tmax, nb_steps, threads=8)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel={"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel])
    b.set_schedule(schedule)
    b.run_grid(tmax, nb_steps)
    assert b.schedule_used == schedule
    np.testing.assert_array_equal(b.samples(), ref)
    assert b.events()[0] == ref_tot
    b.close()


@pytest.mark.parametrize("kernel", ["table", "nvrtc", "prebuilt"])
@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("log2_scale", [-600, -498, 499, 600])
@pytest.mark.parametrize("name,n,tmax,nb_steps", [("sir", 600, 250.0, 25), ("dimers", 200, 0.5, 3)])
def test_totals_outside_the_short_divide_range(gpu, ffi, oracle, kernel, arith, schedule, log2_scale, name, n, tmax,
                                               nb_steps):
    """The pass divides with a short sequence that is exact for totals in [2^-500, 2^500) and sends every other
    total through its side exit (IEEE divide).  Rate constants scaled by 2^s and the horizon by 2^-s put the
    totals below, across and above that range; the oracle divides in IEEE arithmetic throughout."""
    model = models.MODELS[name]()
    scale = 2.0 ** log2_scale
    model["reactions"] = [(k * scale, terms, diff) for k, terms, diff in model["reactions"]]
    model["params"] = [k * scale for k in model["params"]]
    tmax = tmax / scale
    net = models.build_network(model, arith)
    if kernel == "prebuilt" and not net.has_prebuilt:
        pytest.skip("no build-time kernel for this network")
    seeds = numpy_seeds(n, rng=13)
    ref, _, ref_tot = oracle_network(oracle, model, arith).run_batch(model["x0"], seeds, tmax, nb_steps, threads=8)
    assert ref_tot > 10 * n
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel={"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel])
    b.set_schedule(schedule)
    b.run_grid(tmax, nb_steps)
    np.testing.assert_array_equal(b.samples(), ref)
    assert b.events()[0] == ref_tot
    b.close()


@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_dynamic_schedule_is_bit_exact(gpu, ffi, oracle, kernel):
    """More trajectories than resident lanes: lanes claim further trajectories from the work counter; the
    result must not depend on the schedule (every trajectory owns its state and stream)."""
    model = models.sir()
    n = 150_000 if kernel == "nvrtc" else 40_000
    seeds = models.seeds_sequence(n, first=5)
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 60.0, 6, threads=8)
    net = models.build_network(model)
    outs = {}
    for schedule in (1, 2, 3):
        b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=KERNELS[kernel])
        b.set_schedule(schedule)
        b.run_grid(60.0, 6)
        assert b.schedule_used == schedule
        assert b.events()[0] == tot
        outs[schedule] = b.samples()
        b.close()
    for schedule in (1, 2, 3):
        np.testing.assert_array_equal(outs[schedule], ref)


def test_dynamic_schedule_resumes_exactly(gpu, ffi, oracle):
    model = models.dimers()
    n = 120_000
    seeds = models.seeds_sequence(n, first=1)
    net = models.build_network(model, 1)
    finals = {}
    for schedule in (1, 2, 3):
        b = ffi.Batch(net, n, model["x0"], seeds=seeds)
        b.set_schedule(schedule)
        for i in range(4):  # the grid loop of the binding, driven by the caller (src/lib.rs:129-133)
            b.advance_until(0.06 * i / 3)
        finals[schedule] = (b.species(), b.times(), b.events()[0])
        b.close()
    for schedule in (2, 3):
        np.testing.assert_array_equal(finals[1][0], finals[schedule][0])
        np.testing.assert_array_equal(finals[1][1], finals[schedule][1])
        assert finals[1][2] == finals[schedule][2]
    ref, _, tot = oracle_network(oracle, model, 1).run_batch(model["x0"], seeds[:2000], 0.06, 3, threads=8)
    np.testing.assert_array_equal(finals[2][0][:2000].T, ref[-1])


@pytest.mark.parametrize("kernel", ["table", "nvrtc", "prebuilt"])
@pytest.mark.parametrize("name,tmax,arith", [("sir", 60.0, 1), ("dimers", 0.02, 1), ("vilar", 0.5, 1), ("mm_lma", 5.0, 0),
                                              ("ring", 0.3, 0)])
def test_event_log_bit_exact(gpu, ffi, oracle, kernel, name, tmax, arith):
    """nb_steps = 0 (src/pyo3_gillespie.rs:209-223): every row -- time and counts -- of every trajectory equals
    the oracle's log; the batch can be advanced further afterwards (the final state was written back)."""
    model = models.MODELS[name]()
    net = models.build_network(model, arith)
    if kernel == "prebuilt" and not net.has_prebuilt:
        pytest.skip("no build-time kernel for this network")
    n = 70
    seeds = numpy_seeds(n, rng=23)
    kid = {"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel]
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=kid)
    offsets, times, samples = b.run_events(tmax)
    assert b.kernel_used == kid
    onet = oracle_network(oracle, model, arith)
    total_events = 0
    for i in range(n):
        ot, ox = onet.run_events(model["x0"], int(seeds[i]), tmax)
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        np.testing.assert_array_equal(times[lo:hi], ot)
        np.testing.assert_array_equal(samples[:, lo:hi].T, ox)
        total_events += len(ot) - 1 - int(np.isinf(ot[-1]))
    assert b.events()[1] == total_events
    # state written back: species and times are those of the last row
    last = offsets[1:].astype(np.int64) - 1
    np.testing.assert_array_equal(b.species().T, samples[:, last])
    np.testing.assert_array_equal(b.times(), times[last])
    # a subset of the species
    b2 = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=kid)
    o2, t2, s2 = b2.run_events(tmax, save_idx=[0, len(model["species"]) - 1])
    np.testing.assert_array_equal(t2, times)
    np.testing.assert_array_equal(s2, samples[[0, len(model["species"]) - 1]])
    b.close()
    b2.close()


def test_event_log_absorbing_and_empty(gpu, ffi):
    """An absorbing state ends the log with t = +inf (src/gillespie.rs:281-284); a trajectory already at tmax
    has the single initial row."""
    net = models.build_network(models.sir())
    b = ffi.Batch(net, 3, [10, 0, 0], seeds=np.arange(3, dtype=np.uint64))  # no infected: nothing can happen
    offsets, times, samples = b.run_events(50.0)
    assert offsets.tolist() == [0, 2, 4, 6]
    assert np.all(times[0::2] == 0.0) and np.all(np.isinf(times[1::2]))
    np.testing.assert_array_equal(samples[:, 0], [10, 0, 0])
    b.set_time(60.0)
    offsets, times, samples = b.run_events(50.0)
    assert offsets.tolist() == [0, 1, 2, 3] and np.all(times == 60.0)
    b.close()


def test_watchdog_with_unclaimed_trajectories(gpu, ffi, oracle):
    """Dynamic schedule, more trajectories than resident lanes, launches cut short by the watchdog: trajectories
    that were never claimed in a launch must still start from their own seed in a later one."""
    model = models.dimers()
    net = models.build_network(model, 1)
    n = 130_000
    seeds = numpy_seeds(n, rng=17)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds)
    b.set_schedule(2)
    b.set_max_iters(64)
    calls = 0
    while True:
        calls += 1
        try:
            b.advance_until(0.05)
            break
        except ffi.RebopError as e:
            assert e.status == ffi.ERR_ITER_CAP and calls < 400
    assert calls > 1
    ref, _, tot = oracle_network(oracle, model, 1).run_batch(model["x0"], seeds, 0.05, 0, threads=8)
    np.testing.assert_array_equal(b.species().T, ref[-1])
    assert b.events()[0] == tot
    b.close()


def test_segmented_run_with_host_buffer(gpu, ffi, oracle):
    """run_grid(host_out) on a large result runs the grid in segments and copies finished rows while the next
    segment is simulated: same samples, events and final state as one launch, and as the oracle."""
    model = models.sir()
    n = 300_000  # 251 x 3 x n x 4 B = 904 MB > the segmentation threshold
    seeds = models.seeds_sequence(n, first=77)
    net = models.build_network(model)
    a = ffi.Batch(net, n, model["x0"], seeds=seeds)
    host = np.empty((251, 3, n), dtype=np.int32)
    a.run_grid(250.0, 250, host_out=host)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds)
    b.run_grid(250.0, 250)
    np.testing.assert_array_equal(host, b.samples())
    np.testing.assert_array_equal(a.samples(), host)  # the device copy is whole as well
    assert a.events() == b.events()
    np.testing.assert_array_equal(a.species(), b.species())
    np.testing.assert_array_equal(a.times(), b.times())
    ref, _, _ = oracle_network(oracle, model).run_batch(model["x0"], seeds[:3000], 250.0, 250, threads=8)
    np.testing.assert_array_equal(host[:, :, :3000], ref)
    # event-dense network through the dynamic variant
    model = models.dimers()
    n = 2_000_000
    net = models.build_network(model, 1)
    a = ffi.Batch(net, n, model["x0"], seeds=None, seed_base=5)
    a.set_schedule(2)
    host = np.empty((16, 4, n), dtype=np.int32)  # 512 MB
    a.run_grid(0.15, 15, host_out=host)
    assert a.schedule_used == 2
    ref, _, _ = oracle_network(oracle, model, 1).run_batch(model["x0"], models.seeds_sequence(2000, 5), 0.15, 15, threads=8)
    np.testing.assert_array_equal(host[:, :, :2000], ref)
    a.close()
    b.close()


# ---- round 2: resume exactness, sample types, single steps, concurrency --------------------------------------------

@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("kernel", ["table", "nvrtc"])
def test_resume_after_iteration_cap_is_stream_exact(gpu, ffi, oracle, kernel, schedule):
    """Repeating a call the watchdog cut short continues every trajectory where it stopped and leaves finished ones
    alone: a trajectory already at tmax must not draw again (the reference made ONE advance_until call).  The
    interval that follows must therefore still match the oracle: species, times and event totals."""
    model = models.dimers()
    net = models.build_network(model)
    n = 96
    seeds = models.seeds_sequence(n, first=3)
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 0.1, 2, threads=8)  # rows at t = 0, 0.05, 0.1
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel=KERNELS[kernel])
    b.set_schedule(schedule)
    for target, row in ((0.0, 0), (0.05, 1), (0.1, 2)):
        b.set_max_iters(64)
        calls = 0
        while True:
            calls += 1
            try:
                b.advance_until(target)
                break
            except ffi.RebopError as e:
                assert e.status == ffi.ERR_ITER_CAP and calls < 400
        if target > 0:
            assert calls > 2
        np.testing.assert_array_equal(b.species().T, ref[row])
        assert (b.times() == target).all()
    assert b.events()[0] == tot
    b.close()


@pytest.mark.parametrize("schedule", [1, 2, 3])
def test_run_grid_resumes_after_iteration_cap(gpu, ffi, oracle, schedule):
    """run_grid cut short by the watchdog: the same call again continues from the grid points reached; the samples of
    the finished call are those of an uninterrupted run."""
    model = models.sir()
    net = models.build_network(model)
    n = 3000
    seeds = models.seeds_sequence(n, first=9)
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 250.0, 50, threads=8)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds)
    b.set_schedule(schedule)
    b.set_max_iters(128)
    calls = 0
    while True:
        calls += 1
        try:
            b.run_grid(250.0, 50)
            break
        except ffi.RebopError as e:
            assert e.status == ffi.ERR_ITER_CAP and calls < 200
    assert calls > 3
    np.testing.assert_array_equal(b.samples(), ref)
    assert b.events()[0] == tot
    s1, _ = b.sample_sums()
    np.testing.assert_array_equal(s1.reshape(51, 3), ref.sum(axis=2, dtype=np.int64))
    b.close()


@pytest.mark.parametrize("schedule", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.int16, np.int64])
def test_sample_types(gpu, ffi, oracle, dtype, schedule):
    """int16 / int64 samples are produced on the device; values, host copies in every type and the row sums agree
    with the int32 run and the oracle."""
    model = models.sir()
    net = models.build_network(model)
    n = 2100
    seeds = models.seeds_sequence(n, first=21)
    ref, _, _ = oracle_network(oracle, model).run_batch(model["x0"], seeds, 250.0, 40, threads=8)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds)
    b.set_schedule(schedule)
    b.set_sample_dtype(dtype)
    host = np.empty((41, 3, n), dtype=dtype)
    b.run_grid(250.0, 40, host_out=host)
    np.testing.assert_array_equal(host, ref.astype(dtype))
    assert b.samples().dtype == np.dtype(dtype)
    np.testing.assert_array_equal(b.samples(), ref.astype(dtype))
    np.testing.assert_array_equal(b.samples(np.int32), ref)
    np.testing.assert_array_equal(b.samples(np.int64), ref.astype(np.int64))
    s1, s2 = b.sample_sums()
    r64 = ref.astype(np.int64)
    np.testing.assert_array_equal(s1.reshape(41, 3), r64.sum(axis=2))
    np.testing.assert_array_equal(s2.reshape(41, 3).astype(np.int64), (r64 * r64).sum(axis=2))
    b.close()


@pytest.mark.parametrize("schedule", [1, 3])
def test_int16_overflow_is_an_error(gpu, ffi, schedule):
    """A count outside the int16 range must not wrap silently."""
    net = ffi.Network(1, 0)
    net.add_reaction_lma_sparse(1.0e5, [], [1])   # 0 -> A at 1e5 per unit time: about 1e5 molecules at t = 1
    b = ffi.Batch(net, 64, [0], seeds=models.seeds_sequence(64))
    b.set_schedule(schedule)
    b.set_sample_dtype(np.int16)
    with pytest.raises(ffi.RebopError) as e:
        b.run_grid(1.0, 4)
    assert e.value.status == ffi.ERR_LIMIT
    b.set_sample_dtype(np.int32)
    b.set_species([0])
    b.set_time(0.0)
    b.run_grid(1.0, 4)
    assert (b.samples()[-1] > 90000).all()
    b.close()


@pytest.mark.parametrize("kernel", ["table", "nvrtc", "prebuilt"])
def test_advance_one_reaction(gpu, ffi, oracle, kernel):
    """Gillespie::advance_one_reaction (src/gillespie.rs:270-297) for every trajectory: after k calls time, species and
    event count equal k single steps of the oracle; an absorbing state gives t = +inf."""
    model = models.sir()
    arith = 1 if kernel == "prebuilt" else 0
    net = models.build_network(model, arith)
    n = 40
    seeds = numpy_seeds(n, rng=31)
    b = ffi.Batch(net, n, model["x0"], seeds=seeds, kernel={"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel])
    onet = oracle_network(oracle, model, arith)
    for k in (1, 2, 7):
        b.set_species(model["x0"])
        b.set_time(0.0)
        b.seed(seeds)
        for _ in range(k):
            b.advance_one_reaction()
        t, x = b.times(), b.species()
        for i in range(n):
            rt, rx, _ = onet.step_one(model["x0"], int(seeds[i]), k)
            assert t[i] == rt and x[i].tolist() == rx.tolist(), (k, i)
    # nothing can happen: t = +inf, no random word used
    b2 = ffi.Batch(net, 8, [10, 0, 0], seeds=models.seeds_sequence(8), kernel={"table": 1, "nvrtc": 2, "prebuilt": 3}[kernel])
    b2.advance_one_reaction()
    assert np.isinf(b2.times()).all() and (b2.species() == [10, 0, 0]).all()
    b.close()
    b2.close()


@pytest.mark.parametrize("schedule", [1, 2])
def test_more_than_65535_sample_rows(gpu, ffi, schedule):
    """(nb_steps + 1) * n_save beyond the 65535 limit of a grid's y dimension: samples and row sums."""
    model = models.sir()
    net = models.build_network(model)
    n = 40
    b = ffi.Batch(net, n, model["x0"], seeds=models.seeds_sequence(n, 2))
    b.set_schedule(schedule)
    b.run_grid(250.0, 70_000, save_idx=[1])
    out = b.samples().astype(np.int64)
    assert out.shape == (70_001, 1, n)
    s1, s2 = b.sample_sums()
    np.testing.assert_array_equal(s1, out.sum(axis=2).ravel())
    np.testing.assert_array_equal(s2.astype(np.int64), (out * out).sum(axis=2).ravel())
    b.close()


def test_table_kernel_batches_with_different_networks_run_side_by_side(gpu, ffi, oracle):
    """Two host threads, two batches on the same device, two different networks on the table-driven kernel: the
    network travels with the batch (no process-global table), so neither run disturbs the other."""
    import threading
    ma, mb = models.sir(), models.dimers()
    na, nb = 20_000, 600
    sa, sb = models.seeds_sequence(na, 1), models.seeds_sequence(nb, 2)
    ra, _, _ = oracle_network(oracle, ma).run_batch(ma["x0"], sa, 250.0, 25, threads=8)
    rb, _, _ = oracle_network(oracle, mb).run_batch(mb["x0"], sb, 0.3, 3, threads=8)
    results = {}

    def work(key, model, seeds, tmax, nbs, reps):
        net = models.build_network(model)
        outs = []
        for _ in range(reps):
            b = ffi.Batch(net, len(seeds), model["x0"], seeds=seeds, kernel=1)
            b.run_grid(tmax, nbs)
            outs.append(b.samples())
            b.close()
        results[key] = outs

    ta = threading.Thread(target=work, args=("a", ma, sa, 250.0, 25, 6))
    tb = threading.Thread(target=work, args=("b", mb, sb, 0.3, 3, 6))
    ta.start(); tb.start(); ta.join(); tb.join()
    for o in results["a"]:
        np.testing.assert_array_equal(o, ra)
    for o in results["b"]:
        np.testing.assert_array_equal(o, rb)


def test_launches_on_a_caller_stream_do_not_block(gpu, ffi, oracle):
    """On a caller-owned stream run_grid returns before the device is done; rebop_batch_synchronize reports status and
    events, and the samples are those of a blocking run."""
    import torch
    model = models.dimers()
    net = models.build_network(model, 1)
    n = 4000
    seeds = models.seeds_sequence(n, first=4)
    ref, _, tot = oracle_network(oracle, model, 1).run_batch(model["x0"], seeds, 0.2, 4, threads=8)
    stream = torch.cuda.Stream()
    b = ffi.Batch(net, n, model["x0"], seeds=seeds)
    b.set_stream(stream.cuda_stream)
    host = ffi.PinnedBuffer((5, 4, n), np.int32)
    b.run_grid(0.2, 4, host_out=host.array)
    b.synchronize()
    np.testing.assert_array_equal(host.array, ref)
    assert b.events() == (tot, tot)
    b.set_max_iters(16)
    b.set_species(model["x0"])
    b.set_time(0.0)
    b.seed(seeds)
    b.run_grid(0.2, 4)  # returns at once; the cap is reported by synchronize
    with pytest.raises(ffi.RebopError) as e:
        b.synchronize()
    assert e.value.status == ffi.ERR_ITER_CAP
    b.set_stream(None)
    b.close()
    host.close()


def test_ensemble_handle_matches_single_batch(gpu, ffi, oracle):
    """rebop_ensemble_* over every visible device: samples, exact sums and the device-finalised mean / variance equal
    those of one batch on one device (and the oracle), whatever the device count."""
    from rebop_b200 import ensemble as ens
    model = models.sir()
    net = models.build_network(model)
    n = 10_001
    seeds = models.seeds_sequence(n, first=50)
    ref, _, tot = oracle_network(oracle, model).run_batch(model["x0"], seeds, 250.0, 20, threads=8)
    devices = list(range(ffi.device_count()))
    e = ffi.Ensemble(net, n, model["x0"], devices, seeds=seeds)
    assert sum(c for _, _, c in e.shards()) == n
    host = np.empty((21, 3, n), dtype=np.int32)
    e.run_grid(250.0, 20, host_out=host)
    np.testing.assert_array_equal(host, ref)
    np.testing.assert_array_equal(e.samples(), ref)
    assert e.events()[1] == tot
    s1, s2 = e.sums()
    r64 = ref.astype(np.int64)
    np.testing.assert_array_equal(s1.reshape(21, 3), r64.sum(axis=2))
    np.testing.assert_array_equal(s2.reshape(21, 3).astype(np.int64), (r64 * r64).sum(axis=2))
    mean, var = e.stats()
    m2, v2 = ens.finalize_stats(s1, s2, n)
    np.testing.assert_allclose(mean.ravel(), m2, rtol=1e-15, atol=0)
    np.testing.assert_allclose(var.ravel(), v2, rtol=1e-14, atol=0)
    e.close()


def test_ensemble_over_two_devices_is_bit_identical(gpu, ffi):
    """Needs two physical GPUs (skipped otherwise): 1-device and 2-device ensembles give the same samples and the same
    NCCL-all-reduced sums bit for bit."""
    if ffi.device_count() < 2:
        pytest.skip("needs two GPUs")
    model = models.dimers()
    net = models.build_network(model, 1)
    n = 50_001
    one = ffi.Ensemble(net, n, model["x0"], [0], seed_base=7)
    two = ffi.Ensemble(net, n, model["x0"], [0, 1], seed_base=7)
    one.run_grid(0.2, 4)
    two.run_grid(0.2, 4)
    np.testing.assert_array_equal(one.samples(), two.samples())
    a1, a2 = one.sums()
    b1, b2 = two.sums()
    np.testing.assert_array_equal(a1, b1)
    np.testing.assert_array_equal(a2, b2)
    m1, v1 = one.stats()
    m2, v2 = two.stats()
    np.testing.assert_array_equal(m1, m2)
    np.testing.assert_array_equal(v1, v2)
    assert one.events() == two.events()
    one.close()
    two.close()

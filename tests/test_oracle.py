"""Pins the CPU oracle to every known answer the reference holds for this path.

The oracle is the checker the GPU parity tests compare against, so it is validated first:
  * tests/test_rebop.py:30-36   test_fixed_seed: rng=42 => S=0, I=227, R=773 (end-to-end golden vector)
  * src/gillespie.rs:448-472    rate_lma known answers (dense and sparse forms)
  * src/expr.rs:499-548         test_eval 9.049998877643098, test_eval_max
  * tests/test_rebop.py:55-65   dense == sparse
  * src/gillespie_macro.rs:224-253  NaN parameter freezes the system; empty system
  * tests/test_rebop.py:39-52   nb_steps = 0 invariants
  * SURVEY.md Appendix A        RNG known-answer vectors (consistent with the golden vector)
"""
import json
import os

import numpy as np
import pytest

from rebop_b200 import models
from tests.helpers import numpy_seeds, oracle_network

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_numpy_seed_derivation():
    """python/rebop/gillespie.py:139-140."""
    for rng, want in ((0, 11749869230777074270), (1, 9441442522235856126), (42, 14276969152011380359)):
        assert int(np.random.default_rng(rng).integers(np.iinfo(np.uint64).max, dtype=np.uint64)) == want
    many = numpy_seeds(5, rng=3)
    g = np.random.default_rng(3)
    one_by_one = [int(g.integers(np.iinfo(np.uint64).max, dtype=np.uint64)) for _ in range(5)]
    assert many.tolist() == one_by_one


@pytest.mark.parametrize("seed,state,first3,exp1,unif", [
    (0, "e220a8397b1dcdaf 6e789e6aa1b965f4 06c45d188009454f f88bb8a8724c81ec",
     "53175d61490b23df 61da6f3dc380d507 5c0fdf91ec9a7bfc", 0.1970678933693451, 0.38223929651167343),
    (42, "bdd732262feb6e95 28efe333b266f103 47526757130f9f52 581ce1ff0e4ae394",
     "d0764d4f4476689f 519e4174576f3791 fbe07cfb0c24ed8c", 1.0640204579905184, 0.3188210400616611),
    (14276969152011380359, "05d65de5d4f10c4b 283eced139243cb1 75ba3f5cc49388dd 8f26a002642bfb30",
     "f9f2ec6992bb8ac9 af9a1fd23fccaafe eb6ce57c1d30870b", 0.8331471285708154, 0.6859455002120735),
])
def test_rng_known_answers(oracle, seed, state, first3, exp1, unif):
    r = oracle.RngStream(seed)
    assert " ".join("%016x" % v for v in r.state) == state
    assert " ".join("%016x" % r.next_u64() for _ in range(3)) == first3
    r = oracle.RngStream(seed)
    assert r.exp1() == exp1
    assert r.uniform() == unif


def test_exp1_distribution(oracle):
    r = oracle.RngStream(123)
    x = np.array([r.exp1() for _ in range(200000)])
    assert (x >= 0).all()
    assert abs(x.mean() - 1.0) < 0.01 and abs(x.var() - 1.0) < 0.03
    # survival function at a few points
    for q in (0.5, 1.0, 3.0):
        assert abs((x > q).mean() - np.exp(-q)) < 0.005


def test_reference_golden_vector(oracle):
    """tests/test_rebop.py:30-36."""
    seed = int(np.random.default_rng(42).integers(np.iinfo(np.uint64).max, dtype=np.uint64))
    model = models.sir()
    for dense in (False, True):
        for arith in (0, 1):
            times, out, events = oracle_network(oracle, model, arith, dense).run_grid(model["x0"], seed, 250.0, 250)
            assert out[-1].tolist() == [0, 227, 773]
            assert out[0].tolist() == [999, 1, 0]
            assert events == 1772
            np.testing.assert_array_equal(times, np.arange(251.0))


@pytest.mark.parametrize("dense", [False, True])
def test_rate_lma_table(oracle, dense):
    """src/gillespie.rs:448-472: species [5, 3], k = 2."""
    table = [([0, 0], 2.0), ([1, 0], 10.0), ([2, 0], 40.0), ([3, 0], 120.0), ([4, 0], 240.0), ([5, 0], 240.0),
             ([6, 0], 0.0), ([0, 1], 6.0), ([1, 1], 30.0), ([2, 1], 120.0), ([0, 2], 12.0), ([1, 20], 0.0)]
    for exps, want in table:
        net = oracle.Network(2, dense=dense)
        net.add_lma(2.0, [(s, e) for s, e in enumerate(exps) if e > 0 or dense], [0, 0])
        assert net.rate(0, [5, 3]) == want, (exps, want)


def test_expr_eval_known_answers(oracle):
    """src/expr.rs:499-548."""
    A, B, C_, D, E, F = range(6)
    # 1.21 * C + B - A / D ^ E * (F + exp(D)), post-order
    prog = [("const", 0, 1.21), ("species", C_, 0), ("mul", 0, 0), ("species", B, 0), ("add", 0, 0),
            ("species", A, 0), ("species", D, 0), ("species", E, 0), ("pow", 0, 0), ("div", 0, 0),
            ("species", F, 0), ("species", D, 0), ("exp", 0, 0), ("add", 0, 0), ("mul", 0, 0), ("sub", 0, 0)]
    x = np.array([2, 3, 5, 7, 11, 13], dtype=np.int64)
    got = oracle.lib().ora_expr_eval(oracle.make_prog(prog), len(prog), x.ctypes.data_as(oracle.C.POINTER(oracle.C.c_int64)))
    assert got == 9.049998877643098
    three = np.array([3], dtype=np.int64)
    p3 = three.ctypes.data_as(oracle.C.POINTER(oracle.C.c_int64))
    for sub, want in ((None, 3.0), (0.5, 2.5), (3.0, 0.0), (3.5, 0.0)):
        prog = [("species", 0, 0)] + ([("const", 0, sub), ("sub", 0, 0)] if sub is not None else []) + \
               [("const", 0, 0.0), ("max", 0, 0)]
        assert oracle.lib().ora_expr_eval(oracle.make_prog(prog), len(prog), p3) == want


@pytest.mark.parametrize("name", ["sir", "dimers", "vilar", "mm_lma"])
def test_dense_equals_sparse(oracle, name):
    """tests/test_rebop.py:55-65."""
    model = models.MODELS[name]()
    tmax, nb = (model["tmax"], 10) if name != "vilar" else (5.0, 5)
    seeds = numpy_seeds(16, rng=1)
    a, ea, ta = oracle_network(oracle, model, 0, dense=False).run_batch(model["x0"], seeds, tmax, nb)
    b, eb, tb = oracle_network(oracle, model, 0, dense=True).run_batch(model["x0"], seeds, tmax, nb)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ea, eb)
    assert ta == tb


@pytest.mark.parametrize("name", ["sir", "dimers", "vilar"])
def test_macro_specialisation_equals_generic(oracle, name):
    """The hand-expanded define_system! code (timed CPU baseline) is the generic macro-arithmetic walk."""
    model = models.MODELS[name]()
    tmax, nb = (model["tmax"], 10) if name != "vilar" else (10.0, 10)
    seeds = models.seeds_sequence(24)
    a, ea, ta = oracle.run_batch_macro(name, model["params"], model["x0"], seeds, tmax, nb, threads=2)
    b, eb, tb = oracle_network(oracle, model, 1).run_batch(model["x0"], seeds, tmax, nb)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ea, eb)


def test_api_and_macro_arithmetic_agree_on_benchmarks(oracle):
    """same integer trajectories, the flavours differ only in the last bits of t."""
    for name, tmax, nb in (("sir", 250.0, 50), ("dimers", 1.0, 2)):
        model = models.MODELS[name]()
        seeds = models.seeds_sequence(32, 5)
        a, _, _ = oracle_network(oracle, model, 0).run_batch(model["x0"], seeds, tmax, nb)
        b, _, _ = oracle_network(oracle, model, 1).run_batch(model["x0"], seeds, tmax, nb)
        np.testing.assert_array_equal(a, b)


def test_sir_conservation_and_bounds(oracle):
    """src/gillespie.rs:504-516, tests/test_rebop.py:15-27."""
    model = models.sir()
    out, ev, _ = oracle_network(oracle, model).run_batch(model["x0"], models.seeds_sequence(64), 250.0, 250)
    assert (out.sum(axis=1) == 1000).all()
    assert (out >= 0).all()
    assert (np.diff(out[:, 0, :], axis=0) <= 0).all() and (np.diff(out[:, 2, :], axis=0) >= 0).all()
    assert ev.max() <= 1999


def test_dimers_bounds(oracle):
    """src/gillespie.rs:518-530, src/gillespie_macro.rs:193-212."""
    model = models.dimers()
    for arith in (0, 1):
        out, _, _ = oracle_network(oracle, model, arith).run_batch(model["x0"], models.seeds_sequence(8), 1.0, 1)
        assert (out[-1, 0] == 1).all() and (out[-1, 2] > 1000).all() and (out[-1, 3] < 10000).all() and (out[-1, 3] > 1000).all()


def test_nan_parameter_freezes(oracle):
    """src/gillespie_macro.rs:224-238."""
    net = oracle.Network(1, [("lma", 10.0, [], [1]), ("lma", float("nan"), [(0, 1)], [-1])], arith=1)
    times, out, events = net.run_grid([0], 7, 100.0, 4)
    assert events == 0 and (out == 0).all()
    np.testing.assert_array_equal(times, [0.0, 25.0, 50.0, 75.0, 100.0])


def test_empty_system(oracle):
    """src/gillespie_macro.rs:239-253, tests/test_rebop.py:148-152."""
    net = oracle.Network(3, [])
    _, out, events = net.run_grid([42, 1337, 0], 1, 1e20, 3)
    assert events == 0
    assert (out == [42, 1337, 0]).all()


def test_all_reactions_mode(oracle):
    """tests/test_rebop.py:39-52 (nb_steps = 0): one row per event, last time > tmax or inf."""
    model = models.sir()
    net = oracle_network(oracle, model)
    for seed in range(6):
        times, out = net.run_events(model["x0"], seed, 250.0)
        assert times[0] == 0.0 and (np.diff(times) > 0).all()
        assert times[-1] > 250.0 or np.isinf(times[-1])
        d = np.diff(out, axis=0)
        assert set(map(tuple, d[:-1] if np.isinf(times[-1]) else d)) <= {(-1, 1, 0), (0, -1, 1)}
        assert (out.sum(axis=1) == 1000).all()


def test_var_names_subset(oracle):
    """tests/test_rebop.py:68-89."""
    model = models.sir()
    net = oracle_network(oracle, model)
    _, full, _ = net.run_grid(model["x0"], 9, 250.0, 25)
    _, sub, _ = net.run_grid(model["x0"], 9, 250.0, 25, save_idx=[0, 2])
    np.testing.assert_array_equal(sub, full[:, [0, 2]])


def test_golden_fixtures(oracle):
    """Committed fixtures (tests/golden/*.json, written by tests/golden/make_golden.py from the oracle
    after it was pinned above): guards the oracle itself against regressions."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".json"))
    assert files, "no golden fixtures committed"
    for f in files:
        g = json.load(open(os.path.join(GOLDEN, f)))
        model = models.MODELS[g["model"]]()
        net = oracle_network(oracle, model, g["arith"])
        seeds = np.array(g["seeds"], dtype=np.uint64)
        out, ev, tot = net.run_batch(model["x0"], seeds, g["tmax"], g["nb_steps"])
        assert out[-1].T.tolist() == g["final"], f
        assert ev.tolist() == g["events"], f
        assert int(out.astype(np.int64).sum()) == g["checksum"], f

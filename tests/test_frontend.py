"""The Python front end `rebop_b200.Gillespie` against the reference's own Python tests
(tests/test_rebop.py of the reference), plus the new batch-of-trajectories argument.

Host-only behaviour (model building, printing, error messages) runs without a GPU; everything
that simulates is marked gpu and goes through the C ABI.
"""

import numpy as np
import numpy.testing as npt
import pytest

import rebop_b200
from rebop_b200 import models


def sir_model(transmission=1e-4, recovery=0.01):
    sir = rebop_b200.Gillespie()
    sir.add_reaction(transmission, ["S", "I"], ["I", "I"])
    sir.add_reaction(recovery, ["I"], ["R"])
    return sir


# ---- host only ----------------------------------------------------------------------------------
def test_species_order_and_counts():
    """src/pyo3_gillespie.rs:76-80,99-106: first appearance, reactants before products."""
    s = rebop_b200.Gillespie()
    s.add_reaction(0.1, ["A", "B"], ["C"], 0.01)
    s.add_reaction("0.2 * B * C / (5 + C)", ["B"], ["D"])
    assert list(s._species) == ["A", "B", "C", "D"]
    assert s.nb_species() == 4 and s.nb_reactions() == 3  # reverse_rate added the reverse reaction


def test_str():
    """src/pyo3_gillespie.rs:239-252 and Display for PRate (:30-37)."""
    s = sir_model()
    s.add_reaction("1.20*S / (3.5+S)", ["S"], [])
    assert str(s) == ("3 species and 3 reactions\n"
                      "S + I --> I + I @ LMA(0.0001)\n"
                      "I --> R @ LMA(0.01)\n"
                      "S -->  @ ((1.2 * S) / (3.5 + S))\n")
    t = rebop_b200.Gillespie()
    t.add_reaction(14, [], ["A"])
    assert str(t) == "1 species and 1 reactions\n --> A @ LMA(14)\n"


def test_rate_parse_errors():
    """tests/test_rebop.py:92-97."""
    s = rebop_b200.Gillespie()
    with pytest.raises(ValueError, match="Rate expression not understood"):
        s.add_reaction("+", [], ["A"])
    with pytest.raises(ValueError, match="Rate expression not understood"):
        s.add_reaction("1+", [], ["A"])
    assert s.nb_species() == 0 and s.nb_reactions() == 0  # the rate is parsed before anything is registered


def test_set_init_warns_after_storing():
    """src/pyo3_gillespie.rs:119-134."""
    s = rebop_b200.Gillespie()
    s.add_reaction("B", [], ["A"])
    with pytest.raises(UserWarning, match="species are not involved in any reaction"):
        s.set_init({"B": 1})
    assert s._init == {"B": 1} and list(s._species) == ["A", "B"]


def test_lowering_errors_and_tables(ffi):
    """tests/test_rebop.py:155-166 (messages), src/pyo3_gillespie.rs:180-196 (tables)."""
    s = rebop_b200.Gillespie()
    s.add_reaction("k", [], ["A"])
    with pytest.raises(ValueError, match="Parameter k should have a value"):
        s._lower({}, ffi.ARITH_API)
    net = s._lower({"k": 0.4}, ffi.ARITH_API)
    assert net.nb_reactions == 1
    # 2A -> B lowers to exponent 2 on A (reactant multiset), jump (-2, +1)
    d = rebop_b200.Gillespie()
    d.add_reaction(0.001, ["A", "A"], ["B"])
    src = d._lower({}, ffi.ARITH_API).codegen()
    # falling-factorial factor; packed jump: -2 = 0xfe, +1 = 0x01, and the event-count lane (1) behind the species
    assert "__dsub_rn(d0, 1.0)" in src and "0x000101fe" in src


# ---- GPU: the reference's tests -----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [None, *range(10)])
def test_sir(gpu, seed):
    """tests/test_rebop.py:15-27."""
    ds = sir_model().run({"S": 999, "I": 1}, tmax=250, nb_steps=250, rng=seed)
    npt.assert_array_equal(ds.time, np.arange(251))
    for v in (ds.S, ds.I, ds.R):
        assert (np.asarray(v) >= 0).all() and (np.asarray(v) <= 1000).all()
    assert (np.asarray(ds.S) <= 999).all()
    npt.assert_array_equal(np.asarray(ds.S) + np.asarray(ds.I) + np.asarray(ds.R), [1000] * 251)
    assert np.asarray(ds.S).dtype == np.int64 and np.asarray(ds.S).shape == (251,)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "nvrtc", "table"])
def test_fixed_seed(gpu, kernel):
    """tests/test_rebop.py:30-36: the reference's golden vector through the reference's own call."""
    ds = sir_model().run({"S": 999, "I": 1}, tmax=250, nb_steps=250, rng=42, kernel=kernel)
    assert ds.S[-1] == 0
    assert ds.I[-1] == 227
    assert ds.R[-1] == 773


@pytest.mark.gpu
def test_dense_vs_sparse(gpu):
    """tests/test_rebop.py:55-65."""
    sir = sir_model()
    runs = [sir.run({"S": 999, "I": 1}, tmax=250, nb_steps=250, rng=42, sparse=sp) for sp in (None, False, True)]
    for other in runs[1:]:
        for name in ("S", "I", "R"):
            npt.assert_array_equal(np.asarray(runs[0][name]), np.asarray(other[name]))


@pytest.mark.gpu
def test_var_names(gpu):
    """tests/test_rebop.py:68-89 (nb_steps = 250 case)."""
    sir = sir_model()
    init = {"S": 999, "I": 1}
    ds_all = sir.run(init, tmax=250, nb_steps=250, rng=0, var_names=None)
    ds_subset = sir.run(init, tmax=250, nb_steps=250, rng=0, var_names=["S", "I"])
    assert "S" in ds_subset and "I" in ds_subset and "R" not in ds_subset
    for name in ("S", "I"):
        npt.assert_array_equal(np.asarray(ds_all[name]), np.asarray(ds_subset[name]))
    with pytest.raises(KeyError):
        sir.run(init, tmax=250, nb_steps=250, rng=0, var_names=["nope"])


@pytest.mark.gpu
def test_arbitrary_rates(gpu):
    """tests/test_rebop.py:92-114."""
    s = rebop_b200.Gillespie()
    s.add_reaction("B", [], ["A"])
    with pytest.warns(UserWarning, match="species are not involved in any reaction"):
        ds = s.run({"B": 1}, tmax=10, nb_steps=100)
    assert ds.A[-1] >= 1
    npt.assert_equal(np.asarray(ds.B), [1] * 101)

    s = rebop_b200.Gillespie()
    s.add_reaction("B", [], ["A"])
    s.add_reaction(1, [], ["B"])
    assert s.run({}, tmax=10, nb_steps=100).A[-1] >= 1


@pytest.mark.gpu
def test_arbitrary_rates_crossed(gpu):
    """tests/test_rebop.py:117-134: all-zero state with crossed expression rates never moves."""
    s = rebop_b200.Gillespie()
    s.add_reaction("B", [], ["A"])
    s.add_reaction("A", [], ["B"])
    ds = s.run({}, tmax=10, nb_steps=10)
    npt.assert_array_equal(np.asarray(ds.A), [0] * 11)
    npt.assert_array_equal(np.asarray(ds.B), [0] * 11)
    npt.assert_array_equal(np.asarray(ds.time), np.linspace(0, 10, 11))
    ds = s.run({"A": 1}, tmax=10, nb_steps=10)
    assert ds.A[-1] > 1 and ds.B[-1] > 0
    ds = s.run({"B": 1}, tmax=10, nb_steps=10)
    assert ds.A[-1] > 0 and ds.B[-1] > 1


@pytest.mark.gpu
def test_arbitrary_rates_2(gpu):
    """tests/test_rebop.py:137-145."""
    s = rebop_b200.Gillespie()
    s.add_reaction(14, [], ["A"])
    s.add_reaction(0.1, ["A", "B"], ["C"], 0.01)
    s.add_reaction("0.2 * B * C / (5 + C)", ["B"], ["D"])
    ds = s.run({"B": 1000}, tmax=100, nb_steps=100)
    assert (np.asarray(ds.B) + np.asarray(ds.C) + np.asarray(ds.D) == 1000).all()
    assert ds.D[-1] >= 1


@pytest.mark.gpu
def test_run_empty(gpu):
    """tests/test_rebop.py:148-152."""
    ds = rebop_b200.Gillespie().run({}, tmax=10, nb_steps=10)
    assert len(ds.data_vars) == 0
    npt.assert_array_equal(np.asarray(ds.time), np.linspace(0, 10, 11))


@pytest.mark.gpu
def test_parameters(gpu):
    """tests/test_rebop.py:155-166."""
    s = rebop_b200.Gillespie()
    s.add_reaction(4, ["A"], ["B"])
    with pytest.raises(ValueError, match="Species B cannot also be a parameter"):
        s.run({}, 10, 10, params={"B": 4.2})
    s = rebop_b200.Gillespie()
    s.add_reaction("k", [], ["A"])
    with pytest.raises(ValueError, match="Parameter k should have a value"):
        s.run({}, 10, 10)
    assert s.run({}, 10, 10, params={"k": 0.4}).A[-1] > 0


# ---- GPU: the batch argument ----------------------------------------------------------------------
@pytest.mark.gpu
def test_batch_equals_successive_reference_runs(gpu):
    """Trajectory n of a batch == the n-th successive run on the same generator (gillespie.py:139-140)."""
    sir = sir_model()
    init = {"S": 999, "I": 1}
    n = 12
    batch = sir.run(init, tmax=250, nb_steps=50, rng=np.random.default_rng(5), n_trajectories=n)
    assert np.asarray(batch.S).shape == (51, n)
    g = np.random.default_rng(5)
    for i in range(n):
        one = sir.run(init, tmax=250, nb_steps=50, rng=g)
        for name in ("S", "I", "R"):
            npt.assert_array_equal(np.asarray(batch[name])[:, i], np.asarray(one[name]))


@pytest.mark.gpu
def test_michaelis_menten_expression_batch(gpu, oracle):
    """examples/mm.py:10-17 (BASELINE config C3) as a batch, against the oracle's Expr::eval walk."""
    mm = rebop_b200.Gillespie()
    mm.add_reaction("V * A / (Km + A)", ["A"], ["P"])
    n = 4096
    ds = mm.run({"A": 100}, tmax=250, nb_steps=100, params={"V": 1, "Km": 20}, rng=0, n_trajectories=n, dtype=np.int32)
    seeds = np.random.default_rng(0).integers(np.iinfo(np.uint64).max, size=n, dtype=np.uint64)
    prog = [("const", 0, 1.0), ("species", 0, 0), ("mul", 0, 0), ("const", 0, 20.0), ("species", 0, 0), ("add", 0, 0),
            ("div", 0, 0)]
    ref, _, _ = oracle.Network(2, [("expr", prog, [-1, 1])]).run_batch([100, 0], seeds, 250.0, 100, threads=4)
    npt.assert_array_equal(np.asarray(ds.A), ref[:, 0, :])
    npt.assert_array_equal(np.asarray(ds.P), ref[:, 1, :])


@pytest.mark.gpu
def test_reduce_returns_ensemble_moments(gpu):
    sir = sir_model()
    init = {"S": 999, "I": 1}
    full = sir.run(init, tmax=250, nb_steps=25, rng=3, n_trajectories=2000)
    red = sir.run(init, tmax=250, nb_steps=25, rng=3, n_trajectories=2000, reduce=True)
    for name in ("S", "I", "R"):
        x = np.asarray(full[name]).astype(np.float64)
        npt.assert_allclose(np.asarray(red[name + "_mean"]), x.mean(axis=1), rtol=1e-13)
        npt.assert_allclose(np.asarray(red[name + "_var"]), x.var(axis=1, ddof=1), rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_sharding_over_devices_does_not_change_results(gpu, ffi):
    sir = sir_model()
    init = {"S": 999, "I": 1}
    devs = list(range(min(2, ffi.device_count()))) * (2 if ffi.device_count() == 1 else 1)
    a = sir.run(init, tmax=250, nb_steps=20, rng=11, n_trajectories=777)
    b = sir.run(init, tmax=250, nb_steps=20, rng=11, n_trajectories=777, devices=devs)
    for name in ("S", "I", "R"):
        npt.assert_array_equal(np.asarray(a[name]), np.asarray(b[name]))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(10))
def test_all_reactions(gpu, seed):
    """tests/test_rebop.py:39-52 (nb_steps = 0)."""
    tmax = 250
    ds = sir_model().run({"S": 999, "I": 1}, tmax=tmax, nb_steps=0, rng=seed)
    t = np.asarray(ds.time)
    assert t[0] == 0
    assert t[-1] > tmax
    assert np.all(np.diff(t) > 0)
    keep = slice(None, -1) if np.isinf(t[-1]) else slice(None)
    assert set(np.diff(np.asarray(ds.S)[keep])) <= {-1, 0}
    assert set(np.diff(np.asarray(ds.I)[keep])) <= {-1, 1}
    assert set(np.diff(np.asarray(ds.R)[keep])) <= {0, 1}


@pytest.mark.gpu
def test_var_names_all_reactions(gpu):
    """tests/test_rebop.py:68-89 (nb_steps = 0 case)."""
    sir = sir_model()
    init = {"S": 999, "I": 1}
    ds_all = sir.run(init, tmax=250, nb_steps=0, rng=0, var_names=None)
    ds_subset = sir.run(init, tmax=250, nb_steps=0, rng=0, var_names=["S", "I"])
    assert "S" in ds_subset and "I" in ds_subset and "R" not in ds_subset
    npt.assert_array_equal(np.asarray(ds_all.time), np.asarray(ds_subset.time))
    for name in ("S", "I"):
        npt.assert_array_equal(np.asarray(ds_all[name]), np.asarray(ds_subset[name]))


@pytest.mark.gpu
def test_all_reactions_batch_equals_single_runs_and_oracle(gpu, oracle):
    """Event logs of a batch: trajectory i equals the i-th of N successive single runs, and the oracle's log."""
    sir = sir_model()
    init = {"S": 999, "I": 1}
    n = 40
    many = sir.run(init, tmax=100, nb_steps=0, rng=3, n_trajectories=n)
    assert isinstance(many, list) and len(many) == n
    gen = np.random.default_rng(3)
    singles = [sir.run(init, tmax=100, nb_steps=0, rng=gen) for _ in range(n)]
    from rebop_b200 import models
    from tests.helpers import numpy_seeds, oracle_network
    onet = oracle_network(oracle, models.sir())
    seeds = numpy_seeds(n, rng=3)
    for i in range(n):
        npt.assert_array_equal(np.asarray(many[i].time), np.asarray(singles[i].time))
        ot, ox = onet.run_events(models.sir()["x0"], int(seeds[i]), 100.0)
        npt.assert_array_equal(np.asarray(many[i].time), ot)
        for j, name in enumerate("SIR"):
            npt.assert_array_equal(np.asarray(many[i][name]), np.asarray(singles[i][name]))
            npt.assert_array_equal(np.asarray(many[i][name]), ox[:, j])

"""Benchmark networks of the reference, as plain data.

Each model is a dict: species (names, index order), x0, reactions as
(k, [(species_index, exponent), ...] in evaluation order, differences[S]), and the grid the
reference's benchmark/example uses.  Citations are into the reference tree.
"""
from __future__ import annotations

import numpy as np


def _lma(S, k, reactants, products):
    """Mass-action reaction from index lists (pyo3 lowering, src/pyo3_gillespie.rs:180-196)."""
    exps = [0] * S
    diff = [0] * S
    for r in reactants:
        exps[r] += 1
        diff[r] -= 1
    for p in products:
        diff[p] += 1
    terms = [(s, e) for s, e in enumerate(exps) if e > 0]
    return (float(k), terms, diff)


def sir(transmission=1e-4, recovery=0.01):
    """tests/test_rebop.py:8-12; src/lib.rs:122-127."""
    S = 3
    return dict(name="sir", species=["S", "I", "R"], x0=[999, 1, 0], tmax=250.0, nb_steps=250,
                params=[transmission, recovery],
                reactions=[_lma(S, transmission, [0, 1], [1, 1]), _lma(S, recovery, [1], [2])])


def dimers(rtx=25.0, rtl=1000.0, rdi=0.001, rdm=0.1, rdp=1.0):
    """examples/dimers.rs:4-12,17-23; src/gillespie.rs:301-314."""
    S = 4
    return dict(name="dimers", species=["gene", "mRNA", "protein", "dimer"], x0=[1, 0, 0, 0], tmax=1.0, nb_steps=1,
                params=[rtx, rtl, rdi, rdm, rdp],
                reactions=[_lma(S, rtx, [0], [0, 1]), _lma(S, rtl, [1], [1, 2]), _lma(S, rdi, [2, 2], [3]),
                           _lma(S, rdm, [1], []), _lma(S, rdp, [2], [])])


def vilar():
    """benchmarks/benches/vilar/vilar.rs:6-46 (macro) == benchmarks/benches/my_benchmark.rs:235-272 (API)."""
    S = 9
    Da, Dr, Dpa, Dpr, Ma, Mr, A, R, C = range(9)
    aA, apA, aR, apR, bA, bR, dMA, dMR, dA, dR, gA, gR, gC, tA, tR = (
        50.0, 500.0, 0.01, 50.0, 50.0, 5.0, 10.0, 0.5, 1.0, 0.2, 1.0, 1.0, 2.0, 50.0, 100.0)
    rx = [
        _lma(S, gA, [Da, A], [Dpa]), _lma(S, gR, [Dr, A], [Dpr]),
        _lma(S, tA, [Dpa], [Da, A]), _lma(S, tR, [Dpr], [Dr, A]),
        _lma(S, aA, [Da], [Da, Ma]), _lma(S, aR, [Dr], [Dr, Mr]),
        _lma(S, apA, [Dpa], [Dpa, Ma]), _lma(S, apR, [Dpr], [Dpr, Mr]),
        _lma(S, bA, [Ma], [Ma, A]), _lma(S, bR, [Mr], [Mr, R]),
        _lma(S, gC, [A, R], [C]), _lma(S, dA, [C], [R]),
        _lma(S, dMA, [Ma], []), _lma(S, dMR, [Mr], []),
        _lma(S, dA, [A], []), _lma(S, dR, [R], []),
    ]
    return dict(name="vilar", species=["Da", "Dr", "Dpa", "Dpr", "Ma", "Mr", "A", "R", "C"],
                x0=[1, 1, 0, 0, 0, 0, 0, 0, 0], tmax=200.0, nb_steps=200,
                params=[aA, apA, aR, apR, bA, bR, dMA, dMR, dA, dR, gA, gR, gC, tA, tR], reactions=rx)


def mm_lma():
    """Michaelis-Menten, mass-action form: benchmarks/benches/my_benchmark.rs:199-233."""
    S = 4
    return dict(name="mm_lma", species=["E", "S", "ES", "P"], x0=[301, 120, 0, 0], tmax=100.0, nb_steps=100,
                params=[0.0017, 0.5, 0.1],
                reactions=[_lma(S, 0.0017, [0, 1], [2]), _lma(S, 0.5, [2], [0, 1]), _lma(S, 0.1, [2], [3])])


def ring(n=50, k=1.0, a0=1000, tmax=100.0, nb_steps=10):
    """benchmarks/benches/my_benchmark.rs:601-614 (api_ring): X_i -> X_(i+1 mod n), all mass on X_0."""
    rx = [_lma(n, k, [i], [(i + 1) % n]) for i in range(n)]
    return dict(name="ring", species=["A%d" % i for i in range(n)], x0=[a0] + [0] * (n - 1), tmax=float(tmax),
                nb_steps=nb_steps, params=[k], reactions=rx)


def flocculation(n=50, k=1.0, n0=1000, tmax=1000.0, nb_steps=10):
    """benchmarks/benches/my_benchmark.rs:683-700 (api_flocculation): A_i + A_j -> A_(i+j), i <= j, i + j <= n."""
    rx = []
    for i in range(1, n // 2 + 1):
        for j in range(i, n - i + 1):
            rx.append(_lma(n, k, [i - 1, j - 1], [i + j - 1]))
    return dict(name="flocculation", species=["A%d" % (i + 1) for i in range(n)], x0=[n0] + [0] * (n - 1),
                tmax=float(tmax), nb_steps=nb_steps, params=[k], reactions=rx)


def _splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return state, z ^ (z >> 31)


def synthetic(n_species=100, n_reactions=500, gen_seed=20240501, tmax=0.2, nb_steps=100):
    """Deterministic random mass-action network (not in the reference).

    Species s has mass 1 + (s mod 3); reactions come in reversible pairs: with probability 0.4 an
    isomerisation X <-> Y between species of equal mass, otherwise a binding X + Y <-> Z with
    m_Z = m_X + m_Y (X == Y allowed, which exercises the order-2 falling factorial).  Total mass is
    conserved, so counts stay bounded.  Rate constants are log-uniform: bimolecular in
    [1e-3, 1e-1], unimolecular in [1e-1, 10].  x0 uniform in [50, 150].  Everything is drawn from
    one SplitMix64 stream seeded with gen_seed.
    """
    st = [gen_seed]

    def u64():
        st[0], z = _splitmix64(st[0])
        return z

    def unif():
        return (u64() >> 11) * 2.0**-53

    def below(n):
        return u64() % n

    S = n_species
    by_mass = {m: [s for s in range(S) if 1 + s % 3 == m] for m in (1, 2, 3)}
    rx = []
    while len(rx) < n_reactions:
        if unif() < 0.4:
            m = 1 + below(3)
            pool = by_mass[m]
            X = pool[below(len(pool))]
            Y = pool[below(len(pool))]
            if X == Y:
                continue
            kf = 10.0 ** (-1.0 + 2.0 * unif())
            kb = 10.0 ** (-1.0 + 2.0 * unif())
            rx.append(_lma(S, kf, [X], [Y]))
            rx.append(_lma(S, kb, [Y], [X]))
        else:
            mx, my = (1, 1) if unif() < 0.5 else (1, 2)
            X = by_mass[mx][below(len(by_mass[mx]))]
            Y = by_mass[my][below(len(by_mass[my]))]
            Z = by_mass[mx + my][below(len(by_mass[mx + my]))]
            kf = 10.0 ** (-3.0 + 2.0 * unif())
            kb = 10.0 ** (-1.0 + 2.0 * unif())
            rx.append(_lma(S, kf, [X, Y], [Z]))
            rx.append(_lma(S, kb, [Z], [X, Y]))
    rx = rx[:n_reactions]
    x0 = [50 + below(101) for _ in range(S)]
    return dict(name="synthetic", species=["X%d" % s for s in range(S)], x0=x0, tmax=float(tmax), nb_steps=nb_steps,
                params=[r[0] for r in rx], reactions=rx)


MODELS = dict(sir=sir, dimers=dimers, vilar=vilar, mm_lma=mm_lma, synthetic=synthetic, ring=ring, flocculation=flocculation)


def build_network(model, arith=0):
    """Product-side network (C ABI) for a model dict."""
    from rebop_b200 import _ffi
    net = _ffi.Network(len(model["species"]), arith)
    for k, terms, diff in model["reactions"]:
        net.add_reaction_lma_sparse(k, terms, diff)
    return net


def seeds_sequence(n, first=0):
    """seed_i = first + i (the Rust benches seed every run explicitly, my_benchmark.rs:14-31)."""
    return np.arange(first, first + n, dtype=np.uint64)

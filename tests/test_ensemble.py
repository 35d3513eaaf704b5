"""Host side of the multi-GPU path, on CPU: sharding, exact statistics, the all-reduce under gloo (world_size 2).

The simulation data in the 2-rank test comes from the CPU oracle (this is a test; the product's
shards come from the GPU kernel) -- what is under test is rebop_b200.ensemble: the shard ranges,
the integer all-reduce and the mean/variance finalisation, i.e. everything rank-dependent.
"""
import os
import socket
import sys

import numpy as np
import pytest

from rebop_b200 import ensemble, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8, 1000, 10**7):
        for w in (1, 2, 3, 4, 8):
            r = ensemble.shard_ranges(n, w)
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            for (lo, c), (lo2, _) in zip(r, r[1:]):
                assert lo + c == lo2
            assert max(c for _, c in r) - min(c for _, c in r) <= 1
    assert ensemble.shard_range(10**7, 3, 8) == (3_750_000, 1_250_000)
    with pytest.raises(ValueError):
        ensemble.shard_range(10, 2, 2)


def test_finalize_stats_is_exact():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 5000, size=(6, 10000)).astype(np.int64)
    mean, var = ensemble.finalize_stats(x.sum(axis=1), (x * x).sum(axis=1), x.shape[1])
    np.testing.assert_allclose(mean, x.mean(axis=1), rtol=1e-15)
    np.testing.assert_allclose(var, x.var(axis=1, ddof=1), rtol=1e-12)
    # catastrophic-cancellation case: huge offset, tiny spread
    y = np.array([10**6, 10**6 + 1] * 5000, dtype=np.int64)
    m, v = ensemble.finalize_stats([int(y.sum())], [int((y * y).sum())], len(y))
    assert m[0] == 10**6 + 0.5 and abs(v[0] - 0.25 * len(y) / (len(y) - 1)) < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n_total, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from tests.helpers import oracle_network

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    model = models.sir()
    lo, cnt = ensemble.shard_range(n_total, rank, world)
    seeds = models.seeds_sequence(cnt, first=lo)  # trajectory n always uses seed n, whatever the sharding
    out, _, events = oracle_network(O, model).run_batch(model["x0"], seeds, 250.0, 25)
    o64 = out.astype(np.int64)
    part = np.concatenate([o64.sum(axis=2).ravel(), (o64 * o64).sum(axis=2).ravel()])
    t = torch.from_numpy(part.copy())
    ensemble.allreduce_sums(t)
    ev = torch.tensor([events], dtype=torch.int64)
    dist.all_reduce(ev)
    if rank == 0:
        np.save(os.path.join(tmpdir, "sums.npy"), t.numpy())
        np.save(os.path.join(tmpdir, "events.npy"), ev.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process(oracle, tmp_path):
    import torch.multiprocessing as mp

    from tests.helpers import oracle_network

    n_total, world = 301, 2  # odd: the shards differ in size
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    sums = np.load(tmp_path / "sums.npy")
    events = int(np.load(tmp_path / "events.npy")[0])

    model = models.sir()
    out, _, tot = oracle_network(oracle, model).run_batch(model["x0"], models.seeds_sequence(n_total), 250.0, 25)
    o64 = out.astype(np.int64)
    rows = 26 * 3
    np.testing.assert_array_equal(sums[:rows], o64.sum(axis=2).ravel())
    np.testing.assert_array_equal(sums[rows:], (o64 * o64).sum(axis=2).ravel())
    assert events == tot
    mean, var = ensemble.finalize_stats(sums[:rows], sums[rows:], n_total)
    np.testing.assert_allclose(mean.reshape(26, 3), o64.mean(axis=2), rtol=1e-14)
    np.testing.assert_allclose(var.reshape(26, 3), o64.var(axis=2, ddof=1), rtol=1e-10, atol=1e-12)
    # conservation survives the reduction: S + I + R = 1000 on average, at every sample time
    np.testing.assert_allclose(mean.reshape(26, 3).sum(axis=1), 1000.0, rtol=1e-14)

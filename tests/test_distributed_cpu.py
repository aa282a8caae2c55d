"""N > 1 host logic on CPU: world_size-2 gloo group (the data path itself has no collective besides the
single all-reduce of the reference-profile column sums)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infercnvpy_b200 import shard_rows
from infercnvpy_b200._engine import allreduce_sums


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, chunk, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = rng.poisson(0.3, size=(n_rows, 37)).astype(np.float32)
        cat = rng.integers(-1, 2, size=n_rows)  # -1 = not a reference cell
        r0, r1 = shard_rows(n_rows, chunk, rank, world)
        # what icnv_colsum_* produces for this shard: per-category float64 sums and int64 counts
        sums = np.stack([X[r0:r1][cat[r0:r1] == c].sum(axis=0, dtype=np.float64) for c in (0, 1)])
        counts = np.array([(cat[r0:r1] == c).sum() for c in (0, 1)], dtype=np.int64)
        s, c = allreduce_sums(torch.from_numpy(sums), torch.from_numpy(counts))
        want_s = np.stack([X[cat == k].sum(axis=0, dtype=np.float64) for k in (0, 1)])
        want_c = np.array([(cat == k).sum() for k in (0, 1)])
        np.testing.assert_allclose(s.numpy(), want_s, rtol=1e-13)
        np.testing.assert_array_equal(c.numpy(), want_c)
        # every rank ends up with the same profile, equal to the single-process mean
        ref = (s / c[:, None].to(torch.float64)).numpy()
        np.testing.assert_allclose(ref, np.stack([X[cat == k].mean(axis=0, dtype=np.float64) for k in (0, 1)]), rtol=1e-12)
        np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array([r0, r1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,chunk", [(1000, 128), (257, 100), (90, 100)])
def test_allreduce_and_sharding_world2(tmp_path, n_rows, chunk):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_rows, chunk, str(tmp_path)), nprocs=2, join=True)
    a, b = (np.load(tmp_path / f"rows_{r}.npy") for r in (0, 1))
    assert a[0] == 0 and a[1] == b[0] and b[1] == n_rows  # contiguous cover
    assert a[1] % chunk == 0 or a[1] == n_rows            # cut on a chunk boundary (tl/_infercnv.py:123)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_rows_partition(world):
    for n_rows, chunk in [(1_000_000, 5000), (100_000, 5000), (12_345, 5000), (3, 5000), (0, 5000)]:
        pieces = [shard_rows(n_rows, chunk, r, world) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == n_rows
        for (a0, a1), (b0, b1) in zip(pieces, pieces[1:]):
            assert a1 == b0 and a0 <= a1
        for r0, r1 in pieces:
            assert r0 % chunk == 0 or r0 == n_rows
        sizes = [(r1 - r0 + chunk - 1) // chunk for r0, r1 in pieces]
        assert max(sizes) - min(sizes) <= 1  # chunks are spread evenly


def test_single_process_is_identity():
    s, c = torch.ones(2, 3, dtype=torch.float64), torch.ones(2, dtype=torch.int64)
    s2, c2 = allreduce_sums(s, c)
    assert s2 is s and c2 is c


def _label_worker(rank, world, port, out_dir):
    """Host logic of the sharded cnv_score / reference-category validation: shards that see different,
    reordered or missing groups must agree on one label list, and a category missing on ONE rank is not an error."""
    import pandas as pd

    from infercnvpy_b200 import AnnData
    from infercnvpy_b200._engine import allreduce_host_counts, global_label_order
    from infercnvpy_b200.tl._infercnv import _reference_categories

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = [["b", "a", "b", "c"], ["c", "d", "a"]]  # rank 1 never sees "b", sees "c" first
        order = global_label_order(pd.unique(pd.Series(shards[rank])))
        assert order == ["b", "a", "c", "d"] == list(pd.unique(pd.Series(shards[0] + shards[1])))
        # per-label sums indexed by the global list line up across ranks
        codes = pd.Categorical(shards[rank], categories=order).codes
        local = np.bincount(codes, minlength=len(order)).astype(np.int64)
        total = allreduce_host_counts(local)
        np.testing.assert_array_equal(total, [2, 2, 2, 1])
        # reference categories: "T" only lives on rank 0 -> fine on both ranks; "zzz" nowhere -> ValueError on BOTH
        obs = pd.DataFrame({"ct": (["T", "x", "x"] if rank == 0 else ["x", "y", "x"])}, index=[f"c{i}" for i in range(3)])
        ad = AnnData(np.zeros((3, 4), dtype=np.float32), obs=obs)
        row_cat, cats = _reference_categories(ad, "ct", ["T", "y", "T"])
        assert list(cats) == ["T", "y"]  # the duplicate is dropped
        np.testing.assert_array_equal(row_cat, [0, -1, -1] if rank == 0 else [-1, 1, -1])
        with pytest.raises(ValueError, match="reference categories were not found"):
            _reference_categories(ad, "ct", ["T", "zzz"])
        open(os.path.join(out_dir, f"ok_{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_sharded_label_order_and_category_validation_world2(tmp_path):
    port = _free_port()
    mp.spawn(_label_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


def _gather_worker(rank, world, port, out_dir):
    """allgather_rows (pp/_neighbors.py): ragged row shards of coordinates / kNN lists are concatenated in rank order."""
    from infercnvpy_b200.pp._neighbors import allgather_rows

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        cut = [0, 5, 7]  # 5 rows on rank 0, 2 on rank 1
        mine = full[cut[rank] : cut[rank + 1]]
        got, row0 = allgather_rows(mine)
        assert row0 == cut[rank] and torch.equal(got, full)
        ids = torch.arange(cut[rank], cut[rank + 1], dtype=torch.int64)
        got_ids, _ = allgather_rows(ids)
        assert torch.equal(got_ids, torch.arange(7))
        empty, r0 = allgather_rows(full[:0] if rank == 1 else full)  # a rank without rows
        assert torch.equal(empty, full) and r0 == (0 if rank == 0 else 7)
        open(os.path.join(out_dir, f"g_{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_allgather_rows_world2(tmp_path):
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "g_0").exists() and (tmp_path / "g_1").exists()

"""N>1 host logic on CPU: world_size-2 gloo processes partition a gene list, "sample" their shard and
gather per-gene diagnostics on rank 0 — no overlap, full coverage, and a shard's content depends only on
the gene ids it holds (the RNG stream is keyed by gene id, csplotch_b200/csrc/philox.h)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csplotch_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sample(gene_id, seed=7):
    # stands in for the sampler: a pure function of (seed, gene id), like the Philox-keyed chains
    rng = np.random.default_rng([seed, gene_id])
    return {"ess": float(rng.uniform(100, 900)), "n_leapfrog": int(rng.integers(1000, 5000))}


def _worker(rank, world, port, genes, q, costs=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.my_genes(genes) if costs is None else shard.deal(genes, costs)   # bench.py's strong-scaling deal
    local = {g: _fake_sample(g) for g in mine}
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)                      # the barrier + reduction pattern bench.py uses
    allres = shard.gather_diagnostics(local)
    if rank == 0:
        q.put((allres, float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    genes = list(range(3, 26))              # 23 genes, 1-based ids like data_<g>.R
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, genes, q)) for r in range(2)]
    for p in procs:
        p.start()
    allres, total = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert total == len(genes)
    assert sorted(allres) == genes
    for g in genes:
        assert allres[g] == _fake_sample(g)  # identical whichever rank sampled it


def test_two_rank_gloo_cost_sorted_deal():
    """the strong-scaling path of bench.py (c5): one gene list dealt longest-first over the ranks, results keyed by gene id"""
    genes = list(range(3, 26))
    costs = [int(x) for x in np.random.default_rng(1).integers(0, 5000, len(genes))]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, genes, q, costs)) for r in range(2)]
    for p in procs:
        p.start()
    allres, total = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert total == len(genes) and sorted(allres) == genes
    for g in genes:
        assert allres[g] == _fake_sample(g)


def test_deal_properties():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 23, 592, 2368):
        ids = np.arange(100, 100 + n)
        cost = rng.integers(0, 120000, n)
        for w in (1, 2, 4, 8):
            parts = [shard.deal(ids, cost, w, r) for r in range(w)]
            assert sorted(sum(parts, [])) == ids.tolist()                       # a partition of the list
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
            if n >= 8 * w:
                tot = [sum(int(cost[g - 100]) for g in p) for p in parts]
                assert max(tot) - min(tot) <= 2 * cost.max()                   # snake deal: totals within two items
            for p in parts:                                                    # longest first inside a share
                c = [int(cost[g - 100]) for g in p]
                assert c == sorted(c, reverse=True)
            assert parts == [shard.deal(ids, cost, w, r) for r in range(w)]    # deterministic


def test_block_partition_properties():
    for n in (0, 1, 7, 8, 23, 20000, 32768):
        for w in (1, 2, 4, 8):
            blocks = [shard.block_partition(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_contiguous_runs():
    assert shard.contiguous_runs([1, 2, 3, 7, 8, 10], 2) == [[1, 2], [3], [7, 8], [10]]
    assert shard.contiguous_runs([], 4) == []

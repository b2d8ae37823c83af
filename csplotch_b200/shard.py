"""Gene sharding across the GPUs of one box.

Genes never interact, chains of a gene stay on one GPU, and every random number is a function of
(seed, gene id, chain id, iteration, purpose), so the path shards with NO data-path collective: each
rank (one process per GPU, launched by torchrun) takes a contiguous block of the gene list, samples it,
writes its own `OUT/<g//100>/combined_<g>.*` files, and only small per-gene diagnostics are gathered on
rank 0 for the report (`torch.distributed.gather_object`; gloo on CPU in the tests, nccl on the box).
"""
from __future__ import annotations

import os


def block_partition(n_items: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def my_genes(gene_ids, world: int | None = None, rank: int | None = None):
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    lo, hi = block_partition(len(gene_ids), world, rank)
    return list(gene_ids[lo:hi])


def deal(gene_ids, costs, world: int | None = None, rank: int | None = None):
    """Cost-balanced share of rank: the gene list sorted by expected cost (longest first; ties by id) and dealt
    round-robin in snake order (0..w-1, w-1..0, ...), so every rank gets the same number of genes (+-1) and nearly
    the same total cost.  The usual cost is the gene's number of non-zero counts: the tree depth of a sparse gene's
    posterior is lower.  Deterministic, and the union over ranks is the whole list (tests/test_shard_gloo.py).
    Results do not depend on the deal: pass the returned ids to api.Batch(gene_ids=...)."""
    import numpy as np
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    ids = np.asarray(gene_ids)
    c = np.asarray(costs, dtype=np.float64)
    if ids.shape[0] != c.shape[0]:
        raise ValueError("one cost per gene")
    order = np.lexsort((ids, -c))
    pos = np.arange(len(order))
    lap, col = pos // world, pos % world
    owner = np.where(lap % 2 == 0, col, world - 1 - col)
    return [int(g) if isinstance(g, (int, np.integer)) else g for g in ids[order[owner == rank]].tolist()]


def contiguous_runs(gene_ids, max_len: int):
    """Split a sorted gene-id list into runs of consecutive ids (<= max_len): one device batch each, so
    that `gene_id0 + i` is the gene id that keys the RNG stream."""
    runs, cur = [], []
    for g in gene_ids:
        if cur and (g != cur[-1] + 1 or len(cur) >= max_len):
            runs.append(cur)
            cur = []
        cur.append(g)
    if cur:
        runs.append(cur)
    return runs


def gather_diagnostics(local: dict, group=None):
    """Collect {gene_id: diagnostics} dictionaries of all ranks on rank 0 (None elsewhere)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    out = {}
    for part in bucket:
        overlap = set(out) & set(part)
        if overlap:
            raise RuntimeError("genes sampled by more than one rank: %s" % sorted(overlap)[:5])
        out.update(part)
    return out

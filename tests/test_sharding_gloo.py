"""CPU, world_size 2, gloo: the N>1 host logic (shard bounds, logit gather order, counter all-reduce).
The per-clip forward needs no collective, so a deterministic stand-in 'model' (a fixed linear map) is enough here;
the real multi-GPU forward is covered by tests/test_model_gpu.py::test_batch_sharding_equivalence and bench.py."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gava_clip_b200.sharding import gather_logits, reduce_hit_counts, shard_batch, shard_bounds


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 5, 64, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        clips = torch.randn(n_items, 6, generator=g)
        w = torch.randn(6, 3, generator=g)
        labels = torch.randint(0, 3, (n_items,), generator=g)
        mine = shard_batch(clips)
        lo, hi = shard_bounds(n_items, rank, world)
        logits = mine @ w
        full = gather_logits(logits, n_items)
        tot, hit = reduce_hit_counts(logits, labels[lo:hi])
        ref = clips @ w
        ok = torch.allclose(full, ref) and tot == n_items and hit == int((ref.argmax(-1) == labels).sum())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_two_rank_gather_and_counters(n_items):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000) + n_items
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True

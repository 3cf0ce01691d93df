"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in adgs_b200/parallel.py:
view sharding, the flat gradient bucket + single all-reduce, and the densification statistics
(sum of per-view norms, visibility counts, max radii) -- the oracle being the sum / max over all
views computed in one process."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adgs_b200.parallel import DensifyStats, FlatGradBucket, shard_views


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_view_grads(view, n):
    g = torch.Generator().manual_seed(100 + view)
    return {"xyz": torch.randn(n, 3, generator=g), "sh4": torch.randn(12, n, 4, generator=g),
            "background_deform": torch.randn(3, 5, generator=g), "opacity": torch.randn(n, 1, generator=g)}


def _fake_view_stats(view, n):
    g = torch.Generator().manual_seed(200 + view)
    return (torch.randn(n, 3, generator=g), torch.rand(n, generator=g) > 0.4,
            torch.randint(0, 50, (n,), generator=g, dtype=torch.int32))


def _worker(rank, world, port, n, n_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = {k: torch.zeros_like(v) for k, v in _fake_view_grads(0, n).items()}
    bucket = FlatGradBucket(params)
    stats = DensifyStats(n, "cpu")
    for view in shard_views(list(range(n_views)), rank, world):
        bucket.accumulate(_fake_view_grads(view, n))
        stats.add_view(*_fake_view_stats(view, n))
    bucket.all_reduce()
    stats.all_reduce()
    if rank == 0:
        torch.save({"grads": {k: v.clone() for k, v in bucket.views().items()},
                    "norm": stats.grad_norm_sum, "count": stats.visible_count, "radii": stats.max_radii}, out)
    dist.destroy_process_group()


def test_shard_views_round_robin():
    views = list(range(8))
    assert shard_views(views, 0, 2) == [0, 2, 4, 6] and shard_views(views, 1, 2) == [1, 3, 5, 7]
    assert sorted(sum((shard_views(views, r, 4) for r in range(4)), [])) == views
    assert shard_views(views[:3], 3, 4) == []


def test_flat_bucket_layout():
    params = {"a": torch.zeros(5, 3), "b": torch.zeros(7), "c": torch.zeros(2, 2, 4)}
    b = FlatGradBucket(params)
    v = b.views()
    assert all(v[k].shape == params[k].shape for k in params)
    assert all(v[k].data_ptr() % 16 == b.flat.data_ptr() % 16 for k in params)
    v["b"].fill_(3.0)
    assert b.flat.sum().item() == 21.0
    b.accumulate({"a": torch.ones(5, 3), "b": None, "c": torch.ones(2, 2, 4)})
    assert b.flat.sum().item() == 21.0 + 15 + 16


def test_two_rank_all_reduce_equals_sum_over_views(tmp_path):
    n, n_views, world = 300, 5, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, n_views, out), nprocs=world, join=True)
    got = torch.load(out)
    want = {k: sum(_fake_view_grads(v, n)[k] for v in range(n_views)) for k in got["grads"]}
    for k in want:
        assert torch.allclose(got["grads"][k], want[k], atol=1e-5), k
    ref = DensifyStats(n, "cpu")
    for v in range(n_views):
        ref.add_view(*_fake_view_stats(v, n))
    assert torch.allclose(got["norm"], ref.grad_norm_sum, atol=1e-5)
    assert torch.equal(got["count"], ref.visible_count)
    assert torch.equal(got["radii"], ref.max_radii)

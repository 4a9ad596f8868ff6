"""N > 1 host logic of the keyframe-sharded mapping window on CPU: world_size-2 (and 3) gloo process groups.
The CUDA engine cannot run here, so every rank fabricates deterministic per-view gradients; what is under
test is the owner map, the accumulate-then-one-all-reduce protocol and that ranks without views still take
part in the collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_grad(v, n):
    g = torch.Generator().manual_seed(1000 + v)
    return torch.randn(n, generator=g, dtype=torch.float32)


def _worker(rank, world, port, V, n, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import allreduce_window_gradients, shard_views

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    views = shard_views(V, world, rank)
    flat = torch.zeros(n)
    for i, v in enumerate(views):            # first view overwrites, the rest accumulate (engine protocol)
        flat = _view_grad(v, n) if i == 0 else flat + _view_grad(v, n)
    allreduce_window_gradients(flat)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), flat.numpy())
    np.save(os.path.join(out_dir, "v%d.npy" % rank), np.asarray(views, np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,V", [(2, 10), (3, 2)])
def test_window_allreduce_gloo(tmp_path, world, V):
    n = 4099
    mp.spawn(_worker, args=(world, _free_port(), V, n, str(tmp_path)), nprocs=world, join=True)
    expect = sum(_view_grad(v, n) for v in range(V)).numpy()
    seen = []
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npy" % r))
        np.testing.assert_allclose(got, expect, rtol=1e-5, atol=1e-5)
        seen += list(np.load(tmp_path / ("v%d.npy" % r)))
    assert sorted(seen) == list(range(V))          # every view rendered exactly once


def test_shard_map():
    sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))
    from diff_gaussian_rasterization.window import owner_of, shard_views

    assert [len(shard_views(10, 8, r)) for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]
    assert [len(shard_views(32, 8, r)) for r in range(8)] == [4] * 8
    assert shard_views(3, 4, 3) == []
    for v in range(10):
        assert v in shard_views(10, 8, owner_of(v, 8))

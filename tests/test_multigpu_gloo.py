"""View-sharding host logic on CPU: world_size 2, gloo backend (the NCCL path differs only in the backend)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, num_views, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from gaustudio_b200 import parallel
    r, lr, w = parallel.init_distributed(backend="gloo")
    views = parallel.shard_views(num_views, r, w)
    local = torch.tensor([float(v) * 10 + 1 for v in views])  # "loss" of view v = 10 v + 1
    full = parallel.gather_view_losses(local, num_views, r, w)
    ms = parallel.barrier_max_ms(5.0 + r, torch.device("cpu"))
    q.put((r, views, full.tolist(), ms))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_views", [8, 7])
def test_view_sharding_world2(num_views):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, num_views, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, v0, f0, ms0), (r1, v1, f1, ms1) = res
    assert sorted(v0 + v1) == list(range(num_views)) and not set(v0) & set(v1)
    assert v0 == list(range(0, num_views, 2)) and v1 == list(range(1, num_views, 2))
    expect = [10.0 * v + 1 for v in range(num_views)]
    assert f0 == expect and f1 == expect          # every rank sees every view's loss, in view order
    assert ms0 == ms1 == 6.0                      # max over ranks


def test_single_process_paths():
    from gaustudio_b200 import parallel
    assert parallel.shard_views(5, 0, 1) == [0, 1, 2, 3, 4]
    t = torch.arange(5.0)
    assert parallel.gather_view_losses(t, 5, 0, 1) is t
    assert parallel.barrier_max_ms(3.0, torch.device("cpu")) == 3.0

"""Optimizer plugin surface (gaustudio/pipelines/optimizers) and the data-parallel gradient bucket: host logic on CPU,
world_size 2 over gloo.  The step kernel itself is covered by tests/test_gpu_optim.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VANILLA = {"xyz": {"lr": 0.00016}, "opacity": {"lr": 0.05}, "f_dc": {"lr": 0.0025}, "scale": {"lr": 0.005},
           "rot": {"lr": 0.001}}  # configs/vanilla.yaml:36-46


def test_registry_and_groups():
    from gaustudio_b200 import optimizers
    from gaustudio_b200.synthetic import make_scene
    with pytest.raises(ValueError):
        optimizers.make({})
    with pytest.raises(ValueError):
        optimizers.make("no_such_optimizer")
    m = make_scene(50, 1.0, 0.03, 1)
    opt = optimizers.make({"name": "general", "model": m, "optimizer_name": "AdamW", "args": {"lr": 0.0, "eps": 1e-15},
                           "params": VANILLA})
    groups = opt._optimizer.param_groups
    assert [g["name"] for g in groups] == list(VANILLA) and [g["lr"] for g in groups] == [v["lr"] for v in VANILLA.values()]
    assert all(g["eps"] == 1e-15 and g["weight_decay"] == 0.01 and g["betas"] == (0.9, 0.999) for g in groups)  # torch AdamW defaults
    assert isinstance(m._xyz, torch.nn.Parameter) and m._xyz.requires_grad and not m._f_rest.requires_grad
    assert opt._optimizer.decoupled
    m._xyz.sum().backward()
    with pytest.raises(RuntimeError):  # the step has no CPU path
        opt.step()
    opt.zero_grad()
    assert m._xyz.grad.abs().max() == 0
    # other optimizer names go to torch.optim like the reference
    m2 = make_scene(10, 1.0, 0.03, 1)
    sgd = optimizers.make({"name": "general", "model": m2, "optimizer_name": "SGD", "args": {"lr": 0.1},
                           "params": {"xyz": {"lr": 0.5}}})
    assert isinstance(sgd._optimizer, torch.optim.SGD)
    before = m2._xyz.detach().clone()
    m2._xyz.sum().backward(); sgd.step()
    assert torch.allclose(m2._xyz.detach(), before - 0.5)


def test_bucket_aliases_grads():
    from gaustudio_b200.parallel import GradBucket
    a = torch.nn.Parameter(torch.randn(5, 3)); b = torch.nn.Parameter(torch.randn(7))
    bk = GradBucket([a, b])
    assert bk.flat.numel() == 16 + 8 and bk.world == 1 and bk.grad_scale == 1.0 and bk.all_reduce() is None
    assert a.grad.data_ptr() == bk.flat.data_ptr() and (b.grad.data_ptr() - bk.flat.data_ptr()) == 16 * 4
    (a.sum() * 2 + (b * b).sum()).backward()
    assert a.grad.data_ptr() == bk.flat.data_ptr()  # autograd accumulated in place
    assert torch.equal(bk.flat[:15], torch.full((15,), 2.0)) and torch.allclose(bk.flat[16:23], 2 * b.detach())
    bk.zero()
    assert a.grad.abs().max() == 0 and b.grad.abs().max() == 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from gaustudio_b200 import parallel
    parallel.init_distributed(backend="gloo")
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(6, 3)); b = torch.nn.Parameter(torch.randn(5))
    bk = parallel.GradBucket([a, b])
    # each rank's "views" give a different gradient: d/da sum((rank+1) a) = rank+1, d/db sum(b^2 (rank+1)) = 2 b (rank+1)
    ((rank + 1) * a.sum() + (rank + 1) * (b * b).sum()).backward()
    bk.all_reduce(async_op=True); bk.wait()
    q.put((rank, bk.grad_scale, a.grad.clone(), b.grad.clone(), b.detach().clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_all_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, scale, ga, gb, b in res:
        assert scale == 0.5
        assert torch.equal(ga, torch.full((6, 3), 3.0))          # 1 + 2, summed over the ranks
        assert torch.allclose(gb, 6.0 * b)                        # (1 + 2) * 2 b
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])


def test_fused_adam_state_dict_roundtrip():
    """Checkpoint / resume of the optimizer state (host-side bookkeeping only; the step kernel needs a GPU)."""
    from gaustudio_b200.optimizers import FusedAdam
    a = torch.nn.Parameter(torch.randn(4, 3)); b = torch.nn.Parameter(torch.randn(5))
    opt = FusedAdam([{"params": [a], "lr": 0.1}, {"params": [b], "lr": 0.2, "betas": [0.9, 0.999]}], eps=1e-15)
    assert opt.param_groups[1]["betas"] == (0.9, 0.999)
    # per-parameter step counts, like torch (b received gradients on fewer steps than a)
    opt.state[a] = {"step": 7, "exp_avg": torch.ones(4, 3), "exp_avg_sq": torch.full((4, 3), 2.0)}
    opt.state[b] = {"step": 5, "exp_avg": torch.zeros(5), "exp_avg_sq": torch.zeros(5)}
    sd = opt.state_dict()
    assert sd["param_groups"][0]["params"] == [0] and sd["param_groups"][1]["params"] == [1]
    assert sd["state"][0]["step"] == 7 and sd["state"][1]["step"] == 5
    a2 = torch.nn.Parameter(torch.randn(4, 3)); b2 = torch.nn.Parameter(torch.randn(5))
    new = FusedAdam([{"params": [a2], "lr": 1.0}, {"params": [b2], "lr": 1.0}])
    new.load_state_dict(sd)
    assert new.step_count == 7 and new.state[b2]["step"] == 5
    assert new.param_groups[0]["lr"] == 0.1 and new.param_groups[1]["lr"] == 0.2
    assert new.param_groups[0]["eps"] == 1e-15 and torch.equal(new.state[a2]["exp_avg_sq"], torch.full((4, 3), 2.0))
    assert new.state[a2]["exp_avg"].data_ptr() != opt.state[a]["exp_avg"].data_ptr()
    with pytest.raises(ValueError):
        FusedAdam([a2]).load_state_dict(sd)
    with pytest.raises(ValueError):
        FusedAdam([{"params": [a]}, {"params": [b], "eps": 1e-8}], eps=1e-15)  # one launch: one eps
    with pytest.raises(ValueError, match="empty parameter list"):
        FusedAdam([])  # torch.optim raises in the same case


def test_torch_adam_checkpoint_with_uneven_steps_loads():
    """A torch.optim.AdamW checkpoint whose parameters have different step counts is a valid input."""
    from gaustudio_b200.optimizers import FusedAdam
    a = torch.nn.Parameter(torch.randn(4, 3)); b = torch.nn.Parameter(torch.randn(5))
    ref = torch.optim.AdamW([a, b], lr=0.1)
    a.grad = torch.randn(4, 3); b.grad = torch.randn(5); ref.step()
    b.grad = None; a.grad = torch.randn(4, 3); ref.step()      # only `a` moves on the second step
    sd = ref.state_dict()
    sd["decoupled"] = True
    opt = FusedAdam([a, b], lr=0.1)
    opt.load_state_dict(sd)
    assert opt.state[a]["step"] == 2 and opt.state[b]["step"] == 1


def test_point_cloud_container_is_parameter_aware():
    """`GeneralOptimizer` wraps the model attributes as nn.Parameters; .to() / .requires_grad_() must keep working."""
    from gaustudio_b200 import optimizers
    from gaustudio_b200.synthetic import make_scene
    m = make_scene(50, 1.0, 0.05, 0)
    with pytest.raises(ValueError, match="empty parameter list"):
        optimizers.make({"name": "general", "model": m, "optimizer_name": "AdamW", "args": {"lr": 1e-3}})
    optimizers.make({"name": "general", "model": m, "optimizer_name": "SGD", "args": {"lr": 1e-3},
                     "params": {"xyz": {"lr": 1e-4}, "opacity": {"lr": 5e-2}}})
    assert isinstance(m._xyz, torch.nn.Parameter)
    m.requires_grad_(False)
    assert isinstance(m._xyz, torch.nn.Parameter) and not m._xyz.requires_grad and not m._scale.requires_grad
    m.to("cpu").requires_grad_(True)
    assert m._xyz.is_leaf and m._xyz.requires_grad and m._scale.is_leaf

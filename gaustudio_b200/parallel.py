"""View-sharded data parallelism: the rasterizer path has no cross-view state, so ranks take
disjoint views (k mod world == rank), keep Gaussian-parameter gradients local, and exchange only the
per-view loss scalars with ONE all_gather at the end of the pass.  One process per GPU (torchrun); NCCL on
GPUs, gloo in the CPU tests.  The reference has no multi-GPU code (render_gs.py:34-36 only sets
CUDA_VISIBLE_DEVICES)."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_views(num_views, rank, world):
    """Indices of the views rank `rank` renders: k ≡ rank (mod world)."""
    return list(range(rank, num_views, world))


def gather_view_losses(local_losses, num_views, rank, world):
    """local_losses: 1-D tensor of this rank's per-view losses in shard order -> 1-D tensor of all
    `num_views` losses in view order on every rank.  One all_gather (ragged shards are padded)."""
    if world == 1:
        return local_losses
    per = (num_views + world - 1) // world
    pad = torch.full((per,), float("nan"), dtype=local_losses.dtype, device=local_losses.device)
    pad[: local_losses.numel()] = local_losses
    out = torch.empty(world * per, dtype=local_losses.dtype, device=local_losses.device)
    dist.all_gather_into_tensor(out, pad) if hasattr(dist, "all_gather_into_tensor") and out.is_cuda else \
        _all_gather_list(out, pad, world)
    full = out.view(world, per).t().reshape(-1)  # view k lives at [k % world][k // world]
    return full[:num_views]


def _all_gather_list(out, pad, world):
    chunks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(chunks, pad)
    out.copy_(torch.cat(chunks))


def barrier_max_ms(ms, device):
    """max over ranks of a device-measured duration (ms)."""
    if not (dist.is_available() and dist.is_initialized()):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class GradBucket:
    """Every parameter's `.grad` as a view into ONE flat float32 buffer, so data-parallel training needs a single
    all_reduce of 59 P floats per step -- no per-tensor collectives, no flatten copies: autograd
    accumulates straight into the views.  The sum is left un-normalised; the optimizer kernel folds the 1/world in
    (FusedAdam.step(grad_scale=bucket.grad_scale, zero_grad=True) also re-zeroes the bucket in the same pass)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params]
        self.group = group
        pad4 = lambda k: (k + 3) // 4 * 4  # every view starts 16-byte aligned: float4 path of the optimizer kernel
        n = sum(pad4(p.numel()) for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(n, dtype=first.dtype, device=first.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += pad4(p.numel())
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.grad_scale = 1.0 / self.world
        self._work = None

    def all_reduce(self, async_op=True):
        """Sum the bucket over the ranks (NCCL on its own stream when async: overlap it with the next forward)."""
        if self.world == 1:
            return None
        self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return self._work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    def zero(self):
        self.flat.zero_()

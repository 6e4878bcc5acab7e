"""Optimizer plugin surface of the reference (gaustudio/pipelines/optimizers/{__init__,base,general_optimizer}.py)
with the step itself on the device in ONE launch.

    opt = optimizers.make({"name": "general", "model": pcd, "optimizer_name": "AdamW",
                           "args": {"lr": 0.0, "eps": 1e-15},
                           "params": {"xyz": {"lr": 1.6e-4}, "opacity": {"lr": 0.05}, ...}})   # configs/vanilla.yaml:30-46
    loss.backward(); opt.step(); opt.zero_grad()

`Adam` / `AdamW` run through `FusedAdam` (gsr_adam_step, include/gsr.h): every parameter group in one kernel,
optionally with the gradient averaging of data-parallel training (`grad_scale`) and `zero_grad` fused in.  Any other
`optimizer_name` is handed to `torch.optim` exactly as the reference does.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib

optimizers = {}  # name -> class (public, like the reference's module-level dict)


def register(name):
    def _add(cls):
        optimizers[name] = cls
        return cls
    return _add


def make(config):
    """`config`: a registered name, or a mapping with `name` (the mapping itself is handed to the constructor)."""
    name, options = (config, {}) if isinstance(config, str) else (config.get("name"), config)
    if not name:
        raise ValueError("Optimizer name is required")
    try:
        cls = optimizers[name]
    except KeyError:
        raise ValueError(f"Unknown optimizer: {name}") from None
    return cls(options)


class _AdamGroup(C.Structure):  # gsr_adam_group, include/gsr.h
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("lr", C.c_float), ("weight_decay", C.c_float), ("grad2", C.c_void_p)]


MAX_GROUPS = 16


class FusedAdam:
    """torch.optim.Adam / AdamW semantics (no amsgrad, no maximize) with the whole step in one CUDA launch.
    `param_groups` follows torch: an iterable of tensors or of dicts {'params': tensor | [tensors], 'lr': ..., ...}."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=None, decoupled=True, **unsupported):
        if unsupported.get("amsgrad") or unsupported.get("maximize"):
            raise ValueError("FusedAdam: amsgrad / maximize are not supported")
        if weight_decay is None:
            weight_decay = 1e-2 if decoupled else 0.0  # torch defaults: AdamW 0.01, Adam 0
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.decoupled = bool(decoupled)
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        self.param_groups = []
        for g in params:
            g = dict(g)
            ps = g["params"]
            g["params"] = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            g["betas"] = tuple(g["betas"])
            self.param_groups.append(g)
        betas_eps = {(g["betas"], g["eps"]) for g in self.param_groups}
        if len(betas_eps) > 1:
            raise ValueError("FusedAdam: betas and eps must be the same for every group (one launch)")
        if not any(g["params"] for g in self.param_groups):
            raise ValueError("optimizer got an empty parameter list")  # torch.optim's message
        self.state = {}

    def _tensors(self):
        for g in self.param_groups:
            for p in g["params"]:
                yield g, p

    def step(self, grad_scale=1.0, zero_grad=False, closure=None, extra_grads=None):
        """One optimizer step over every parameter that has a gradient.  grad_scale multiplies the gradients first
        (1/world_size after a summing all-reduce); zero_grad=True leaves the gradients zeroed by the same kernel.
        extra_grads: optional list (aligned with the parameters in group order, entries may be None) of second
        gradient tensors -- e.g. the second all-reduce bucket of a data-parallel step -- added to `.grad` inside the
        kernel.  Like torch, every parameter keeps its own step count (bias correction of a parameter that only
        sometimes receives a gradient); parameters that share a count move in one launch."""
        if closure is not None:
            raise ValueError("FusedAdam: closures are not supported")
        by_step, keep = {}, []
        dev = None
        extra = list(extra_grads) if extra_grads is not None else None
        for k, (g, p) in enumerate(self._tensors()):
            if p.grad is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters (there is no CPU path)")
            grad = p.grad
            if not grad.is_contiguous():
                grad = p.grad = grad.contiguous()
            g2 = extra[k] if extra is not None and k < len(extra) else None
            if g2 is not None and (g2.shape != p.shape or not g2.is_contiguous() or g2.dtype != torch.float32):
                raise RuntimeError("FusedAdam: extra_grads entries must be contiguous float32 tensors shaped like their parameter")
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p, memory_format=torch.contiguous_format),
                                      "exp_avg_sq": torch.zeros_like(p, memory_format=torch.contiguous_format)}
            st["step"] += 1
            dev = p.device
            by_step.setdefault(st["step"], []).append(
                _AdamGroup(p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                           float(g["lr"]), float(g["weight_decay"]), g2.data_ptr() if g2 is not None else None))
            keep.append((grad, g2))
        if not by_step:
            return
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        L = _lib.lib()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for step, rows in by_step.items():
                for i in range(0, len(rows), MAX_GROUPS):
                    chunk = rows[i:i + MAX_GROUPS]
                    arr = (_AdamGroup * len(chunk))(*chunk)
                    rc = L.gsr_adam_step(len(chunk), C.cast(arr, C.c_void_p), float(b1), float(b2), float(eps), int(step),
                                         int(self.decoupled), float(grad_scale), int(bool(zero_grad)), stream)
                    if rc < 0:
                        raise RuntimeError("gsr_adam_step failed: " + _lib.last_error())

    @property
    def step_count(self):
        """Largest per-parameter step count (all equal when every parameter gets a gradient every step)."""
        return max((st["step"] for st in self.state.values()), default=0)

    def state_dict(self):
        """torch.optim layout: parameters are numbered in group order; state holds step / exp_avg / exp_avg_sq."""
        index = {id(p): i for i, (_, p) in enumerate(self._tensors())}
        groups = []
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = [index[id(p)] for p in g["params"]]
            groups.append(d)
        state = {index[id(p)]: {"step": st["step"], "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}
                 for p, st in self.state.items()}
        return {"state": state, "param_groups": groups, "decoupled": self.decoupled}

    def load_state_dict(self, sd):
        """Resume (also from a torch.optim.Adam/AdamW checkpoint): moments are copied onto the current parameters'
        devices, per-parameter step counts are kept; hyper-parameters come from the checkpoint."""
        params = [p for _, p in self._tensors()]
        if [len(g["params"]) for g in sd["param_groups"]] != [len(g["params"]) for g in self.param_groups]:
            raise ValueError("FusedAdam.load_state_dict: parameter groups do not match")
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})
            g["betas"] = tuple(g["betas"])
        self.state = {}
        for i, st in sd["state"].items():
            p = params[int(i)]
            if st["exp_avg"].shape != p.shape:
                raise ValueError("FusedAdam.load_state_dict: moment shape does not match its parameter")
            self.state[p] = {"step": int(st["step"]),
                             "exp_avg": st["exp_avg"].detach().to(device=p.device, dtype=torch.float32).contiguous().clone(),
                             "exp_avg_sq": st["exp_avg_sq"].detach().to(device=p.device, dtype=torch.float32).contiguous().clone()}
        self.decoupled = bool(sd.get("decoupled", self.decoupled))

    def zero_grad(self, set_to_none=False):
        """Keeps the gradient tensors by default (the flat all-reduce bucket of parallel.GradBucket aliases them)."""
        for _, p in self._tensors():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()


class BaseOptimizer:
    """gaustudio/pipelines/optimizers/base.py:7-34."""

    def __init__(self, config, **kwargs):
        self.config = config
        self.model = config["model"]
        self._initialize_internal_state()
        self.setup_optimizer()

    def setup_optimizer(self):
        name = self.config["optimizer_name"]
        args = dict(self.config.get("args") or {})
        if name in ("Adam", "AdamW"):
            self._optimizer = FusedAdam(self.param_groups, decoupled=(name == "AdamW"), **args)
        else:
            self._optimizer = getattr(torch.optim, name)(self.param_groups, **args)

    def _initialize_internal_state(self):
        raise NotImplementedError

    def step(self, **kw):
        self._optimizer.step(**kw)

    def zero_grad(self):
        self._optimizer.zero_grad()


@register("general")
class GeneralOptimizer(BaseOptimizer):
    """gaustudio/pipelines/optimizers/general_optimizer.py:9-21: one group per entry of config['params'] (the model
    attribute `_<name>` becomes an nn.Parameter), else every model parameter in one group."""

    def _initialize_internal_state(self):
        spec = self.config.get("params") if hasattr(self.config, "get") else None
        if spec is None:
            self.param_groups = self.model.parameters()
            return
        self.param_groups = []
        for name, options in spec.items():
            attr = "_" + name
            leaf = nn.Parameter(getattr(self.model, attr), requires_grad=True)
            setattr(self.model, attr, leaf)
            self.param_groups.append(dict(options or {}, params=leaf, name=name))

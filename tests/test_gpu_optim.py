"""-m gpu: the fused Adam / AdamW step (gsr_adam_step) against torch.optim on the same seeded gradients, and one
render -> backward -> step of the reference's optimizer config through the plugin surface."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).to(DEV) for s in shapes]


@pytest.mark.parametrize("name", ["AdamW", "Adam"])
def test_fused_adam_matches_torch(name):
    from gaustudio_b200.optimizers import FusedAdam
    shapes = [(1000, 3), (1000, 15, 3), (1000, 1), (333,), (7, 5), (1,)]  # float4 bodies, ragged tails, scalar path
    lrs = [1.6e-4, 2.5e-3, 0.05, 5e-3, 1e-3, 0.1]
    wd = [0.01, 0.0, 0.1, 0.01, 0.02, 0.0]
    ours = [torch.nn.Parameter(p.clone()) for p in _params(1, shapes)]
    ref = [torch.nn.Parameter(p.clone()) for p in _params(1, shapes)]
    mk = lambda ps: [{"params": [p], "lr": lr, "weight_decay": w} for p, lr, w in zip(ps, lrs, wd)]
    fo = FusedAdam(mk(ours), lr=0.0, eps=1e-15, decoupled=(name == "AdamW"))
    to = getattr(torch.optim, name)(mk(ref), lr=0.0, eps=1e-15)
    for step in range(6):
        grads = _params(100 + step, shapes)
        for p, q, g in zip(ours, ref, grads):
            p.grad = g.clone() * (1.0 if step != 3 else 1e-6)  # one step with tiny gradients (eps 1e-15 regime)
            q.grad = p.grad.clone()
        fo.step(); to.step()
        for i, (p, q) in enumerate(zip(ours, ref)):
            err = (p - q).abs().max().item()
            assert err <= 2e-6 * max(1.0, q.abs().max().item()), (name, step, i, err)
    for p, q in zip(ours, ref):
        st, rt = fo.state[p], to.state[q]
        # the moments are sums with cancellation: absolute tolerance at float32 epsilon of the O(1) gradients
        assert torch.allclose(st["exp_avg"], rt["exp_avg"], rtol=1e-5, atol=5e-7)
        assert torch.allclose(st["exp_avg_sq"], rt["exp_avg_sq"], rtol=1e-5, atol=1e-8)


def test_fused_scale_and_zero_grad():
    from gaustudio_b200.optimizers import FusedAdam
    a = torch.nn.Parameter(torch.randn(4097, device=DEV)); b = torch.nn.Parameter(a.detach().clone())
    g = torch.randn(4097, device=DEV)
    fa = FusedAdam([a], lr=1e-2, eps=1e-8); fb = FusedAdam([b], lr=1e-2, eps=1e-8)
    a.grad = g.clone() * 4; b.grad = g.clone()
    fa.step(grad_scale=0.25, zero_grad=True); fb.step()
    assert torch.allclose(a, b, rtol=0, atol=1e-7) and a.grad.abs().max() == 0 and b.grad.abs().max() > 0
    c = torch.nn.Parameter(torch.randn(8, device=DEV))  # no gradient -> untouched, no state
    fc = FusedAdam([c]); before = c.detach().clone(); fc.step()
    assert torch.equal(c.detach(), before) and fc.step_count == 0
    with pytest.raises(RuntimeError):
        d = torch.nn.Parameter(torch.randn(8)); d.grad = torch.ones(8); FusedAdam([d]).step()
    with pytest.raises(ValueError):
        FusedAdam([c], amsgrad=True)
    many = [torch.nn.Parameter(torch.randn(10, device=DEV)) for _ in range(20)]  # > 16 groups: two launches
    ref = [torch.nn.Parameter(p.detach().clone()) for p in many]
    for p, q in zip(many, ref):
        p.grad = torch.ones_like(p); q.grad = torch.ones_like(q)
    FusedAdam(many, lr=1e-2).step(); torch.optim.AdamW(ref, lr=1e-2).step()
    assert all(torch.allclose(p, q, atol=1e-6) for p, q in zip(many, ref))


def test_training_step_through_plugins():
    """configs/vanilla.yaml optimizer block on a synthetic scene: render, L1 loss, backward, fused step."""
    from gaustudio_b200 import optimizers, renderers
    from gaustudio_b200.parallel import GradBucket
    from gaustudio_b200.synthetic import build_config
    model, cams, c = build_config("cfg1", P=4000, W=128, H=96, K=4)
    dev = torch.device(DEV); model.to(dev)
    opt = optimizers.make({"name": "general", "model": model, "optimizer_name": "AdamW", "args": {"lr": 0.0, "eps": 1e-15},
                           "params": {"xyz": {"lr": 0.00016}, "opacity": {"lr": 0.05}, "f_dc": {"lr": 0.0025},
                                      "scale": {"lr": 0.005}, "rot": {"lr": 0.001}}})
    trained = [model._xyz, model._opacity, model._f_dc, model._scale, model._rot]
    bucket = GradBucket(trained)
    before = [p.detach().clone() for p in trained]; rest = model._f_rest.clone()
    r = renderers.make({"name": "vanilla_renderer"})
    target = torch.rand(3, 96, 128, device=dev)
    losses = []
    for it in range(8):
        out = r.render(cams[it % 4].to(dev), model)
        loss = (out["render"] - target).abs().mean()
        loss.backward()
        assert model._xyz.grad.data_ptr() == bucket.flat.data_ptr() and bucket.flat.abs().max() > 0
        opt.step(grad_scale=bucket.grad_scale, zero_grad=True)
        assert bucket.flat.abs().max() == 0
        losses.append(float(loss))
    assert all(torch.isfinite(p).all() for p in trained)
    assert all((p.detach() - b).abs().max() > 0 for p, b in zip(trained, before)) and torch.equal(model._f_rest, rest)
    assert sum(losses[4:]) < sum(losses[:4])  # same four views again: the loss went down


def test_second_gradient_tensor_and_per_parameter_steps():
    """extra_grads (the second all-reduce bucket of a data-parallel step) is summed inside the kernel and cleared by
    zero_grad; a parameter that misses a step keeps torch's per-parameter bias correction."""
    from gaustudio_b200.optimizers import FusedAdam
    shapes = [(1001, 3), (257,)]
    ours = [torch.nn.Parameter(p.clone()) for p in _params(3, shapes)]
    ref = [torch.nn.Parameter(p.clone()) for p in _params(3, shapes)]
    fo = FusedAdam(ours, lr=1e-2, eps=1e-12)
    to = torch.optim.AdamW(ref, lr=1e-2, eps=1e-12)
    for step in range(5):
        ga, gb = _params(200 + step, shapes), _params(300 + step, shapes)
        skip_second = step in (1, 3)  # the second parameter receives no gradient on these steps
        extra = []
        for i, (p, q) in enumerate(zip(ours, ref)):
            if i == 1 and skip_second:
                p.grad = None; q.grad = None; extra.append(None)
                continue
            p.grad = ga[i].clone(); extra.append(gb[i].clone())
            q.grad = (ga[i] + gb[i]) * 0.5
        fo.step(grad_scale=0.5, zero_grad=True, extra_grads=extra)
        to.step()
        for p, q, e in zip(ours, ref, extra):
            assert torch.allclose(p, q, rtol=0, atol=3e-6 * max(1.0, float(q.abs().max())))
            if e is not None:
                assert float(e.abs().max()) == 0 and float(p.grad.abs().max()) == 0
    assert fo.state[ours[0]]["step"] == 5 and fo.state[ours[1]]["step"] == 3

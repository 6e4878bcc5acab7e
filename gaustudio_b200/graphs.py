"""CUDA-graph capture of a whole per-view step (render -> loss -> backward [-> post]).

Every launch of the library has a grid that depends only on the number of Gaussians and the image size -- the
per-view instance count R stays on the device -- so in fixed-capacity mode the complete forward + backward is
capturable: one `cudaGraphLaunch` per view instead of ~45 kernel launches and ~35 framework ops, which makes the
throughput independent of host speed.  (The reference reads R back to the host inside every forward,
rasterizer_impl.cu:284, and launches on the legacy default stream: it cannot be captured.)

All replayed cameras must share image size and field of view (they are kernel parameters baked into the graph);
the camera matrices live in static device tensors that `__call__` refreshes before each replay.
"""
import torch

from . import _C


class _StaticCamera:
    def __init__(self, cam, dev):
        self.FoVx, self.FoVy = cam.FoVx, cam.FoVy
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.world_view_transform = cam.world_view_transform.to(dev).clone()
        self.full_proj_transform = cam.full_proj_transform.to(dev).clone()
        self.camera_center = cam.camera_center.to(dev).clone()

    def load(self, cam):
        self.world_view_transform.copy_(cam.world_view_transform, non_blocking=True)
        self.full_proj_transform.copy_(cam.full_proj_transform, non_blocking=True)
        self.camera_center.copy_(cam.camera_center, non_blocking=True)


class GraphedViewStep:
    """step = GraphedViewStep(renderer, model, loss_fn, example_cameras, capacity=None, post_fn=None)
    loss = step(camera)      # gradients of the model parameters are in their (static) .grad tensors

    `step.out` is the render's result dict (static tensors, refreshed by every replay).

    `loss_fn=None` captures a forward-only step (rendered under no_grad; `step(camera)` returns the result dict of
    static output tensors, `post_fn(camera, out)`'s value is in `step.extra`) -- the extraction-pass shape of cfg 5.

    `capacity`: binning capacity (tile instances) baked into the graph; default = 1.3 x the largest count seen on
    `example_cameras` (rendered eagerly once each).  `step.max_rendered()` returns the largest count any replay
    needed -- compare it with `step.capacity` (an overflowing view is rendered incompletely, never out of bounds).

    `accumulate=True` keeps the parameters' CURRENT `.grad` tensors (e.g. the views of a `parallel.GradBucket`) and
    captures the in-place accumulation into them, so several replays sum their gradients into the same buffers -- the
    caller zeroes them (the fused optimizer step does, `zero_grad=True`).  They are zeroed once after the capture.

    `tile_order` (baked into the captured launches): graphs exist to keep several views in flight on different streams,
    where raster order (0) packs better than the library's longest-first default (`_C.set_tile_order`); restored after.

    The fixed-capacity forward mode is only active during warm-up and capture: the caller's own forward mode
    (`_C.set_pipelined`) is restored before the constructor returns, so eager renders afterwards behave as before.

    Create it BEFORE running an eager backward on the same parameter tensors: autograd binds a leaf's gradient
    accumulator to the stream of its first backward, and a legacy-default-stream binding is illegal under capture
    (cudaErrorStreamCaptureImplicit).  Eager steps after the capture are fine."""

    def __init__(self, renderer, model, loss_fn, example_cameras, capacity=None, post_fn=None, device=None,
                 accumulate=False, tile_order=0):
        dev = device or model._xyz.device
        self.dev, self.model = dev, model
        cams = list(example_cameras)
        saved = _C.pipeline_state()
        saved_order = _C.set_tile_order(tile_order)
        try:
            if capacity is None:  # largest instance count over the example views (eager, exact mode) + 30 %
                _C.set_pipelined(False)
                worst = max(self._count(renderer, cam, model, dev) for cam in cams)
                capacity = _C._quantise(worst, 1.3)
            self.capacity = int(capacity)
            self.cam = _StaticCamera(cams[0], dev)
            self.train = loss_fn is not None
            params = model.parameters_list() if self.train else []
            _C.set_pipelined(True, fixed_capacity=self.capacity)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))

            def body():
                if not self.train:
                    with torch.no_grad():
                        out = renderer.render(self.cam, model)
                        extra = post_fn(self.cam, out) if post_fn is not None else None
                    return out, extra
                out = renderer.render(self.cam, model)
                self.out = out  # static output tensors of the captured step (refreshed by every replay)
                loss = loss_fn(out)
                loss.backward()
                extra = post_fn(self.cam, out) if post_fn is not None else None
                return loss.detach(), extra

            if accumulate and any(p.grad is None for p in params):
                raise ValueError("GraphedViewStep(accumulate=True): every parameter needs a .grad tensor to accumulate into")
            with torch.cuda.stream(side):
                for _ in range(3):  # warm-up outside capture (allocator, lazy module loads)
                    if not accumulate:
                        for p in params:
                            p.grad = None
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            if not accumulate:
                for p in params:
                    p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.extra = body()
            self.grads = [p.grad for p in params]
            if accumulate:
                for g in self.grads:
                    g.zero_()
            self.rmax = _C._pl().rmax  # device-side running max of num_rendered, updated inside the graph
            self.rmax.zero_()
        finally:
            _C.restore_pipeline(saved)
            _C.set_tile_order(saved_order)

    @staticmethod
    def _count(renderer, cam, model, dev):
        """num_rendered of one eager (exact-mode) forward."""
        out = renderer.render(_StaticCamera(cam, dev), model)
        fn = out["render"].grad_fn
        if fn is None:
            raise RuntimeError("GraphedViewStep: the model's parameters must require grad to size the capacity "
                               "(or pass capacity= explicitly)")
        return _C.last_num_binned()  # what sizes the binning buffer (fn.num_rendered keeps the reference's meaning)

    def __call__(self, camera):
        self.cam.load(camera)
        self.graph.replay()
        return self.loss

    def max_rendered(self):
        return int(self.rmax.item())

"""Device time of the extraction post-pass kernels at 1920x1080 (CUDA events, 50 iterations after warm-up), next to
the CPU path the reference takes for the same step (D2H + cv2.dilate + cv2.bilateralFilter + H2D), if cv2 is there."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gaustudio_b200 import extract, _lib
import ctypes as C

dev = torch.device("cuda"); H, W = 1080, 1920
g = torch.Generator().manual_seed(0)
depth = (2 + torch.rand(H, W, generator=g) * 3).to(dev); opacity = torch.rand(H, W, generator=g).to(dev)
median = depth.clone(); mask = opacity > 0.1


def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n


out = {}
out["masked_bilateral_ms"] = timed(lambda: extract.masked_bilateral_filter(depth, mask))
f, fg = extract.masked_bilateral_filter(depth, mask)
rot = torch.eye(3, device=dev).contiguous(); cam_n = torch.empty(H, W, 3, device=dev); negw = torch.empty(H, W, 3, device=dev)
valid = torch.empty(H, W, dtype=torch.bool, device=dev); L = _lib.lib(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
out["extract_normals_ms"] = timed(lambda: L.gsr_extract_normals(p(f), p(fg), p(opacity), p(median), W, H, 1182.0, 1182.0, 960.0, 540.0, p(rot), 1e9, 0.5, p(cam_n), p(negw), p(valid), st))
P, n = 1_000_000, 1_000_000
xyz = torch.randn(P, 3, device=dev); ids = torch.randint(0, P, (n,), device=dev); nrm = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=1)
conf = torch.rand(n, device=dev); sums = torch.zeros(P, 3, device=dev); wts = torch.zeros(P, device=dev); touched = torch.zeros(P, dtype=torch.uint8, device=dev)
out["fusion_pass_1M_ms"] = timed(lambda: L.gsr_normal_fusion_pass(n, p(ids), p(nrm), p(conf), P, p(xyz), 1.0, 2.0, 3.0, None, 0.8, p(sums), p(wts), p(touched), st))
# algorithmic bytes: bilateral = depth 4 + mask 1 read, mask 1 + depth 4 written, + (mask 1 + depth 4) re-read by the filter
px = H * W
out["masked_bilateral_GBps"] = px * 15 / out["masked_bilateral_ms"] / 1e6
out["extract_normals_GBps"] = px * (4 + 1 + 4 + 4 + 12 + 12 + 1) / out["extract_normals_ms"] / 1e6
try:
    import cv2
    def cpu():
        d = depth.cpu().numpy(); m = mask.cpu().numpy()
        inv = (1 - m).astype(np.uint8); nm = 1 - cv2.dilate(inv, np.ones((3, 3), np.uint8))
        v = nm == 1; lo, hi = d[v].min(), d[v].max(); nd = (d - lo) / (hi - lo); nd[~v] = 0
        fl = cv2.bilateralFilter(nd.astype(np.float32), d=3, sigmaColor=75, sigmaSpace=75) * (hi - lo) + lo
        fl[~v] = d[~v]; return torch.from_numpy(fl).to(dev), torch.from_numpy(nm).to(dev)
    cpu(); t = time.perf_counter()
    for _ in range(5): cpu()
    torch.cuda.synchronize(); out["cpu_opencv_path_ms"] = (time.perf_counter() - t) / 5 * 1e3
    out["cv2_threads"] = cv2.getNumThreads()
except Exception as e:
    out["cpu_opencv_path_ms"] = None; out["cpu_note"] = repr(e)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))

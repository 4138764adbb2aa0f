"""x2 bilinear / trilinear upsampling forward + backward per launch (developer tool): python tools/upsample_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chap_b200 import ops

CASES = [("2D 16ch 12x128x128", (12, 16, 128, 128)), ("2D 64ch 12x32x32", (12, 64, 32, 32)), ("3D 16ch 2x56x56x40", (2, 16, 56, 56, 40)),
         ("3D 32ch 2x28x28x20", (2, 32, 28, 28, 20))]
for name, shp in CASES:
    fmt = torch.channels_last if len(shp) == 4 else torch.channels_last_3d
    x = torch.randn(*shp, device="cuda").contiguous(memory_format=fmt).requires_grad_(True)
    y = ops.upsample2x(x)
    gy = torch.randn_like(y)
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps
    tf = timeit(lambda: ops.upsample2x(x.detach()))
    tb = timeit(lambda: torch.autograd.grad(ops.upsample2x(x), x, gy)) - tf
    mb = (x.numel() + y.numel()) * 4 / 1e6
    print(f"{name:24s} {mb:7.1f} MB  fwd {tf:6.1f} us ({mb / tf:5.2f} TB/s)  bwd {tb:6.1f} us ({mb / tb:5.2f} TB/s)")
# raw kernel time through the C ABI (no autograd / Python dispatch in the loop body beyond one ctypes call): CUDA graph of 20 launches
import ctypes
from chap_b200.ops import lib, _p, _stream
for name, shp in CASES:
    nd = len(shp) - 2
    fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
    x = torch.randn(*shp, device="cuda").contiguous(memory_format=fmt)
    n, c = shp[0], shp[1]
    d, h, w = (1, shp[2], shp[3]) if nd == 2 else shp[2:]
    yshape = (n, c) + tuple(2 * s for s in shp[2:])
    y = torch.empty(yshape, device="cuda").contiguous(memory_format=fmt)
    dx = torch.empty_like(x)
    def run(fn):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 40
    tf = run(lambda: lib().chap_upsample2x_fwd(_p(x), nd, n, d, h, w, c, _p(y), _stream()))
    tb = run(lambda: lib().chap_upsample2x_bwd(_p(y), nd, n, d, h, w, c, _p(dx), _stream()))
    mb = (x.numel() + y.numel()) * 4 / 1e6
    print(f"graph {name:24s} {mb:7.1f} MB  fwd {tf:6.1f} us ({mb / tf:5.2f} TB/s)  bwd {tb:6.1f} us ({mb / tb:5.2f} TB/s)")

"""Perturbation generator at the benchmark shapes (five 2D levels of 12 samples, or three 3D levels of 2): one chap_perturb_fwd call,
timed with CUDA events; also the unit for `ncu -k regex:"perturb_rows|chan_sq"` captures.   python tools/perturb_bench.py [3d]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from chap_b200 import ops  # noqa: E402

dev = "cuda:0"
if len(sys.argv) > 1 and sys.argv[1] == "3d":
    shapes = [(2, 16, 112, 112, 80), (2, 32, 56, 56, 40), (2, 64, 28, 28, 20), (2, 128, 14, 14, 10), (2, 256, 7, 7, 5)]
else:
    shapes = [(12, 16, 256, 256), (12, 32, 128, 128), (12, 64, 64, 64), (12, 128, 32, 32), (12, 256, 16, 16)]
gs = [ops.cl(torch.randn(s, device=dev)) for s in shapes]
fs = [ops.cl(torch.randn(s, device=dev)) for s in shapes]
elems = sum(g.numel() for g in gs)
for _ in range(3):
    ops.perturb(gs, fs, 6.0, "channel_spatial", g_scale=10.0)
torch.cuda.synchronize()
reps = 1 if os.environ.get("NCU") else 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.perturb(gs, fs, 6.0, "channel_spatial", g_scale=10.0)
e1.record()
torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / reps
print("perturb all levels: %.1f us per call, %.0f MB algorithmic (12 B/elt) -> %.0f GB/s; touched (20 B/elt) %.0f GB/s"
      % (us, 12e-6 * elems, 12.0 * elems / us / 1e3, 20.0 * elems / us / 1e3))

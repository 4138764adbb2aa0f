"""Sliding-window inference throughput: VNet on synthetic LA-like volumes, patch 112x112x80, stride 18/18/4.
    python tools/sw_bench.py [W H D] [--cases K]
Prints cases/s through chap_b200.test_3D_util.test_single_case (volume on the host in, label map on the host out) and the
aggregation kernel's own time / bandwidth from the library timers."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from chap_b200 import _lib, networks  # noqa: E402
from chap_b200.test_3D_util import test_single_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shape", nargs="*", type=int, default=[192, 192, 88])
ap.add_argument("--cases", type=int, default=3)
ap.add_argument("--batch-windows", type=int, default=4)
args = ap.parse_args()
torch.manual_seed(1337)
net = networks.net_factory_3d("vnet", in_chns=1, class_num=2, mode="test", device="cuda:0")
rng = np.random.RandomState(0)
vols = [rng.randn(*args.shape).astype(np.float32) for _ in range(args.cases)]
test_single_case(net, vols[0], 18, 4, (112, 112, 80), num_classes=2, batch_windows=args.batch_windows)      # warm-up
torch.cuda.synchronize()
_lib.timing_enable(True)
t0 = time.perf_counter()
for v in vols:
    lab = test_single_case(net, v, 18, 4, (112, 112, 80), num_classes=2, batch_windows=args.batch_windows)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.cases
rep = _lib.timing_report()
_lib.timing_enable(False)
agg = rep.get("sw_aggregate")
print("volume %s: %.3f s per case = %.2f cases/s (batch of %d windows per forward)" % ("x".join(map(str, args.shape)), dt, 1.0 / dt, args.batch_windows))
if agg:
    ms = agg["ms"] / agg["launches"]
    print("sw_aggregate: %.3f ms per case, %.0f GB/s algorithmic (%.0f MB)" % (ms, agg["bytes"] / agg["launches"] / ms / 1e6, agg["bytes"] / agg["launches"] / 1e6))
tot = sum(v["ms"] for v in rep.values()) / args.cases
print("kernel time per case by family (ms):", {k.split(":")[0]: round(v["ms"] / args.cases, 2) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:8]}, "sum %.1f" % tot)

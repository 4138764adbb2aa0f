"""Kernel launch mix of ONE CUDA-graph-replayed CHAP iteration (developer tool): torch.profiler (CUPTI) around a single replay.
    python tools/step_kernels.py [unet2d|vnet3d] [--eager]
Prints kernel name, launches, total microseconds; numbers are for finding launch overhead, not a bench value."""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from chap_b200.train_step import ChapTrainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "unet2d"
w = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev)
tr = ChapTrainer(model, n_classes=w["classes"], labeled_bs=w["labeled"], max_iterations=30000, use_graph="--eager" not in sys.argv, graph_warmup=2)
data = [tuple(t.to(dev) for t in bench.synth_batch(w, i)) for i in range(2)]
for i in range(5):
    tr.step(*data[i % 2])
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    tr.step(*data[0])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*", "", e.name)[:100]
        agg[n][0] += 1
        agg[n][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot_n = sum(v[0] for v in agg.values()); tot_t = sum(v[1] for v in agg.values())
print("%d device activities, %.1f us summed" % (tot_n, tot_t))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-102s %5d %9.1f" % (n, c, t))

import sys, os, re, collections
sys.path.insert(0, os.getcwd())
import torch, bench
from chap_b200.test_3D_util import test_single_case
from chap_b200.networks.vnet import VNet
import numpy as np
dev = torch.device("cuda:0")
net = VNet(n_channels=1, n_classes=2, normalization='batchnorm', has_dropout=False).to(dev).eval()
img = np.random.rand(192, 192, 88).astype(np.float32)
for _ in range(2): test_single_case(net, img, 18, 4, (112, 112, 80), num_classes=2)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    test_single_case(net, img, 18, 4, (112, 112, 80), num_classes=2)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*", "", e.name)[:90]
        agg[n][0] += 1; agg[n][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("total us", tot)
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-92s %5d %9.1f" % (n, c, t))

import sys, os
sys.path.insert(0, os.getcwd())
import torch
from chap_b200 import ops
dev = "cuda"
for (n, c, shp) in [(4, 16, (112, 112, 80)), (4, 32, (56, 56, 40))]:
    conv = torch.nn.Conv3d(c, c, 3, padding=1).to(dev)
    bn = torch.nn.BatchNorm3d(c).to(dev).eval()
    x = torch.randn(n, c, *shp, device=dev).contiguous(memory_format=torch.channels_last_3d)
    f = lambda: ops.conv_bn_act_eval(x, conv.weight, conv.bias, 0, bn, 0.0)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(c, shp, "%.1f us" % (e0.elapsed_time(e1) * 100))

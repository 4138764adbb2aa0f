import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import golden, rel_err, seeded_model
from chap_b200 import ops
for c in (0, 16, 32, 1<<20):
    ops.set_conv_precision(c)
    g = golden("unet2d.npz"); m = seeded_model("dualdecoder2d").to("cuda").train()
    x = torch.from_numpy(g["x"]).cuda()
    with ops.bn_tracking(False):
        o1, o2 = m(x)
    print("2d", c, rel_err(o1, g["o1"]), rel_err(o2, g["o2"]))
    g = golden("vnet3d.npz"); m = seeded_model("dualdecoder3d").to("cuda").train()
    x = torch.from_numpy(g["x"]).cuda()
    with ops.bn_tracking(False):
        o1, o2 = m(x)
    print("3d", c, rel_err(o1, g["o1"]), rel_err(o2, g["o2"]))
ops.set_conv_precision(0)

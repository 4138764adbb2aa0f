import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import replay
from replay import d64, rel, _torch_conv
from conftest import seeded_model
from chap_b200 import ops
DEV = "cuda:0"
ops.set_conv_precision(ops.PRECISE_ALL)
for kind in ("2d", "3d"):
    m = seeded_model("dualdecoder3d" if kind == "3d" else "dualdecoder2d", seed=17).to(DEV).train()
    x = (torch.randn(2, 1, 16, 16, 16) if kind == "3d" else torch.rand(2, 1, 48, 48)).to(DEV)
    def run():
        o1, o2 = m(x)
        (torch.softmax(o1, 1)[:, 0].mean() + (torch.softmax(o2, 1)[:, 1] ** 2).mean()).backward()
    rec = replay.record(run)
    for i, e in enumerate(rec):
        if e["name"] != "conv_stats" or e["gout"] is None: continue
        a, kw, gout = e["a"], e["kw"], e["gout"]
        xx, w, b, k = a[0], a[1], a[2], a[3]
        cat = kw.get("cat")
        if not xx.dtype.is_floating_point: continue
        nd = xx.dim() - 2
        xg = xx.detach().clone().requires_grad_(True)
        xr, wr, br = d64(xx).requires_grad_(True), d64(w), d64(b)
        if cat is None:
            y = ops.conv_stats(xg, w.detach(), b.detach(), k, False)[0]
            (gx,) = torch.autograd.grad(y, (xg,), gout)
            yr = _torch_conv(k, nd, xr, wr, br)
            (gr,) = torch.autograd.grad(yr, (xr,), d64(gout))
            print(kind, i, "kind", k, tuple(xx.shape), "->", tuple(y.shape), "fwd %.2e dx %.2e" % (rel(y, yr), rel(gx, gr)))
        else:
            cg, cr = cat.detach().clone().requires_grad_(True), d64(cat).requires_grad_(True)
            y = ops.conv_stats(xg, w.detach(), b.detach(), k, False, cat=cg)[0]
            gs = torch.autograd.grad(y, (xg, cg), gout)
            yr = _torch_conv(k, nd, torch.cat([xr, cr], 1), wr, br)
            grr = torch.autograd.grad(yr, (xr, cr), d64(gout))
            print(kind, i, "kind", k, "cat", tuple(xx.shape), "->", tuple(y.shape), "fwd %.2e dx %.2e %.2e" % (rel(y, yr), rel(gs[0], grr[0]), rel(gs[1], grr[1])))

"""Op-by-op replay checker (test infrastructure).  Runs a network forward + backward on the GPU while
recording every chap_b200.ops call with its inputs and the gradient that reached its output; then each op
is re-executed in isolation with the CUDA kernels and with a float64 torch restatement on the SAME inputs
and upstream gradient.  Unlike an end-to-end gradient comparison this is immune to the derivative
discontinuities of (Leaky)ReLU / max-pool (a single pre-activation whose sign differs between two fp32
evaluations changes a whole channel's BatchNorm gradient by ~1/sqrt(N) -- measured 4e-2 in one channel,
see DESIGN.md "conditioning"), so it can use a tight tolerance."""
import torch
import torch.nn.functional as F

from chap_b200 import _lib, ops

WRAPPED = ("conv_stats", "bn_act", "maxpool2", "upsample2x", "concat_channels")


def d64(t):
    return t.detach().double().cpu().contiguous()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def record(run):
    rec, orig = [], {k: getattr(ops, k) for k in WRAPPED}

    def wrap(name):
        fn = orig[name]

        def w(*a, **kw):
            out = fn(*a, **kw)
            t = out[0] if isinstance(out, tuple) else out
            e = dict(name=name, a=a, kw=kw, gout=None)
            if t.requires_grad:
                t.register_hook(lambda g, e=e: e.__setitem__("gout", g.detach().clone()))
            rec.append(e)
            return out
        return w
    for k in WRAPPED:
        setattr(ops, k, wrap(k))
    try:
        run()
    finally:
        for k in WRAPPED:
            setattr(ops, k, orig[k])
    return rec


def _torch_conv(kind, nd, x, w, b):
    conv = F.conv2d if nd == 2 else F.conv3d
    convt = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
    if kind == _lib.CONV_K3:
        return conv(x, w, b, padding=1)
    if kind == _lib.CONV_K1:
        return conv(x, w, b)
    if kind == _lib.CONV_DOWN2:
        return conv(x, w, b, stride=2)
    return convt(x, w, b, stride=2)


def check(rec):
    """Returns a list of (index, description, forward error, worst backward error)."""
    rows = []
    for i, e in enumerate(rec):
        if e["gout"] is None:
            continue
        name, a, kw, gout = e["name"], e["a"], e["kw"], e["gout"]
        if name == "conv_stats":
            x, w, b, kind = a[0], a[1], a[2], a[3]
            cat = kw.get("cat")                                            # conv on the channel concat (x, cat)
            nd = x.dim() - 2
            xg = x.detach().clone().requires_grad_(x.dtype.is_floating_point)
            wg, bg = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
            xr, wr, br = d64(x).requires_grad_(True), d64(w).requires_grad_(True), d64(b).requires_grad_(True)
            if cat is None:
                y = ops.conv_stats(xg, wg, bg, kind, False)[0]
                gs = torch.autograd.grad(y, (xg, wg, bg), gout)
                yr = _torch_conv(kind, nd, xr, wr, br)
                gr = torch.autograd.grad(yr, (xr, wr, br), d64(gout))
                berr = max(rel(u, v) for u, v in zip(gs[:2], gr[:2]))      # bias grads can be analytically ~0
            else:
                cg, cr = cat.detach().clone().requires_grad_(True), d64(cat).requires_grad_(True)
                y = ops.conv_stats(xg, wg, bg, kind, False, cat=cg)[0]
                gs = torch.autograd.grad(y, (xg, cg, wg), gout)
                yr = _torch_conv(kind, nd, torch.cat([xr, cr], 1), wr, br)
                gr = torch.autograd.grad(yr, (xr, cr, wr), d64(gout))
                berr = max(rel(u, v) for u, v in zip(gs, gr))
            rows.append((i, "conv kind %d %s->%s" % (kind, tuple(x.shape), tuple(y.shape)), rel(y, yr), berr))
        elif name == "bn_act":
            y, bn, slope = a[0], a[1], a[2]
            res, dnc, dele = kw.get("residual"), kw.get("drop_nc"), kw.get("drop_el")
            yg = y.detach().clone().requires_grad_(True)
            resg = None if res is None else res.detach().clone().requires_grad_(True)
            train = bn.training
            with ops.bn_tracking(False):
                out = ops.bn_act(yg, bn, slope, residual=resg, drop_nc=dnc, drop_el=dele)
            ins = [yg] + ([resg] if res is not None else [])
            gs = torch.autograd.grad(out, ins, gout)
            yr = d64(y).requires_grad_(True)
            if train:
                z = F.batch_norm(yr, None, None, d64(bn.weight), d64(bn.bias), True, 0.1, bn.eps)
            else:
                z = F.batch_norm(yr, d64(bn.running_mean), d64(bn.running_var), d64(bn.weight), d64(bn.bias), False, 0.1, bn.eps)
            outr = F.leaky_relu(z, slope)
            if dnc is not None:
                outr = outr * d64(dnc).reshape(tuple(dnc.shape) + (1,) * (y.dim() - 2))
            if dele is not None:
                outr = outr * d64(dele)
            insr = [yr]
            if res is not None:
                rr = d64(res).requires_grad_(True)
                outr = outr + rr
                insr.append(rr)
            gr = torch.autograd.grad(outr, insr, d64(gout))
            rows.append((i, "bn_act %s" % (tuple(y.shape),), rel(out, outr), max(rel(u, v) for u, v in zip(gs, gr))))
        else:
            ins = [t.detach().clone().requires_grad_(True) for t in a]
            out = getattr(ops, name)(*ins)
            gs = torch.autograd.grad(out, ins, gout)
            insr = [d64(t).requires_grad_(True) for t in a]
            nd = a[0].dim() - 2
            fn = {"maxpool2": lambda t: F.max_pool2d(t, 2),
                  "upsample2x": lambda t: F.interpolate(t, scale_factor=2, mode="bilinear" if nd == 2 else "trilinear", align_corners=True),
                  "concat_channels": lambda s, t: torch.cat([s, t], 1)}[name]
            outr = fn(*insr)
            gr = torch.autograd.grad(outr, insr, d64(gout))
            rows.append((i, "%s %s" % (name, tuple(a[0].shape)), rel(out, outr), max(rel(u, v) for u, v in zip(gs, gr))))
    return rows

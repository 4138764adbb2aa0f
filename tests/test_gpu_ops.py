"""GPU: every kernel of libchap_b200 against a plain PyTorch fp32 restatement of the same op
(run on the CPU in float64/float32) -- through the public ops / the C ABI."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, max_err, rel_err
from oracle import chap_losses as L

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# tolerance on convolution results: the tensor-core path multiplies in TF32 (10-bit mantissa, the
# reference's default cudnn.allow_tf32 behaviour), accumulates in fp32; CUDA-core path is fp32.
CONV_TOL = 3e-3
SIMT_TOL = 2e-5


def _ops():
    from chap_b200 import ops
    return ops


def _to_cl(x):
    return _ops().cl(x.to(DEV))


CONV_CASES = [
    # kind, nd, n, spatial, cin, cout
    ("k3", 2, 2, (12, 16), 1, 16), ("k3", 2, 3, (9, 7), 16, 16), ("k3", 2, 2, (16, 16), 32, 64),
    ("k3", 2, 1, (8, 8), 16, 4), ("k3", 2, 2, (5, 6), 256, 128),
    ("k1", 2, 2, (8, 8), 64, 32), ("k1", 3, 2, (4, 6, 5), 16, 2),
    ("up2", 2, 2, (6, 5), 32, 16), ("up2", 3, 2, (3, 4, 5), 32, 16),
    ("down2", 3, 2, (4, 6, 8), 16, 32), ("down2", 2, 2, (6, 8), 16, 32),
    ("k3", 3, 2, (6, 5, 7), 1, 16), ("k3", 3, 1, (5, 8, 6), 16, 16), ("k3", 3, 2, (4, 4, 5), 64, 32),
]


def _torch_conv(kind, nd, x, w, b):
    conv = F.conv2d if nd == 2 else F.conv3d
    convt = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
    if kind == "k3":
        return conv(x, w, b, padding=1)
    if kind == "k1":
        return conv(x, w, b)
    if kind == "down2":
        return conv(x, w, b, stride=2)
    return convt(x, w, b, stride=2)


@pytest.mark.parametrize("force_simt", [True, False])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "%s-%dd-n%d-%s-%d-%d" % (c[0], c[1], c[2], "x".join(map(str, c[3])), c[4], c[5]))
def test_conv_fwd_dgrad_wgrad(case, force_simt):
    from chap_b200 import _lib
    ops = _ops()
    kind, nd, n, sp, cin, cout = case
    kcode = {"k3": _lib.CONV_K3, "k1": _lib.CONV_K1, "down2": _lib.CONV_DOWN2, "up2": _lib.CONV_UP2}[kind]
    k = {"k3": 3, "k1": 1, "down2": 2, "up2": 2}[kind]
    g = torch.Generator().manual_seed(CONV_CASES.index(case))
    x = torch.randn((n, cin) + sp, generator=g, dtype=torch.float64)
    wshape = ((cin, cout) if kind == "up2" else (cout, cin)) + (k,) * nd
    w = torch.randn(wshape, generator=g, dtype=torch.float64) / (cin * k ** nd) ** 0.5
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y_ref = _torch_conv(kind, nd, x, w, b)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (x, w, b), gy)
    ops.set_force_simt(force_simt)
    try:
        xg = _to_cl(x.detach().float()).requires_grad_(True)
        wg = w.detach().float().to(DEV).requires_grad_(True)
        bg = b.detach().float().to(DEV).requires_grad_(True)
        y, sums = ops.conv_stats(xg, wg, bg, kcode, True)
        gx, gw, gb = torch.autograd.grad(y, (xg, wg, bg), gy.float().to(DEV))
        torch.cuda.synchronize()
    finally:
        ops.set_force_simt(False)
    tol = SIMT_TOL if force_simt else CONV_TOL
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) < tol, "fwd"
    assert rel_err(gx, gx_ref) < tol, "dgrad"
    assert rel_err(gw, gw_ref) < tol, "wgrad"
    assert rel_err(gb, gb_ref) < tol, "dbias"
    yd = y.detach().double().cpu()
    dims = (0,) + tuple(range(2, 2 + nd))
    sums = sums.reshape(_lib.STAT_SLOTS, 2 * cout).sum(0)          # partial statistics slots of the conv epilogue
    assert rel_err(sums[:cout], yd.sum(dims)) < 1e-4 + tol and rel_err(sums[cout:], (yd * yd).sum(dims)) < 1e-4 + tol, "fused BN statistics"


LARGE_CONV_CASES = [
    # many tiles per persistent CTA, ragged borders, several images; the big wgrad pixel blocks (16x16, 16x8)
    ("k3", 2, 5, (144, 136), 16, 16), ("k3", 2, 12, (128, 128), 32, 32), ("k3", 2, 8, (128, 128), 64, 32),
    ("k3", 2, 3, (200, 264), 16, 32), ("k3", 2, 6, (96, 96), 32, 16), ("k3", 3, 2, (24, 40, 48), 16, 16),
    ("k1", 2, 9, (64, 64), 64, 32), ("k3", 2, 2, (72, 72), 128, 128), ("k3", 2, 4, (64, 80), 16, 4), ("k3", 2, 12, (128, 128), 32, 64),
    ("up2", 2, 6, (64, 72), 32, 16), ("up2", 2, 3, (32, 32), 256, 128), ("up2", 3, 2, (12, 14, 10), 32, 16), ("up2", 2, 12, (128, 128), 32, 16),
    ("down2", 3, 2, (24, 40, 48), 16, 32), ("down2", 2, 6, (64, 72), 32, 64), ("down2", 3, 2, (8, 12, 10), 128, 256), ("up2", 3, 1, (14, 14, 10), 128, 64),
    # round 2: 3D kx-in-N convolution tile (one k chunk, N <= 32, W >= 32: tiles of 30 output columns, partial last tile; 32 -> 32 keeps 108 KB of
    # weights resident with one CTA per SM) and the taps-in-N weight gradient (x channel groups split over CTAs when 3 * cout * groups > 512 columns)
    ("k3", 3, 2, (10, 12, 36), 32, 32), ("k3", 3, 1, (8, 10, 33), 16, 32), ("k3", 3, 1, (6, 9, 61), 32, 16), ("k3", 3, 3, (5, 6, 32), 16, 16),
    ("k3", 2, 2, (40, 48), 128, 64), ("k3", 2, 3, (24, 40), 64, 64), ("k3", 2, 2, (32, 32), 256, 32), ("k3", 2, 5, (56, 40), 64, 32),
]


@pytest.mark.parametrize("case", LARGE_CONV_CASES, ids=lambda c: "%s-%dd-n%d-%s-%d-%d" % (c[0], c[1], c[2], "x".join(map(str, c[3])), c[4], c[5]))
def test_conv_large_shapes_tensor_core(case):
    """Persistent multi-tile path: reference is torch fp64 on the GPU (same op, no TF32)."""
    from chap_b200 import _lib
    ops = _ops()
    kind, nd, n, sp, cin, cout = case
    kcode = {"k3": _lib.CONV_K3, "k1": _lib.CONV_K1, "up2": _lib.CONV_UP2, "down2": _lib.CONV_DOWN2}[kind]
    k = {"k3": 3, "k1": 1, "up2": 2, "down2": 2}[kind]
    g = torch.Generator().manual_seed(100 + LARGE_CONV_CASES.index(case))
    x = torch.randn((n, cin) + sp, generator=g).to(DEV)
    wshape = ((cin, cout) if kind == "up2" else (cout, cin)) + (k,) * nd
    w = (torch.randn(wshape, generator=g) / (cin * k ** nd) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    y_ref = _torch_conv(kind, nd, xd, wd, bd)
    gy = torch.randn(y_ref.shape, generator=g).to(DEV)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (xd, wd, bd), gy.double())
    xg = ops.cl(x).requires_grad_(True)
    wg = w.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    y, sums = ops.conv_stats(xg, wg, bg, kcode, True)
    gx, gw, gb = torch.autograd.grad(y, (xg, wg, bg), ops.cl(gy))
    torch.cuda.synchronize()
    assert rel_err(y, y_ref) < CONV_TOL, "fwd"
    assert rel_err(gx, gx_ref) < CONV_TOL, "dgrad"
    assert rel_err(gw, gw_ref) < CONV_TOL, "wgrad"
    assert rel_err(gb, gb_ref) < CONV_TOL, "dbias"
    dims = (0,) + tuple(range(2, 2 + nd))
    sums = sums.reshape(_lib.STAT_SLOTS, 2 * cout).sum(0)
    yd = y.detach().double()
    assert rel_err(sums[:cout], yd.sum(dims)) < 1e-4 and rel_err(sums[cout:], (yd * yd).sum(dims)) < 1e-4, "fused BN statistics"


@pytest.mark.parametrize("case", [(2, 3, (40, 56), 16, 16, 16), (2, 12, (128, 128), 32, 32, 32), (2, 2, (16, 16), 128, 128, 128),
                                  (2, 2, (12, 12), 16, 8, 16)])
def test_conv_on_channel_concat_fused_split(case):
    """conv(cat(a, b)): forward and the two data gradients written by the split epilogue (or the fallback split pass)."""
    from chap_b200 import _lib
    ops = _ops()
    nd, n, sp, ca, cb, cout = case
    g = torch.Generator().manual_seed(7)
    a = torch.randn((n, ca) + sp, generator=g).to(DEV)
    b = torch.randn((n, cb) + sp, generator=g).to(DEV)
    w = (torch.randn((cout, ca + cb, 3, 3), generator=g) / (9 * (ca + cb)) ** 0.5).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    ad, bd, wd, biasd = (t.double().requires_grad_(True) for t in (a, b, w, bias))
    y_ref = F.conv2d(torch.cat([ad, bd], 1), wd, biasd, padding=1)
    gy = torch.randn(y_ref.shape, generator=g).to(DEV)
    refs = torch.autograd.grad(y_ref, (ad, bd, wd), gy.double())
    ag, bg, wg = ops.cl(a).requires_grad_(True), ops.cl(b).requires_grad_(True), w.clone().requires_grad_(True)
    y, _ = ops.conv_stats(ag, wg, bias, _lib.CONV_K3, True, cat=bg)
    outs = torch.autograd.grad(y, (ag, bg, wg), ops.cl(gy))
    torch.cuda.synchronize()
    assert rel_err(y, y_ref) < CONV_TOL
    for name, o, r in zip(("d/da", "d/db", "d/dw"), outs, refs):
        assert o.shape == r.shape and rel_err(o, r) < CONV_TOL, name


@pytest.mark.parametrize("nd,train,with_res,drop", list(itertools.product([2, 3], [True, False], [False, True], ["none", "nc", "el"])))
def test_bn_act_fwd_bwd(nd, train, with_res, drop):
    ops = _ops()
    torch.manual_seed(nd * 7 + int(train))
    n, c = 3, 16
    sp = (6, 5) if nd == 2 else (3, 4, 5)
    y = torch.randn((n, c) + sp, dtype=torch.float64) * 2 + 0.5
    bn = (torch.nn.BatchNorm2d if nd == 2 else torch.nn.BatchNorm3d)(c).double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
        bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
    bn.train(train)
    res = torch.randn_like(y) if with_res else None
    dnc = (torch.rand(n, c, dtype=torch.float64) > 0.5).double() * 2 if drop == "nc" else None
    dele = (torch.rand_like(y) > 0.3).double() / 0.7 if drop == "el" else None
    slope = 0.01 if nd == 2 else 0.0
    bn_g = (torch.nn.BatchNorm2d if nd == 2 else torch.nn.BatchNorm3d)(c).to(DEV)
    bn_g.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    bn_g.train(train)
    yr = y.clone().requires_grad_(True)
    resr = res.clone().requires_grad_(True) if with_res else None
    out_ref = F.leaky_relu(bn(yr), slope)
    if dnc is not None:
        out_ref = out_ref * dnc.reshape((n, c) + (1,) * nd)
    if dele is not None:
        out_ref = out_ref * dele
    if with_res:
        out_ref = out_ref + resr
    gout = torch.randn_like(out_ref)
    ins = [yr, bn.weight, bn.bias] + ([resr] if with_res else [])
    grads_ref = torch.autograd.grad(out_ref, ins, gout)
    yg = _to_cl(y.float()).requires_grad_(True)
    resg = _to_cl(res.float()).requires_grad_(True) if with_res else None
    out = ops.bn_act(yg, bn_g, slope, residual=resg, drop_nc=None if dnc is None else dnc.float().to(DEV),
                     drop_el=None if dele is None else _to_cl(dele.float()))
    ins_g = [yg, bn_g.weight, bn_g.bias] + ([resg] if with_res else [])
    grads = torch.autograd.grad(out, ins_g, gout.float().to(DEV))
    assert rel_err(out, out_ref) < 1e-5
    for a, b_, nm in zip(grads, grads_ref, ["dy", "dgamma", "dbeta", "dres"]):
        assert rel_err(a, b_) < 2e-5, nm
    if train:
        assert rel_err(bn_g.running_mean, bn.running_mean) < 1e-6 and rel_err(bn_g.running_var, bn.running_var) < 1e-6
        assert int(bn_g.num_batches_tracked) == int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("p,c", [(0.05, 16), (0.5, 256), (0.3, 64)])
def test_generated_dropout_equals_its_own_mask_and_has_the_right_rate(p, c):
    """nn.Dropout(p) generated inside the BatchNorm / activation kernels (chap_bn_act_{fwd,bwd}_rng, unet.py:53): the mask the
    forward applied is recovered with an identity BatchNorm on ones; the generated backward must equal the explicit-mask backward
    on that mask; keep rate = 1 - p within 5 sigma; subsequence and the device-side epoch word change the draw, equal keys repeat it."""
    ops = _ops()
    torch.manual_seed(3)
    n, h, w = 4, 24, 20
    bn = torch.nn.BatchNorm2d(c).to(DEV).eval()                       # identity: scale 1 / sqrt(1 + eps), shift 0
    ones = _to_cl(torch.ones(n, c, h, w, device=DEV))
    epoch = torch.zeros(1, dtype=torch.int64, device=DEV)
    rng = (p, 1234, 7, epoch)
    inv = float(torch.sqrt(torch.tensor(1.0 + bn.eps)))
    mask = ops.bn_act(ones, bn, 1.0, drop_rng=rng) * inv
    vals = torch.unique(mask)
    keep = 1.0 - p
    assert vals.numel() == 2 and abs(float(vals[0])) == 0.0 and abs(float(vals[1]) - 1.0 / keep) < 1e-5
    rate = float((mask > 0).double().mean())
    assert abs(rate - keep) < 5.0 * (keep * p / mask.numel()) ** 0.5, (rate, keep)
    for ch in (0, c - 1):                                              # no channel / position structure
        assert abs(float((mask[:, ch] > 0).double().mean()) - keep) < 6.0 * (keep * p / mask[:, ch].numel()) ** 0.5
    assert torch.equal(mask, ops.bn_act(ones, bn, 1.0, drop_rng=rng) * inv)
    assert not torch.equal(mask, ops.bn_act(ones, bn, 1.0, drop_rng=(p, 1234, 8, epoch)) * inv)
    epoch.fill_(5)
    assert not torch.equal(mask, ops.bn_act(ones, bn, 1.0, drop_rng=rng) * inv)
    epoch.fill_(0)
    # train-mode layer: generated forward / backward == explicit-mask forward / backward on the recovered mask
    bn_a = torch.nn.BatchNorm2d(c).to(DEV).train()
    bn_b = torch.nn.BatchNorm2d(c).to(DEV).train()
    y = _to_cl(torch.randn(n, c, h, w, device=DEV) * 2 + 0.3)
    gout = _to_cl(torch.randn(n, c, h, w, device=DEV))
    ya, yb = y.clone().requires_grad_(True), y.clone().requires_grad_(True)
    out_a = ops.bn_act(ya, bn_a, 0.01, drop_rng=rng)
    out_b = ops.bn_act(yb, bn_b, 0.01, drop_el=mask)
    assert torch.equal(out_a, out_b)
    ga = torch.autograd.grad(out_a, [ya, bn_a.weight, bn_a.bias], gout)
    gb = torch.autograd.grad(out_b, [yb, bn_b.weight, bn_b.bias], gout)
    for a_, b_, nm in zip(ga, gb, ["dy", "dgamma", "dbeta"]):
        assert rel_err(a_, b_) < 1e-6, nm


def test_bn_tracking_disabled_keeps_running_stats():
    ops = _ops()
    bn = torch.nn.BatchNorm2d(16).to(DEV).train()
    y = _to_cl(torch.randn(2, 16, 4, 4))
    with ops.bn_tracking(False):
        ops.bn_act(y, bn, 0.01)
    assert float(bn.running_mean.abs().sum()) == 0.0 and int(bn.num_batches_tracked) == 0


def test_maxpool_upsample_concat():
    ops = _ops()
    torch.manual_seed(0)
    x = torch.randn(2, 16, 8, 6, dtype=torch.float64, requires_grad=True)
    ref = F.max_pool2d(x, 2)
    g = torch.randn_like(ref)
    (gx_ref,) = torch.autograd.grad(ref, x, g)
    xg = _to_cl(x.detach().float()).requires_grad_(True)
    out = ops.maxpool2(xg)
    (gx,) = torch.autograd.grad(out, xg, g.float().to(DEV))
    assert max_err(out, ref) < 1e-6 and max_err(gx, gx_ref) < 1e-6
    # (the larger shapes have several compact pixel tiles per dimension with partial tiles at the borders: the vector kernels' tile walk)
    for nd, shape in ((2, (2, 8, 5, 7)), (3, (2, 4, 3, 5, 4)), (3, (1, 8, 1, 2, 3)), (2, (1, 4, 1, 1)), (2, (3, 16, 21, 19)), (3, (2, 16, 7, 9, 5)),
                      (3, (1, 64, 3, 5, 6)), (2, (2, 256, 6, 5)), (2, (1, 6, 4, 3))):
        x = torch.randn(shape, dtype=torch.float64, requires_grad=True)
        ref = F.interpolate(x, scale_factor=2, mode="bilinear" if nd == 2 else "trilinear", align_corners=True)
        g = torch.randn_like(ref)
        (gx_ref,) = torch.autograd.grad(ref, x, g)
        xg = _to_cl(x.detach().float()).requires_grad_(True)
        out = ops.upsample2x(xg)
        (gx,) = torch.autograd.grad(out, xg, g.float().to(DEV))
        assert rel_err(out, ref) < 1e-6, shape
        assert rel_err(gx, gx_ref) < 1e-6, shape
    a = torch.randn(2, 16, 4, 5, requires_grad=True)
    b = torch.randn(2, 32, 4, 5, requires_grad=True)
    ag, bg = _to_cl(a.detach()).requires_grad_(True), _to_cl(b.detach()).requires_grad_(True)
    out = ops.concat_channels(ag, bg)
    g = torch.randn(2, 48, 4, 5)
    ga, gb = torch.autograd.grad(out, (ag, bg), g.to(DEV))
    assert torch.equal(out.cpu(), torch.cat([a, b], 1)) and torch.equal(ga.cpu(), g[:, :16]) and torch.equal(gb.cpu(), g[:, 16:])
    s = torch.rand(2, 16)
    assert rel_err(ops.channel_scale(ag, s.to(DEV)), a.detach() * s[:, :, None, None]) < 1e-7
    m = (torch.rand(4, 5) > 0.5).long()
    c1, c2 = torch.randn(2, 1, 4, 5), torch.randn(2, 1, 4, 5)
    assert torch.equal(ops.mask_mix(c1.to(DEV), c2.to(DEV), m.to(DEV)).cpu(), c1 * m + c2 * (1 - m))


@pytest.mark.parametrize("c,nd", [(4, 2), (2, 3)])
def test_loss_kernels_match_frozen_oracle(c, nd):
    ops = _ops()
    from chap_b200.utils import losses
    torch.manual_seed(c)
    sp = (12, 16) if nd == 2 else (4, 8, 8)
    n = 3
    logits = (torch.randn((n, c) + sp) * 2).requires_grad_(True)
    logits2 = torch.randn((n, c) + sp) * 2
    lab_a, lab_b = torch.randint(0, c, (n,) + sp), torch.randint(0, c, (n,) + sp).float()
    mask = L.generate_mask(sp, tuple(1 for _ in sp))
    lmask = mask.unsqueeze(0).expand((n,) + sp)
    ref = L.mix_loss(logits, lab_a, lab_b, lmask, c, unlab=True)
    (gref,) = torch.autograd.grad(ref[2] + 0.3 * ref[0], logits)
    lg = _to_cl(logits.detach()).requires_grad_(True)
    out = losses.mix_loss(lg, lab_a.to(DEV), lab_b.to(DEV), lmask.to(DEV), unlab=True)
    (gg,) = torch.autograd.grad(out[2] + 0.3 * out[0], lg)
    for a, b in zip(out, ref):
        assert abs(float(a) - float(b)) < 1e-5 * max(1.0, abs(float(b)))
    assert rel_err(gg, gref) < 1e-4
    # pseudo-label block + patch mask
    s1, s2, a1, a2, know = ops.pseudo_label(_to_cl(logits.detach()), _to_cl(logits2))
    rs1, rs2, ra1, ra2, rknow = L.pseudo_label_block(logits.detach(), logits2)
    assert rel_err(s1, rs1) < 1e-6 and rel_err(s2, rs2) < 1e-6 and rel_err(know, rknow) < 1e-5
    assert torch.equal(a1.cpu(), ra1) and torch.equal(a2.cpu(), ra2)
    dm = ops.patch_topk_mask(know, a1, a2, 4, 0.25)
    rdm = L.create_mask_v1(ra1, ra2, rknow, 4, 0.25)
    assert float((dm.cpu() != rdm).float().mean()) < 0.02          # a patch can flip only through a float tie at the threshold
    # consistency distances, masked and unmasked, both types
    for losstype, m, red in itertools.product(("kl", "dice"), (None, rdm), ("mean", "batchmean")):
        r = L.consistency_distance(logits, rs2, m, losstype, red)
        (gr,) = torch.autograd.grad(r, logits)
        o = losses.consistency_distance(lg, s2, None if m is None else m.to(DEV), losstype, red)
        (go,) = torch.autograd.grad(o, lg)
        assert abs(float(o) - float(r)) < 1e-5 * abs(float(r)) + 1e-9, (losstype, m is None, red)      # north_star: <= 1e-3 on losses
        assert rel_err(go, gr) < 1e-4, (losstype, m is None, red)
    assert torch.equal(ops.argmax(lg).cpu(), logits.detach().argmax(1))
    assert rel_err(ops.softmax(lg), torch.softmax(logits.detach(), 1)) < 1e-6


def test_loss_kernels_match_frozen_fixture():
    from chap_b200.utils import losses
    g = golden("frozen_losses.npz")
    lg = _to_cl(torch.from_numpy(g["logits"]))
    out = losses.mix_loss(lg, torch.from_numpy(g["lab_a"]).to(DEV), torch.from_numpy(g["lab_b"]).to(DEV),
                          torch.from_numpy(g["mask"]).to(DEV), unlab=True)
    np.testing.assert_allclose([float(v) for v in out], g["mix"], rtol=1e-5)


@pytest.mark.parametrize("mode", ["channel_spatial", "channel", "spatial", "sample"])
def test_perturbation_generator(mode):
    """tolerance from BASELINE.json north_star: <= 1e-4 relative on perturbation tensors."""
    ops = _ops()
    torch.manual_seed(1)
    shapes = [(3, 16, 16, 16), (3, 32, 8, 8), (3, 64, 4, 4), (3, 128, 2, 2), (3, 256, 1, 1)]
    gs = [torch.randn(s) * (10.0 ** -i) for i, s in enumerate(shapes)]
    fs = [torch.randn(s) for s in shapes]
    outs = ops.perturb([_to_cl(t) for t in gs], [_to_cl(t) for t in fs], 6.0, mode, g_scale=10.0)
    for o, g_, f in zip(outs, gs, fs):
        r_ref = L.perturbation(g_ * 10.0, 6.0, mode)
        r = o.cpu() - f
        assert rel_err(r, r_ref) < 1e-4
        assert rel_err(o, f + r_ref) < 1e-6
    g3 = torch.randn(2, 16, 4, 6, 5)
    (o3,) = ops.perturb([_to_cl(g3)], None, 2.0, mode)
    assert rel_err(o3, L.perturbation(g3, 2.0, mode)) < 1e-4
    gf = golden("frozen_losses.npz")
    key = {"channel_spatial": "r_cs", "channel": "r_c", "spatial": "r_s", "sample": "r_n"}[mode]
    (og,) = ops.perturb([_to_cl(torch.from_numpy(gf["gfield"]))], None, 6.0, mode)
    assert rel_err(og, gf[key]) < 1e-4


def test_perturbation_generator_all_levels_one_call():
    """The five pyramid levels (different C, rows) of 3 samples through ONE chap_perturb_fwd call = 3 launches + a zero-fill
    (every phase runs for all levels through a descriptor table): each level <= 1e-4 of the oracle, with and without f."""
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    shapes = [(3, 16, 32, 32), (3, 32, 16, 16), (3, 64, 8, 8), (3, 128, 4, 4), (3, 256, 2, 2)]
    gs = [torch.randn(s, generator=g) * (10.0 ** (-i)) for i, s in enumerate(shapes)]          # very different magnitudes per level
    fs = [torch.randn(s, generator=g) for s in shapes]
    for mode in ("channel_spatial", "channel", "spatial", "sample"):
        outs = ops.perturb([_to_cl(t) for t in gs], [_to_cl(t) for t in fs], 6.0, mode, g_scale=10.0)
        for o, gg, ff in zip(outs, gs, fs):
            want = ff + L.perturbation(gg * 10.0, 6.0, mode)
            assert rel_err(o - ff.to(DEV), want - ff) < 1e-4, mode          # the perturbation itself, not f + r
    outs = ops.perturb([_to_cl(t) for t in gs], None, 6.0, "channel_spatial")
    for o, gg in zip(outs, gs):
        assert rel_err(o, L.perturbation(gg, 6.0, "channel_spatial")) < 1e-4
    # 3D levels
    g3 = [torch.randn(2, c, s, s, s, generator=g) for c, s in ((16, 8), (32, 4), (64, 2))]
    outs = ops.perturb([_to_cl(t) for t in g3], None, 6.0, "channel_spatial")
    for o, gg in zip(outs, g3):
        assert rel_err(o, L.perturbation(gg, 6.0, "channel_spatial")) < 1e-4


def test_l2n_axpy_and_sgd():
    ops = _ops()
    torch.manual_seed(2)
    d, base = torch.randn(3, 16, 5, 7) - 0.5, torch.randn(3, 16, 5, 7)
    out = ops.l2n_sample_axpy(_to_cl(d), _to_cl(base), 10.0)
    assert rel_err(out, base + 10.0 * L.l2n_sample(d)) < 1e-6
    # all levels of the VAT probe in two launches (ragged shapes, one level without a base)
    ds = [torch.randn(3, c, s, s + 1) - 0.5 for c, s in ((16, 12), (32, 6), (64, 3), (128, 2), (256, 1))]
    bs = [torch.randn_like(t) for t in ds]
    bs[3] = None
    outs = ops.l2n_sample_axpy_all([_to_cl(t) for t in ds], [None if t is None else _to_cl(t) for t in bs], 10.0)
    for o, t, bb in zip(outs, ds, bs):
        assert rel_err(o, (0 if bb is None else bb) + 10.0 * L.l2n_sample(t)) < 1e-6
    p, g = torch.randn(1003), torch.randn(1003)
    pr, buf = p.clone(), [None]
    pg, gg, bg = p.to(DEV), g.to(DEV), torch.zeros(1003, device=DEV)
    for it in range(3):
        L.sgd_momentum_step([pr], [g * (it + 1)], buf, 0.01 * (it + 1))
        ops.sgd_momentum_(pg, gg * (it + 1), bg, 0.01 * (it + 1), 0.9, 1e-4, 1.0, it == 0)
    assert rel_err(pg, pr) < 1e-6 and rel_err(bg, buf[0]) < 1e-6


def test_tensor_core_path_is_taken_and_accurate_to_tf32():
    """Supported layers must run on the tcgen05 kernel (not silently on CUDA cores) and agree with the fp32
    CUDA-core result to TF32 rounding (10-bit mantissa products, fp32 accumulation)."""
    from chap_b200 import _lib
    ops = _ops()
    torch.manual_seed(0)
    x = _to_cl(torch.randn(2, 64, 24, 40))
    w = (torch.randn(128, 64, 3, 3) / 24.0).to(DEV)
    b = torch.randn(128).to(DEV)
    _lib.timing_enable(True)
    y_tc, s_tc = ops.conv_stats(x, w, b, _lib.CONV_K3, True)
    fam = _lib.timing_report()
    _lib.timing_enable(False)
    launches = sum(v["launches"] for k, v in fam.items() if k.split(":")[0] == "conv_tc_fwd")   # (names carry the shape under CHAP_TIMING_DETAIL)
    assert launches == 1, fam
    ops.set_force_simt(True)
    try:
        y_ref, s_ref = ops.conv_stats(x, w.clone(), b, _lib.CONV_K3, True)
    finally:
        ops.set_force_simt(False)
    err = rel_err(y_tc, y_ref)
    assert 1e-6 < err < 1e-3, err                       # > 1e-6: it really went through TF32; < 1e-3: correct
    s_tc = s_tc.reshape(_lib.STAT_SLOTS, -1).sum(0)
    s_ref = s_ref.reshape(_lib.STAT_SLOTS, -1).sum(0)
    assert rel_err(s_tc, s_ref) < 1e-3


def test_feature_dropout_kernel_matches_reference_fixture():
    """perform_dropout on the fused CUDA kernel, given the factors the reference applied (tests/golden/filter_dropout.npz,
    generated from the unmodified reference): outputs bit-exact; backward equals autograd through the oracle restatement."""
    from chap_b200.networks import FilterDropout as fd
    from oracle import filter_dropout as ofd
    g = golden("filter_dropout.npz")
    feats = [torch.from_numpy(g["feat%d" % i]) for i in range(3)]
    for name in ("binomial_comp", "dropout2d", "scores", "scores_comp"):
        masks = [(torch.from_numpy(g["%s_m1_%d" % (name, i)]), torch.from_numpy(g["%s_m2_%d" % (name, i)])) if i in (0, 2) else None
                 for i in range(3)]
        fg = [_to_cl(f).requires_grad_(True) for f in feats]
        mg = [None if m is None else (m[0].to(DEV), m[1].to(DEV)) for m in masks]
        o1, o2 = fd.perform_dropout(fg, [0, 2], None, False, masks=mg)
        for i in range(3):
            assert o1[i].shape[0] == 6
            np.testing.assert_allclose(o1[i].detach().cpu().numpy(), g["%s_fp1_%d" % (name, i)], rtol=2e-7, atol=0)
            np.testing.assert_allclose(o2[i].detach().cpu().numpy(), g["%s_fp2_%d" % (name, i)], rtol=2e-7, atol=0)
        fr = [f.clone().requires_grad_(True) for f in feats]
        r1, r2 = ofd.perform_dropout(fr, masks)
        ws = [torch.randn_like(t) for t in r1 + r2]
        gr = torch.autograd.grad(sum((t * w).sum() for t, w in zip(r1 + r2, ws)), fr)
        gg = torch.autograd.grad(sum((t * w.to(DEV)).sum() for t, w in zip(o1 + o2, ws)), fg)
        for a, b in zip(gg, gr):
            assert rel_err(a, b) < 1e-6
    # drawing on the device (no explicit masks): shapes and the mean-preserving property of Dropout2d(0.5) factors
    o1, o2 = fd.perform_dropout([_to_cl(f) for f in feats], [0, 1, 2], None, False)
    assert all(t.shape[0] == 6 for t in o1 + o2)
    ratio = (o1[0][4:] / _to_cl(feats[0])[2:]).flatten()
    assert set(torch.unique(ratio[torch.isfinite(ratio)].round()).tolist()) <= {0.0, 2.0}

"""Host-side operators: torch.autograd.Function wrappers over the C ABI of libchap_b200.so.

PyTorch is plumbing here (device memory, streams, autograd tape); every device computation
is a kernel of this repository launched through ctypes with raw pointers.  Tensors are fp32
and channels-last (logical [N, C, *spatial], physical [N, *spatial, C]).  There is no CPU
path: every operator raises if its inputs are not CUDA tensors.
"""
import contextlib
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from ._lib import ConvDesc, check

_state = {"persistent_stats": True, "bn_epoch": 0, "epoch": 0, "bn_tracking": True, "weight_grad": True, "bias_grad_none": False, "grad_sink": False}


def lib():
    return _lib.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("chap_b200 operators need CUDA tensors (there is no CPU fallback)")


def mem_format(ndim):
    return torch.channels_last if ndim == 4 else torch.channels_last_3d


def cl(x):
    """fp32, channels-last contiguous view/copy of a [N, C, *spatial] tensor."""
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous(memory_format=mem_format(x.dim()))


def empty_cl(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device, memory_format=mem_format(len(shape)))


def _spatial3(x):
    """(nd, D, H, W) with D = 1 for 2D tensors."""
    if x.dim() == 4:
        return 2, 1, x.shape[2], x.shape[3]
    return 3, x.shape[2], x.shape[3], x.shape[4]


# ----------------------------------------------------------------------------- global switches
def invalidate_weight_cache():
    """Call after parameters were modified behind autograd's back (the fused SGD kernel)."""
    _state["epoch"] += 1


@contextlib.contextmanager
def bn_tracking(flag):
    """BatchNorm keeps using batch statistics in train mode but does not touch the running
    statistics while flag is False (the VAT passes)."""
    old, _state["bn_tracking"] = _state["bn_tracking"], flag
    try:
        yield
    finally:
        _state["bn_tracking"] = old


@contextlib.contextmanager
def no_weight_grad():
    """Convolutions record data gradients only (the decoder data-grad pass of VAT)."""
    old, _state["weight_grad"] = _state["weight_grad"], False
    try:
        yield
    finally:
        _state["weight_grad"] = old


@contextlib.contextmanager
def zero_bias_grad_as_none(flag=True):
    """Inside: the analytically-zero gradient of a bias that feeds a train-mode BatchNorm is returned as None (autograd's
    zero) instead of a zeros tensor -- saves a fill and an accumulate kernel per layer for optimisers that treat a missing
    .grad as zero (FlatSGD)."""
    old = _state["bias_grad_none"]
    _state["bias_grad_none"] = bool(flag)
    try:
        yield
    finally:
        _state["bias_grad_none"] = old


@contextlib.contextmanager
def grad_sink(flag=True):
    """Inside: parameters that carry a `_chap_sink` attribute (set by FlatSGD: a view of the flat gradient arena plus, for conv
    weights, a persistent zeroed scratch) get their gradients ADDED straight into the arena by the kernels
    (chap_conv_wgrad_acc / chap_bn_act_bwd_acc) and autograd receives None for them: no AccumulateGrad add kernels for
    parameters used by several passes, no zero-fills of fresh gradient tensors, no gather copy before the optimiser."""
    old = _state["grad_sink"]
    _state["grad_sink"] = bool(flag)
    try:
        yield
    finally:
        _state["grad_sink"] = old


def _persistent_zeros(holder, attr, n):
    """A float64 zero buffer of n elements that lives on `holder` (a BatchNorm module).  The kernels that use it hand it back
    zeroed, so it is allocated and cleared exactly once (saves one zero-fill launch per convolution / BatchNorm backward)."""
    buf = getattr(holder, attr, None)
    dev = holder.weight.device
    if buf is None or buf.numel() != n or buf.device != dev:
        buf = torch.zeros(n, dtype=torch.float64, device=dev)
        setattr(holder, attr, buf)
    return buf


def _sink_of(p):
    return getattr(p, "_chap_sink", None) if p is not None else None


def set_force_simt(flag):
    lib().chap_set_force_simt(1 if flag else 0)


def set_pdl(flag):
    """Programmatic dependent launch between the library's kernels (default off: measured without gain; results are identical either way)."""
    lib().chap_set_pdl(1 if flag else 0)
    invalidate_weight_cache()


def set_conv_precision(max_channels):
    """0: plain TF32 tensor-core convolutions; c > 0: split-operand 3xTF32 for layers with max(Cin, Cout) <= c
    (>= 1024: every layer, the "precise mode"); see chap_set_conv_precision in include/chap_b200.h."""
    lib().chap_set_conv_precision(int(max_channels))


PRECISE_ALL = 1 << 20


# ----------------------------------------------------------------------------- convolution
def _conv_desc(kind, x, cin, cout):
    nd, d, h, w = _spatial3(x)
    return ConvDesc(kind, nd, x.shape[0], d, h, w, cin, cout)


def _packs(weight, kind, nd):
    """Packed forward / data-gradient operands of a conv weight.  Cached ON the tensor object (an
    nn.Parameter lives as long as its module), keyed by the tensor version, the global epoch bumped
    by the fused optimiser, and the dispatch mode -- never by address, which the allocator reuses."""
    tag = (weight._version, _state["epoch"], kind, nd, lib().chap_get_force_simt())
    hit = getattr(weight, "_chap_pack", None)
    if hit is not None and hit[0] == tag:
        return hit[1], hit[2]
    transposed = kind == _lib.CONV_UP2
    cin = weight.shape[0] if transposed else weight.shape[1]
    cout = weight.shape[1] if transposed else weight.shape[0]
    desc = ConvDesc(kind, nd, 1, 2, 2, 2, cin, cout) if nd == 3 else ConvDesc(kind, nd, 1, 1, 2, 2, cin, cout)
    n = lib().chap_conv_packed_elems(ctypes.byref(desc))
    wf = torch.empty(n, dtype=torch.float32, device=weight.device)
    wd = torch.empty(n, dtype=torch.float32, device=weight.device)
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    check(lib().chap_conv_pack_weights(ctypes.byref(desc), _p(w), _p(wf), _p(wd), _stream()))
    try:
        weight._chap_pack = (tag, wf, wd)
    except AttributeError:
        pass
    return wf, wd


def _conv_kind_of(module):
    """(kind, nd) of an nn.Conv / nn.ConvTranspose parameter holder on the hot path, None for anything else."""
    import torch.nn as nn
    if isinstance(module, (nn.ConvTranspose2d, nn.ConvTranspose3d)):
        return _lib.CONV_UP2, (2 if isinstance(module, nn.ConvTranspose2d) else 3)
    if isinstance(module, (nn.Conv2d, nn.Conv3d)):
        nd = 2 if isinstance(module, nn.Conv2d) else 3
        k, s = module.kernel_size[0], module.stride[0]
        kind = {(3, 1): _lib.CONV_K3, (1, 1): _lib.CONV_K1, (2, 2): _lib.CONV_DOWN2}.get((k, s))
        return (kind, nd) if kind is not None else None
    return None


def pack_all(model):
    """Re-pack every conv weight of `model` in ONE batched launch (chap_conv_pack_weights_batched) and refresh the per-weight
    pack cache that `conv_stats` consults -- called by the trainer at the top of each iteration, right after the optimiser
    changed the weights (72 single-layer pack launches per iteration otherwise)."""
    plan = getattr(model, "_chap_pack_plan", None)
    mode = lib().chap_get_force_simt()
    if plan is None or plan["mode"] != mode:
        entries = []
        for mod in model.modules():
            kn = _conv_kind_of(mod)
            if kn is None:
                continue
            kind, nd = kn
            w = mod.weight
            transposed = kind == _lib.CONV_UP2
            cin, cout = (w.shape[0], w.shape[1]) if transposed else (w.shape[1], w.shape[0])
            desc = ConvDesc(kind, nd, 1, 2, 2, 2, cin, cout) if nd == 3 else ConvDesc(kind, nd, 1, 1, 2, 2, cin, cout)
            n = lib().chap_conv_packed_elems(ctypes.byref(desc))
            entries.append((w, kind, nd, desc, torch.empty(n, dtype=torch.float32, device=w.device),
                            torch.empty(n, dtype=torch.float32, device=w.device)))
        plan = {"mode": mode, "entries": entries, "items": (_lib.PackItem * len(entries))()}
        model._chap_pack_plan = plan
    items = plan["items"]
    for i, (w, kind, nd, desc, wf, wd) in enumerate(plan["entries"]):
        if not w.is_contiguous():
            raise RuntimeError("pack_all: parameters must be contiguous")
        items[i] = _lib.PackItem(w.data_ptr(), wf.data_ptr(), wd.data_ptr(), desc)      # p.data may have been re-pointed (flat arena)
    check(lib().chap_conv_pack_weights_batched(items, len(plan["entries"]), _stream()))
    for (w, kind, nd, desc, wf, wd) in plan["entries"]:
        w._chap_pack = ((w._version, _state["epoch"], kind, nd, mode), wf, wd)


def _out_shape(kind, x, cout):
    sp = list(x.shape[2:])
    if kind == _lib.CONV_DOWN2:
        sp = [s // 2 for s in sp]
    elif kind == _lib.CONV_UP2:
        sp = [s * 2 for s in sp]
    return [x.shape[0], cout] + sp


class _Conv(Function):
    """y = conv(cat(x, xb)) (xb optional).  With xb the channel concat of the U-Net skip connection is part of the op: the
    backward's data gradient is written straight into the two parts by the tensor-core epilogue (no split pass)."""

    @staticmethod
    def forward(ctx, x, xb, weight, bias, kind, want_stats, wf, wd, zero_bias_grad=False, bn=None):
        _require_cuda(x, weight, bias)
        # no zero tensors for the gradients of the (non-differentiable) statistics outputs: autograd would otherwise fill three
        # of them per convolution backward (216 fill kernels per 2D iteration, measured with tools/step_kernels.py)
        ctx.set_materialize_grads(False)
        x = cl(x)
        ctx.split = None
        if xb is not None:
            xb = cl(xb)
            ca, cb = x.shape[1], xb.shape[1]
            cat = empty_cl([x.shape[0], ca + cb] + list(x.shape[2:]), x.device)
            check(lib().chap_concat_channels(_p(x), _p(xb), x.numel() // ca, ca, cb, _p(cat), _stream()))
            ctx.split = (ca, cb)
            x = cat
        transposed = kind == _lib.CONV_UP2
        cin = weight.shape[0] if transposed else weight.shape[1]
        cout = weight.shape[1] if transposed else weight.shape[0]
        if x.shape[1] != cin:
            raise RuntimeError("conv: input has %d channels, weight expects %d" % (x.shape[1], cin))
        desc = _conv_desc(kind, x, cin, cout)
        y = empty_cl(_out_shape(kind, x, cout), x.device)
        fold = bn is not None and want_stats                      # +1 double: the kernel's block ticket counter
        persistent = fold and len(bn) > 7 and bn[7] is not None
        if persistent:
            sums = bn[7]            # per-layer persistent statistics buffer: zero on entry, zeroed again by the kernel's finalizing block
        else:
            sums = torch.empty(_lib.STAT_SLOTS * 2 * cout + (1 if fold else 0), dtype=torch.float64, device=x.device) if want_stats else None
        b = None if bias is None else bias.detach()
        mi = ss = None
        if fold:
            # train-mode BatchNorm follows: its finalize step (scale / shift, running statistics) runs inside the conv kernel
            gamma, beta, eps, momentum, rm, rv, nbt = bn[:7]
            mi = torch.empty(2 * cout, dtype=torch.float32, device=x.device)
            ss = torch.empty(2 * cout, dtype=torch.float32, device=x.device)
            args = _lib.BnTrainArgs(_p(gamma.detach()), _p(beta.detach()), eps, momentum, _p(rm), _p(rv), _p(nbt), _p(mi), _p(ss),
                                    1 if persistent else 0, 0)
            check(lib().chap_conv_bn_fwd(ctypes.byref(desc), _p(x), _p(wf), _p(b), _p(y), _p(sums), ctypes.byref(args), _stream()))
        else:
            check(lib().chap_conv_fwd(ctypes.byref(desc), _p(x), _p(wf), _p(b), _p(y), _p(sums), _stream()))
        ctx.desc, ctx.has_bias, ctx.wd = desc, bias is not None, wd
        ctx.zero_bias_grad = zero_bias_grad
        ctx.w_sink, ctx.b_sink = _sink_of(weight), _sink_of(bias)      # FlatSGD's gradient arena views (None outside a trainer)
        ctx.wshape = tuple(weight.shape)
        ctx.save_for_backward(x if ctx.needs_input_grad[2] else None)
        if want_stats:
            if mi is not None:
                ctx.mark_non_differentiable(sums, mi, ss)
                return y, sums, mi, ss
            ctx.mark_non_differentiable(sums)
            return y, sums, None, None
        return y, None, None, None

    @staticmethod
    def backward(ctx, dy, _dsums, _dmi, _dss):
        (x,) = ctx.saved_tensors
        desc = ctx.desc
        if dy is None:                                   # y itself unused downstream
            return (None,) * 10
        dy = cl(dy)
        dx = dxb = dw = db = None
        nd = desc.nd
        sp = [desc.in_h, desc.in_w] if nd == 2 else [desc.in_d, desc.in_h, desc.in_w]
        if ctx.split is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            ca, cb = ctx.split
            dx, dxb = empty_cl([desc.n, ca] + sp, dy.device), empty_cl([desc.n, cb] + sp, dy.device)
            if lib().chap_conv_dgrad_split_supported(ctypes.byref(desc), ca):
                check(lib().chap_conv_dgrad_split(ctypes.byref(desc), _p(dy), _p(ctx.wd), _p(dx), ca, _p(dxb), _stream()))
            else:
                both = empty_cl([desc.n, ca + cb] + sp, dy.device)
                check(lib().chap_conv_dgrad(ctypes.byref(desc), _p(dy), _p(ctx.wd), _p(both), _stream()))
                check(lib().chap_split_channels(_p(both), both.numel() // (ca + cb), ca, cb, _p(dx), _p(dxb), _stream()))
            if not ctx.needs_input_grad[0]:
                dx = None
            if not ctx.needs_input_grad[1]:
                dxb = None
        elif ctx.needs_input_grad[0]:
            dx = empty_cl([desc.n, desc.cin] + sp, dy.device)
            check(lib().chap_conv_dgrad(ctypes.byref(desc), _p(dy), _p(ctx.wd), _p(dx), _stream()))
        if ctx.needs_input_grad[2] and _state["grad_sink"] and ctx.w_sink is not None:
            # gradient-sink mode: dW (and db) are ADDED into the optimiser's flat arena by the kernels; autograd gets None
            g_view, scratch = ctx.w_sink
            want_b = ctx.has_bias and ctx.needs_input_grad[3] and not ctx.zero_bias_grad and ctx.b_sink is not None
            ws_bytes = lib().chap_conv_wgrad_workspace_bytes(ctypes.byref(desc))      # bias sums + scratch of the pair-packed 16 -> 16 path
            ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.float64, device=dy.device)
            check(lib().chap_conv_wgrad_acc(ctypes.byref(desc), _p(x), _p(dy), _p(g_view), _p(ctx.b_sink[0]) if want_b else None,
                                            _p(ws), ws_bytes, _p(scratch), _stream()))
        elif ctx.needs_input_grad[2]:
            dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dy.device)
            # A bias that feeds a train-mode BatchNorm has an analytically ZERO gradient (BN subtracts the batch mean:
            # sum_r dy = scale * (sum dz - N mean(dz) - mean(dz xhat) * sum xhat) = 0); the reference computes rounding
            # noise there.  The gradient is returned as exact zeros (weight decay still applies in the optimiser).
            want_b = ctx.has_bias and ctx.needs_input_grad[3] and not ctx.zero_bias_grad
            db = torch.empty(desc.cout, dtype=torch.float32, device=dy.device) if want_b else None
            ws_bytes = lib().chap_conv_wgrad_workspace_bytes(ctypes.byref(desc))
            ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.float64, device=dy.device)
            check(lib().chap_conv_wgrad(ctypes.byref(desc), _p(x), _p(dy), _p(dw), _p(db), _p(ws), ws_bytes, _stream()))
            if ctx.has_bias and ctx.needs_input_grad[3] and ctx.zero_bias_grad and not _state["bias_grad_none"]:
                db = torch.zeros(desc.cout, dtype=torch.float32, device=dy.device)
        return dx, dxb, dw, db, None, None, None, None, None, None


class BnStats(tuple):
    """(sums, mean_invstd, scale_shift) of a convolution whose BatchNorm finalize already ran inside the conv kernel."""


def conv_stats(x, weight, bias, kind, want_stats=True, feeds_train_bn=False, cat=None, bn=None):
    """(y, sums): sums = per-channel sum / sum-of-squares of y as float64[2*Cout] (None if not wanted).
    feeds_train_bn: y goes straight into a train-mode BatchNorm -> the bias gradient is exactly zero and is not computed.
    cat: optional second input; the convolution runs on the channel concat (x, cat)  (U-Net skip connection)."""
    _require_cuda(x, weight)
    wf, wd = _packs(weight, kind, x.dim() - 2)
    if not _state["weight_grad"]:
        weight = weight.detach()
        bias = None if bias is None else bias.detach()
    pack = None
    if bn is not None and want_stats and bn.training:
        # bn: the nn.BatchNormNd holder that consumes y -> (y, BnStats) and bn_act skips its own finalize launch
        update = bn.track_running_stats and _state["bn_tracking"]
        if update:
            _state["bn_epoch"] += 1                  # the kernel's last block updates the running statistics
        momentum = 0.1 if bn.momentum is None else bn.momentum
        pack = (bn.weight, bn.bias, float(bn.eps), float(momentum),
                bn.running_mean if update else None, bn.running_var if update else None, bn.num_batches_tracked if update else None,
                _persistent_zeros(bn, "_chap_stats", _lib.STAT_SLOTS * 2 * bn.weight.numel() + 1) if _state["persistent_stats"] else None)
    y, sums, mi, ss = _Conv.apply(x, cat, weight, bias, kind, bool(want_stats), wf, wd, bool(feeds_train_bn), pack)
    if mi is not None:
        return y, BnStats((sums, mi, ss))
    return y, sums


def conv(x, weight, bias, kind):
    return conv_stats(x, weight, bias, kind, False)[0]


def _eval_scale_shift(bn):
    """float[2C] (scale, shift) of an eval-mode BatchNorm holder, cached on the module until a parameter / running statistic
    changes (optimiser step, train-mode forward, load_state_dict)."""
    tag = (_state["epoch"], _state["bn_epoch"], bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.weight.data_ptr(), bn.running_mean.data_ptr())
    hit = getattr(bn, "_chap_eval_ss", None)
    if hit is not None and hit[0] == tag:
        return hit[1]
    c = bn.weight.numel()
    mi = torch.empty(2 * c, dtype=torch.float32, device=bn.weight.device)
    ss = torch.empty(2 * c, dtype=torch.float32, device=bn.weight.device)
    check(lib().chap_bn_eval_params(_p(bn.weight.detach()), _p(bn.bias.detach()), _p(bn.running_mean), _p(bn.running_var),
                                    float(bn.eps), _p(mi), _p(ss), c, _stream()))
    bn._chap_eval_ss = (tag, ss)
    return ss


def eval_fusable(bn):
    """True when conv -> bn -> act can run as ONE kernel: eval-mode BatchNorm with running statistics and no autograd."""
    return (not bn.training) and bn.track_running_stats and not torch.is_grad_enabled()


def conv_bn_act_eval(x, weight, bias, kind, bn, slope, residual=None, cat=None):
    """act(BatchNorm_eval(conv(x))) + residual in one kernel (chap_conv_bn_act_fwd); inference only (no autograd)."""
    _require_cuda(x, weight)
    x = cl(x.detach())
    if cat is not None:
        x = concat_channels(x, cl(cat.detach()))
    wf, _ = _packs(weight, kind, x.dim() - 2)
    transposed = kind == _lib.CONV_UP2
    cin = weight.shape[0] if transposed else weight.shape[1]
    cout = weight.shape[1] if transposed else weight.shape[0]
    desc = _conv_desc(kind, x, cin, cout)
    y = empty_cl(_out_shape(kind, x, cout), x.device)
    res = None if residual is None else cl(residual.detach())
    check(lib().chap_conv_bn_act_fwd(ctypes.byref(desc), _p(x), _p(wf), _p(None if bias is None else bias.detach()),
                                     _p(_eval_scale_shift(bn)), float(slope), _p(res), _p(y), _stream()))
    return y


# ----------------------------------------------------------------------------- BN + activation
class _BnAct(Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, residual, sums, running, train, update, slope, eps, momentum, drop_nc, drop_el, pre=None, rng=None):
        _require_cuda(y, gamma, beta)
        y = cl(y)
        n, c = y.shape[0], y.shape[1]
        rps = y.numel() // (n * c)
        dev = y.device
        g, b = gamma.detach(), beta.detach()
        if pre is not None:
            mi, ss = pre                      # finalize (and the running-statistics update) already done by the conv kernel
        else:
            mi = torch.empty(2 * c, dtype=torch.float32, device=dev)
            ss = torch.empty(2 * c, dtype=torch.float32, device=dev)
        if pre is not None:
            pass
        elif train:
            slots = _lib.STAT_SLOTS
            if sums is None:
                slots = 1
                sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
                check(lib().chap_channel_stats(_p(y), n * rps, c, _p(sums), _stream()))
            rm, rv, nbt = running if (update and running is not None) else (None, None, None)
            check(lib().chap_bn_finalize(_p(sums), slots, n * rps, _p(g), _p(b), eps, momentum, _p(rm), _p(rv), _p(nbt),
                                         _p(mi), _p(ss), c, _stream()))
        else:
            rm, rv, _ = running
            check(lib().chap_bn_eval_params(_p(g), _p(b), _p(rm), _p(rv), eps, _p(mi), _p(ss), c, _stream()))
        if residual is not None:
            residual = cl(residual)
        if drop_el is not None:
            drop_el = cl(drop_el)
        out = empty_cl(y.shape, dev)
        ctx.rng = rng
        if rng is not None:           # generated elementwise dropout: (p, seed, subsequence, epoch tensor or None)
            check(lib().chap_bn_act_fwd_rng(_p(y), _p(ss), slope, _p(drop_nc), ctypes.byref(_rng_struct(rng)), _p(residual), n, rps, c,
                                            _p(out), _stream()))
        else:
            check(lib().chap_bn_act_fwd(_p(y), _p(ss), slope, _p(drop_nc), _p(drop_el), _p(residual), n, rps, c, _p(out), _stream()))
        ctx.save_for_backward(y, ss, mi, g, drop_nc, drop_el)
        ctx.cfg = (n, rps, c, bool(train), float(slope), residual is not None)
        ctx.sinks = (_sink_of(gamma), _sink_of(beta))
        ctx.bwd_sums = getattr(gamma, "_chap_bwd_sums", None)          # persistent zeroed reduction buffer (set up by FlatSGD)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, ss, mi, g, drop_nc, drop_el = ctx.saved_tensors
        n, rps, c, train, slope, has_res = ctx.cfg
        dout = cl(dout)
        dev = dout.device
        dy = empty_cl(y.shape, dev)
        want_pg = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        rng = None if ctx.rng is None else ctypes.byref(_rng_struct(ctx.rng))
        psums = ctx.bwd_sums if _state["persistent_stats"] else None         # zero on entry, handed back zeroed: no zero-fill launch
        sink = want_pg and _state["grad_sink"] and ctx.sinks[0] is not None and ctx.sinks[1] is not None
        dres = dout if (has_res and ctx.needs_input_grad[3]) else None
        if sink or (not want_pg and psums is not None):
            # parameter gradients ADDED into the optimiser's arena (sink), or none wanted at all (the feature-gradient probe of VAT2d)
            sums = psums if psums is not None else torch.empty(2 * c, dtype=torch.float64, device=dev)
            dg, db = (ctx.sinks[0][0], ctx.sinks[1][0]) if sink else (None, None)
            if rng is not None:
                check(lib().chap_bn_act_bwd_rng(_p(dout), _p(y), _p(ss), _p(mi), slope, _p(drop_nc), rng, n, rps, c, 1 if train else 0, _p(sums),
                                                1 if psums is not None else 0, _p(dy), _p(dg), _p(db), 1, _stream()))
            else:
                check(lib().chap_bn_act_bwd_acc(_p(dout), _p(y), _p(ss), _p(mi), slope, _p(drop_nc), _p(drop_el), n, rps, c,
                                                1 if train else 0, _p(sums), 1 if psums is not None else 0, _p(dy), _p(dg), _p(db), _stream()))
            return (dy, None, None, dres) + (None,) * 11
        sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
        dgamma = torch.empty(c, dtype=torch.float32, device=dev) if want_pg else None
        dbeta = torch.empty(c, dtype=torch.float32, device=dev) if want_pg else None
        if rng is not None:
            check(lib().chap_bn_act_bwd_rng(_p(dout), _p(y), _p(ss), _p(mi), slope, _p(drop_nc), rng, n, rps, c, 1 if train else 0, _p(sums), 0,
                                            _p(dy), _p(dgamma), _p(dbeta), 0, _stream()))
        else:
            check(lib().chap_bn_act_bwd(_p(dout), _p(y), _p(ss), _p(mi), _p(g), slope, _p(drop_nc), _p(drop_el), n, rps, c,
                                        1 if train else 0, _p(sums), _p(dy), _p(dgamma), _p(dbeta), _stream()))
        return (dy, dgamma, dbeta, dres) + (None,) * 11


def _rng_struct(rng):
    p, seed, sub, epoch = rng
    return _lib.DropoutRng(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sub) & 0xFFFFFFFFFFFFFFFF, _p(epoch))


def set_generated_dropout(flag):
    """Elementwise dropout masks generated inside the BatchNorm / activation kernels (default on) instead of bernoulli_ mask tensors."""
    _state["drop_rng"] = bool(flag)


def set_dropout_epoch(epoch_dev):
    """int64 CUDA tensor (or None) mixed into the dropout generator's key at RUN time: a trainer hands in its device-side iteration
    counter so that every replay of a captured iteration draws new masks."""
    _state["drop_epoch"] = epoch_dev


def dropout_rng(p, c):
    """Generator description for bn_act(drop_rng=...) or None when the generated form does not apply (inactive, unsupported layout)."""
    if not _state.get("drop_rng", True) or not (0.0 < p < 1.0) or c % 4 != 0 or 256 % (c // 4) != 0:
        return None
    _state["drop_calls"] = _state.get("drop_calls", 0) + 1
    seed = torch.initial_seed()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        seed += 0x9E3779B97F4A7C15 * torch.distributed.get_rank()      # data-parallel ranks draw different masks from the same torch seed
    return (p, seed, _state["drop_calls"], _state.get("drop_epoch"))


def bn_act(y, bn, slope, sums=None, residual=None, drop_nc=None, drop_el=None, drop_rng=None):
    """act(BatchNorm(y)) * drop + residual with the semantics of the nn.BatchNormNd holder `bn`
    (train/eval, eps, momentum, running statistics).  drop_rng: dropout_rng(p, C) description -- nn.Dropout(p) with the mask
    generated inside the kernels (exclusive with drop_el)."""
    train = bn.training or not bn.track_running_stats
    running = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if bn.track_running_stats else None
    update = bn.training and bn.track_running_stats and _state["bn_tracking"]
    if update:
        _state["bn_epoch"] += 1                      # running statistics change: cached eval scale / shift are stale
    momentum = 0.1 if bn.momentum is None else bn.momentum
    pre = None
    if isinstance(sums, BnStats):
        sums, pre = sums[0], (sums[1], sums[2])
    return _BnAct.apply(y, bn.weight, bn.bias, residual, sums, running, train, update, float(slope), float(bn.eps),
                        float(momentum), drop_nc, drop_el, pre, drop_rng)


# ----------------------------------------------------------------------------- pooling / upsampling / concat
class _MaxPool2(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = cl(x)
        n, c, h, w = x.shape
        y = empty_cl([n, c, h // 2, w // 2], x.device)
        check(lib().chap_maxpool2_fwd(_p(x), n, h, w, c, _p(y), _stream()))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = cl(dy)
        dx = empty_cl(x.shape, x.device)
        check(lib().chap_maxpool2_bwd(_p(x), _p(dy), n, h, w, c, _p(dx), _stream()))
        return dx


def maxpool2(x):
    return _MaxPool2.apply(x)


class _Upsample2x(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = cl(x)
        nd, d, h, w = _spatial3(x)
        n, c = x.shape[0], x.shape[1]
        y = empty_cl([n, c] + [2 * s for s in x.shape[2:]], x.device)
        check(lib().chap_upsample2x_fwd(_p(x), nd, n, d, h, w, c, _p(y), _stream()))
        ctx.cfg = (nd, n, d, h, w, c, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        nd, n, d, h, w, c, shape = ctx.cfg
        dy = cl(dy)
        dx = empty_cl(list(shape), dy.device)
        check(lib().chap_upsample2x_bwd(_p(dy), nd, n, d, h, w, c, _p(dx), _stream()))
        return dx


def upsample2x(x):
    """bilinear (4D) / trilinear (5D) x2, align_corners=True."""
    return _Upsample2x.apply(x)


class _Concat(Function):
    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        a, b = cl(a), cl(b)
        ca, cb = a.shape[1], b.shape[1]
        rows = a.numel() // ca
        out = empty_cl([a.shape[0], ca + cb] + list(a.shape[2:]), a.device)
        check(lib().chap_concat_channels(_p(a), _p(b), rows, ca, cb, _p(out), _stream()))
        ctx.cfg = (tuple(a.shape), tuple(b.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        sa, sb = ctx.cfg
        dout = cl(dout)
        rows = dout.numel() // dout.shape[1]
        da = empty_cl(list(sa), dout.device) if ctx.needs_input_grad[0] else None
        db = empty_cl(list(sb), dout.device) if ctx.needs_input_grad[1] else None
        if da is not None or db is not None:
            check(lib().chap_split_channels(_p(dout), rows, sa[1], sb[1], _p(da), _p(db), _stream()))
        return da, db


def concat_channels(a, b):
    return _Concat.apply(a, b)


class _ChannelScale(Function):
    @staticmethod
    def forward(ctx, x, scale_nc):
        _require_cuda(x, scale_nc)
        x = cl(x)
        s = scale_nc.detach().reshape(x.shape[0], x.shape[1]).float().contiguous()
        n, c = x.shape[0], x.shape[1]
        out = empty_cl(x.shape, x.device)
        check(lib().chap_channel_scale(_p(x), _p(s), n, x.numel() // (n * c), c, _p(out), _stream()))
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, dout):
        (s,) = ctx.saved_tensors
        dout = cl(dout)
        n, c = dout.shape[0], dout.shape[1]
        dx = empty_cl(dout.shape, dout.device)
        check(lib().chap_channel_scale(_p(dout), _p(s), n, dout.numel() // (n * c), c, _p(dx), _stream()))
        return dx, None


def channel_scale(x, scale_nc):
    """x * scale[n, c] (Dropout2d/3d-style channel masks; the mask itself is not differentiated)."""
    return _ChannelScale.apply(x, scale_nc)


class _FeatureDropout(Function):
    @staticmethod
    def forward(ctx, feat, m1, m2, nu):
        _require_cuda(feat, m1, m2)
        feat = cl(feat)
        n, c = feat.shape[0], feat.shape[1]
        rps = feat.numel() // (n * c)
        m1 = None if m1 is None else m1.detach().reshape(nu, c).float().contiguous()
        m2 = None if m2 is None else m2.detach().reshape(nu, c).float().contiguous()
        shape = [n + nu, c] + list(feat.shape[2:])
        out1, out2 = empty_cl(shape, feat.device), empty_cl(shape, feat.device)
        check(lib().chap_feature_dropout_fwd(_p(feat), _p(m1), _p(m2), n, nu, rps, c, _p(out1), _p(out2), _stream()))
        ctx.save_for_backward(m1, m2)
        ctx.cfg = (n, nu, rps, c, tuple(feat.shape))
        return out1, out2

    @staticmethod
    def backward(ctx, d1, d2):
        m1, m2 = ctx.saved_tensors
        n, nu, rps, c, shape = ctx.cfg
        d1 = None if d1 is None else cl(d1)
        d2 = None if d2 is None else cl(d2)
        dev = (d1 if d1 is not None else d2).device
        dfeat = empty_cl(list(shape), dev)
        check(lib().chap_feature_dropout_bwd(_p(d1), _p(d2), _p(m1), _p(m2), n, nu, rps, c, _p(dfeat), _stream()))
        return dfeat, None, None, None


def feature_dropout(feat, m1, m2, nu):
    """(cat(feat, feat[-nu:] * m1), cat(feat, feat[-nu:] * m2)) along the batch, m [nu, C] or None (= 1), in one fused pass."""
    return _FeatureDropout.apply(feat, m1, m2, int(nu))


def axpy(a, b, alpha):
    """a + alpha * b (flat, no autograd)."""
    _require_cuda(a, b)
    a, b = cl(a.detach()), cl(b.detach())
    out = torch.empty_like(a)
    check(lib().chap_axpy(_p(a), _p(b), float(alpha), a.numel(), _p(out), _stream()))
    return out


def mask_mix(a, b, mask):
    """a*m + b*(1-m) with an int64 spatial mask broadcast over batch and channels (no grad)."""
    _require_cuda(a, b, mask)
    a, b = cl(a.detach()), cl(b.detach())
    m = mask.to(torch.int64).contiguous()
    n, c = a.shape[0], a.shape[1]
    out = torch.empty_like(a)
    check(lib().chap_mask_mix(_p(a), _p(b), _p(m), n, a.numel() // (n * c), c, _p(out), _stream()))
    return out


# ----------------------------------------------------------------------------- losses
def _rows_c(logits):
    c = logits.shape[1]
    return logits.numel() // c, c


def pseudo_label(pre1, pre2):
    """softmax1, softmax2, argmax1, argmax2 (int64), knowledge = CE(pre1, arg2) + CE(pre2, arg1)."""
    _require_cuda(pre1, pre2)
    pre1, pre2 = cl(pre1.detach()), cl(pre2.detach())
    rows, c = _rows_c(pre1)
    sp = [pre1.shape[0]] + list(pre1.shape[2:])
    s1, s2 = torch.empty_like(pre1), torch.empty_like(pre2)
    a1 = torch.empty(sp, dtype=torch.int64, device=pre1.device)
    a2 = torch.empty(sp, dtype=torch.int64, device=pre1.device)
    k = torch.empty(sp, dtype=torch.float32, device=pre1.device)
    check(lib().chap_pseudo_label(_p(pre1), _p(pre2), rows, c, _p(s1), _p(s2), _p(a1), _p(a2), _p(k), _stream()))
    return s1, s2, a1, a2, k


def softmax(logits):
    _require_cuda(logits)
    x = cl(logits.detach())
    rows, c = _rows_c(x)
    out = torch.empty_like(x)
    check(lib().chap_softmax(_p(x), rows, c, _p(out), _stream()))
    return out


def argmax(a, b=None):
    """argmax over classes of softmax(a) or of softmax((a + b) / 2); int64 [N, *spatial]."""
    _require_cuda(a, b)
    a = cl(a.detach())
    b = None if b is None else cl(b.detach())
    rows, c = _rows_c(a)
    out = torch.empty([a.shape[0]] + list(a.shape[2:]), dtype=torch.int64, device=a.device)
    check(lib().chap_argmax(_p(a), _p(b), rows, c, _p(out), _stream()))
    return out


def _label_arg(labels):
    if labels.dtype == torch.int64:
        return labels.contiguous(), _lib.LABEL_I64
    return labels.float().contiguous(), _lib.LABEL_F32


class _DiceCeSums(Function):
    @staticmethod
    def forward(ctx, logits, labels, mask, invert):
        _require_cuda(logits, labels, mask)
        x = cl(logits)
        n, c = x.shape[0], x.shape[1]
        rps = x.numel() // (n * c)
        lab, dt = _label_arg(labels)
        m = mask.to(torch.int64).contiguous()
        if m.numel() != rps:
            raise RuntimeError("dice_ce: mask must have one entry per spatial position")
        sums = torch.empty(3 * c + 2, dtype=torch.float64, device=x.device)
        check(lib().chap_dice_ce_fwd(_p(x), _p(lab), dt, _p(m), invert, n, rps, c, _p(sums), _stream()))
        ctx.save_for_backward(x, lab, m)
        ctx.cfg = (dt, invert, n, rps, c)
        return sums

    @staticmethod
    def backward(ctx, dsums):
        x, lab, m = ctx.saved_tensors
        dt, invert, n, rps, c = ctx.cfg
        coef = torch.cat([dsums[:2 * c], dsums[3 * c:3 * c + 1]]).float().contiguous()
        dl = torch.empty_like(x)
        check(lib().chap_dice_ce_bwd(_p(x), _p(lab), dt, _p(m), invert, n, rps, c, _p(coef), 0, _p(dl), _stream()))
        return dl, None, None, None


def dice_ce_sums(logits, labels, mask, invert=False):
    """float64[3C+2]: inter[C], sum(s^2 m)[C], sum(t m)[C], sum(CE m), sum(m); differentiable in logits."""
    return _DiceCeSums.apply(logits, labels, mask, 1 if invert else 0)


class _MixLoss(Function):
    """mix_loss (code/train_ours_2D.py:198-216) as ONE autograd node: two fused masked Dice + CE passes over the logits, a
    one-block kernel for the scalar tail, and in the backward one coefficient kernel + two passes accumulating into dlogits."""

    @staticmethod
    def forward(ctx, logits, img_l, patch_l, mask, w_img, w_patch):
        _require_cuda(logits, img_l, patch_l, mask)
        x = cl(logits)
        n, c = x.shape[0], x.shape[1]
        rps = x.numel() // (n * c)
        lab1, dt1 = _label_arg(img_l)
        lab2, dt2 = _label_arg(patch_l)
        m = mask.to(torch.int64).contiguous()
        if m.numel() != rps:
            raise RuntimeError("mix_loss: mask must have one entry per spatial position")
        sums = torch.empty(2, 3 * c + 2, dtype=torch.float64, device=x.device)
        check(lib().chap_dice_ce_fwd(_p(x), _p(lab1), dt1, _p(m), 0, n, rps, c, _p(sums[0]), _stream()))
        check(lib().chap_dice_ce_fwd(_p(x), _p(lab2), dt2, _p(m), 1, n, rps, c, _p(sums[1]), _stream()))
        out = torch.empty(3, dtype=torch.float32, device=x.device)
        check(lib().chap_mix_loss_finalize(_p(sums[0]), _p(sums[1]), c, w_img, w_patch, _p(out), _stream()))
        ctx.save_for_backward(x, lab1, lab2, m, sums)
        ctx.cfg = (dt1, dt2, n, rps, c, w_img, w_patch)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, lab1, lab2, m, sums = ctx.saved_tensors
        dt1, dt2, n, rps, c, w_img, w_patch = ctx.cfg
        g = dout.float().contiguous()
        coef = torch.empty(2, 2 * c + 1, dtype=torch.float32, device=x.device)
        check(lib().chap_mix_loss_coef(_p(sums[0]), _p(sums[1]), c, w_img, w_patch, _p(g), _p(coef[0]), _p(coef[1]), _stream()))
        dl = torch.empty_like(x)
        check(lib().chap_dice_ce_bwd(_p(x), _p(lab1), dt1, _p(m), 0, n, rps, c, _p(coef[0]), 0, _p(dl), _stream()))
        check(lib().chap_dice_ce_bwd(_p(x), _p(lab2), dt2, _p(m), 1, n, rps, c, _p(coef[1]), 1, _p(dl), _stream()))
        return dl, None, None, None, None, None


def mix_loss_fused(logits, img_l, patch_l, mask, w_img, w_patch):
    """float32[3] = (loss_image, loss_patch, total); see _MixLoss."""
    return _MixLoss.apply(logits, img_l, patch_l, mask, float(w_img), float(w_patch))


class _ConsistencySums(Function):
    @staticmethod
    def forward(ctx, logits, target, mask, dist):
        _require_cuda(logits, target, mask)
        x, t = cl(logits), cl(target.detach())
        rows, c = _rows_c(x)
        m = None if mask is None else mask.detach().float().contiguous()
        sums = torch.empty(3 * c + 1, dtype=torch.float64, device=x.device)
        check(lib().chap_consistency_fwd(_p(x), _p(t), _p(m), dist, rows, c, _p(sums), _stream()))
        ctx.save_for_backward(x, t, m)
        ctx.cfg = (dist, rows, c)
        return sums

    @staticmethod
    def backward(ctx, dsums):
        x, t, m = ctx.saved_tensors
        dist, rows, c = ctx.cfg
        coef = dsums[:2 * c].float().contiguous()
        dl = torch.empty_like(x)
        check(lib().chap_consistency_bwd(_p(x), _p(t), _p(m), dist, rows, c, _p(coef), _p(dl), _stream()))
        return dl, None, None, None


def consistency_sums(logits, target, mask, dist):
    return _ConsistencySums.apply(logits, target, mask, dist)


def patch_topk_mask(knowledge, arg1, arg2, scale_factor, topk):
    """create_maskV1 (frozen spec, oracle/chap_losses.py create_mask_v1): float mask [N, *spatial]."""
    _require_cuda(knowledge, arg1, arg2)
    k = knowledge.detach().float().contiguous()
    a1, a2 = arg1.contiguous(), arg2.contiguous()
    n = k.shape[0]
    nd = k.dim() - 1
    d, h, w = (1, k.shape[1], k.shape[2]) if nd == 2 else tuple(k.shape[1:])
    s = scale_factor
    npatch = (d // s if nd == 3 else 1) * (h // s) * (w // s)
    score = torch.empty(n, npatch, dtype=torch.float32, device=k.device)
    check(lib().chap_patch_score(_p(k), _p(a1), _p(a2), nd, n, d, h, w, s, _p(score), _stream()))
    kk = max(1, int(topk * npatch))
    kth = torch.topk(score, kk, dim=1).values[:, -1].contiguous()     # selection on a [N, P] tensor: plumbing
    mask = torch.empty_like(k)
    check(lib().chap_patch_mask(_p(score), _p(kth), nd, n, d, h, w, s, _p(mask), _stream()))
    return mask


def largest_cc(seg, n_classes):
    """get_ACDC_2DLargestCC on the device: int64 class map [N, *spatial] -> float32 map keeping, per sample and
    foreground class, only the largest connected component (full connectivity, first component wins ties)."""
    _require_cuda(seg)
    seg = seg.to(torch.int64).contiguous()
    n = seg.shape[0]
    nd = seg.dim() - 1
    d, h, w = (1, seg.shape[1], seg.shape[2]) if nd == 2 else tuple(seg.shape[1:])
    nbytes = lib().chap_largest_cc_workspace_bytes(n, d, h, w, n_classes)
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=seg.device)
    out = torch.empty(seg.shape, dtype=torch.float32, device=seg.device)
    check(lib().chap_largest_cc(_p(seg), nd, n, d, h, w, n_classes, _p(out), _p(ws), nbytes, _stream()))
    return out


# ----------------------------------------------------------------------------- perturbation generator
def perturb(grads, feats, eps, mode="channel_spatial", g_scale=1.0):
    """[f_l + eps * normalise(g_l)] for all levels in one library call (f_l may be None -> r_l only).
    g is multiplied by g_scale on load (VAT differentiates w.r.t. f + xi*d, so dL/dd = xi * dL/d(f + xi d)).
    The perturbation is a constant w.r.t. autograd (g is detached by the frozen spec); the caller restores
    d out / d f = I with attach_identity_grad()."""
    n = grads[0].shape[0]
    levels = (_lib.Level * len(grads))()
    keep, outs = [], []
    for i, g in enumerate(grads):
        _require_cuda(g)
        g = cl(g.detach())
        f = None if feats is None or feats[i] is None else cl(feats[i].detach())
        out = torch.empty_like(g)
        c = g.shape[1]
        levels[i] = _lib.Level(g.data_ptr(), 0 if f is None else f.data_ptr(), out.data_ptr(), g.numel() // (n * c), c, 0)
        keep += [g, f]
        outs.append(out)
    n_ws = lib().chap_perturb_workspace_elems(levels, len(grads), n)
    ws = torch.empty(n_ws, dtype=torch.float64, device=grads[0].device)
    check(lib().chap_perturb_fwd(levels, len(grads), n, _lib.PERTURB_MODES[mode], float(eps), float(g_scale), _p(ws), n_ws,
                                 _stream()))
    return outs


class _ReplaceValue(Function):
    """Forward: returns `value` (computed from f by a kernel as f + const); backward: identity to f."""
    @staticmethod
    def forward(ctx, f, value):
        return value.view_as(value)

    @staticmethod
    def backward(ctx, dout):
        return dout, None


def attach_identity_grad(f, value):
    return _ReplaceValue.apply(f, value)


def l2n_sample_axpy(d, base, xi):
    """base + xi * d / (||d||_2 per sample + 1e-8); no autograd (inputs are leaves of VAT step 2)."""
    _require_cuda(d, base)
    d = cl(d.detach())
    base = None if base is None else cl(base.detach())
    n = d.shape[0]
    norms = torch.empty(n, dtype=torch.float64, device=d.device)
    out = torch.empty_like(d)
    check(lib().chap_l2n_sample_axpy(_p(d), _p(base), float(xi), n, d.numel() // n, _p(norms), _p(out), _stream()))
    return out


def l2n_sample_axpy_all(ds, bases, xi):
    """[base_l + xi * d_l / (||d_l||_2 per sample + 1e-8)] for all levels in TWO launches (VAT step 2); no autograd."""
    n = ds[0].shape[0]
    levels = (_lib.Level * len(ds))()
    keep, outs = [], []
    for i, d in enumerate(ds):
        _require_cuda(d)
        d = cl(d.detach())
        base = None if bases is None or bases[i] is None else cl(bases[i].detach())
        out = torch.empty_like(d)
        levels[i] = _lib.Level(d.data_ptr(), 0 if base is None else base.data_ptr(), out.data_ptr(), d.numel() // n, 1, 0)
        keep += [d, base]
        outs.append(out)
    norms = torch.empty(len(ds) * n, dtype=torch.float64, device=ds[0].device)
    check(lib().chap_l2n_sample_axpy_batched(levels, len(ds), n, float(xi), _p(norms), _stream()))
    return outs


# ----------------------------------------------------------------------------- optimiser
def sgd_momentum_(flat_p, flat_g, flat_buf, lr, momentum, weight_decay, grad_scale=1.0, first_step=False):
    _require_cuda(flat_p, flat_g, flat_buf)
    check(lib().chap_sgd_momentum(_p(flat_p), _p(flat_g), _p(flat_buf), flat_p.numel(), float(lr), float(momentum),
                                  float(weight_decay), float(grad_scale), 1 if first_step else 0, _stream()))
    invalidate_weight_cache()


def sgd_momentum_lrdev_(flat_p, flat_g, flat_buf, lr_dev, momentum, weight_decay, grad_scale=1.0):
    """SGD-momentum with the learning rate in a device tensor (CUDA-graph replayable); flat_buf starts as zeros."""
    _require_cuda(flat_p, flat_g, flat_buf, lr_dev)
    check(lib().chap_sgd_momentum_lrdev(_p(flat_p), _p(flat_g), _p(flat_buf), flat_p.numel(), _p(lr_dev), float(momentum),
                                        float(weight_decay), float(grad_scale), _stream()))
    invalidate_weight_cache()


def schedule_step(iter_dev, base_lr, max_iterations, consistency, rampup, lr_dev, cw_dev, ramp_div=150):
    """lr_dev <- poly LR, cw_dev <- consistency weight for the iteration held in iter_dev (int64, device), then
    iter_dev += 1 -- all on the device (code/train_ours_2D.py:356,387)."""
    _require_cuda(iter_dev, lr_dev, cw_dev)
    check(lib().chap_schedule_step(_p(iter_dev), float(base_lr), float(max_iterations), float(consistency), float(rampup),
                                   int(ramp_div), _p(lr_dev), _p(cw_dev), _stream()))


# ----------------------------------------------------------------------------- 2D validation helpers
def zoom_index(in_size, out_size):
    """scipy.ndimage.zoom(order=0, mode='constant', grid_mode=False) index map for one axis: output o reads input
    floor(o * (in - 1) / (out - 1) + 0.5), evaluated in double like ni_interpolation.c (NI_ZoomShift) does; -1 = fill value 0."""
    import numpy as np
    if out_size <= 1:
        return np.zeros(max(out_size, 0), dtype=np.int32)
    ratio = float(in_size - 1) / float(out_size - 1)
    cc = np.arange(out_size, dtype=np.float64) * ratio
    idx = np.floor(cc + 0.5).astype(np.int32)
    # mode='constant' (the default the reference uses): a coordinate outside [0, in - 1] -- which happens for the LAST output
    # when o * ratio overshoots in - 1 by one ulp, e.g. 256 -> 200 -- reads the fill value cval = 0 (map_coordinate, NI_EXTEND_CONSTANT)
    idx[(cc < 0) | (cc > in_size - 1)] = -1
    return idx


def zoom_nearest(x, out_h, out_w):
    """scipy.ndimage.zoom(x[s], (out_h / h, out_w / w), order=0) for every slice of a CUDA stack [S, h, w] (float32 or int64)."""
    _require_cuda(x)
    if x.dtype not in (torch.float32, torch.int64):
        raise RuntimeError("zoom_nearest: float32 or int64 stacks only")
    x = x.contiguous()
    s, h, w = x.shape
    iy = torch.from_numpy(zoom_index(h, out_h)).to(x.device)
    ix = torch.from_numpy(zoom_index(w, out_w)).to(x.device)
    out = torch.empty((s, out_h, out_w), dtype=x.dtype, device=x.device)
    check(lib().chap_gather2d(_p(x), x.element_size(), _p(iy), _p(ix), s, h, w, out_h, out_w, _p(out), _stream()))
    return out


def label_overlap(pred, gt, classes):
    """int64 [classes, 3] = (|pred == c & gt == c|, |pred == c|, |gt == c|) for two int64 CUDA label volumes."""
    _require_cuda(pred, gt)
    pred, gt = pred.to(torch.int64).contiguous(), gt.to(torch.int64).contiguous()
    if pred.numel() != gt.numel():
        raise RuntimeError("label_overlap: shapes differ")
    counts = torch.empty((classes, 3), dtype=torch.int64, device=pred.device)
    check(lib().chap_label_overlap(_p(pred), _p(gt), pred.numel(), classes, _p(counts), _stream()))
    return counts


# ----------------------------------------------------------------------------- sliding window
def sw_desc(vol, patch, nwin, stride, c):
    d = _lib.SwDesc()
    for a in range(3):
        d.vol[a], d.patch[a], d.nwin[a], d.stride[a] = int(vol[a]), int(patch[a]), int(nwin[a]), int(stride[a])
    d.c = int(c)
    return d


def sw_extract(desc, volume, first, count):
    _require_cuda(volume)
    out = torch.empty([count, 1, desc.patch[0], desc.patch[1], desc.patch[2]], dtype=torch.float32, device=volume.device)
    check(lib().chap_sw_extract(ctypes.byref(desc), _p(volume), first, count, _p(out), _stream()))
    return out


def sw_aggregate(desc, win, is_prob, want_maps=False):
    _require_cuda(win)
    dev = win.device
    vol = (desc.vol[0], desc.vol[1], desc.vol[2])
    label = torch.empty(vol, dtype=torch.int64, device=dev)
    score = torch.empty((desc.c,) + vol, dtype=torch.float32, device=dev) if want_maps else None
    cnt = torch.empty(vol, dtype=torch.float32, device=dev) if want_maps else None
    check(lib().chap_sw_aggregate(ctypes.byref(desc), _p(win), 1 if is_prob else 0, _p(score), _p(cnt), _p(label), _stream()))
    return (label, score, cnt) if want_maps else label

"""TEST INFRASTRUCTURE.  "Ideal TF32" yardstick: the oracle networks evaluated in float64 with every convolution that the
product runs on the tensor cores given TF32-rounded operands (cvt.rna emulation) -- forward (x, w), data gradient (dy, w) and
weight gradient (x, dy) -- and exact accumulation.  This is what ANY faithful TF32 execution converges to; the TF32 mode of
the product is asserted against the deviation of this evaluation from the unrounded one, instead of against a constant.

(torch-eager + cuDNN with allow_tf32 is NOT such a yardstick: measured on B200, cuDNN runs its fp32 kernels on the
16-channel layers, see profiles/r02_tf32_probe.md.)"""
import contextlib

import torch
import torch.nn.functional as F


def rna(x):
    """cvt.rna.tf32.f32 on the fp32 image of x (any float dtype), returned in x's dtype."""
    i = x.detach().float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)


class _RoundFwd(torch.autograd.Function):
    """forward: rna(x); backward: identity (the rounding is part of the conv arithmetic, not of the function)."""
    @staticmethod
    def forward(ctx, x):
        return rna(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """forward: identity; backward: rna(g) -- the dy operand of the data / weight gradient GEMMs."""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return rna(g)


def _on_tensor_cores(cin, cout, ksize):
    # product dispatch (chap_b200/csrc/conv_tc.cu tc_supports): Cin = 1 stems and the 1x1x1 class head run on the CUDA cores
    return cin >= 16 and (cout >= 16 or ksize == 3)


@contextlib.contextmanager
def ideal_tf32():
    orig = {n: getattr(F, n) for n in ("conv2d", "conv3d", "conv_transpose2d", "conv_transpose3d")}

    def wrap(name, transposed):
        fn = orig[name]

        def conv(x, w, b=None, *a, **kw):
            cin = w.shape[0] if transposed else w.shape[1]
            cout = w.shape[1] if transposed else w.shape[0]
            if not _on_tensor_cores(cin, cout, w.shape[-1]):
                return fn(x, w, b, *a, **kw)
            y = fn(_RoundFwd.apply(x), _RoundFwd.apply(w), None, *a, **kw)
            y = _RoundBwd.apply(y)
            if b is not None:
                y = y + b.reshape((1, -1) + (1,) * (y.dim() - 2))
            return y
        return conv
    try:
        for n in orig:
            setattr(F, n, wrap(n, "transpose" in n))
        yield
    finally:
        for n, fn in orig.items():
            setattr(F, n, fn)

"""TEST / MEASUREMENT TOOL (uses oracle/).  Settles what tcgen05.mma.kind::tf32 does with the low 13 mantissa bits of its fp32 operands.

One convolution layer, operands pre-rounded to TF32 on the host (emulated cvt.rna) or left raw, compared with an fp64
evaluation; then the whole 2D / 3D network in the library's rounding modes against the fp64 oracle and against
torch-eager + cuDNN-TF32 on the same GPU.  Writes gpurun_out/tf32_probe.json.

    python tools/tf32_probe.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chap_b200 import networks, ops  # noqa: E402
from chap_b200._lib import CONV_K3  # noqa: E402
from oracle import nets as onets  # noqa: E402

DEV = "cuda:0"


def rna(x):
    """cvt.rna.tf32.f32 emulated: add half an ulp of the 10-bit mantissa to the magnitude, clear the low 13 bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def stats(o, ref):
    o, ref = o.double().flatten(), ref.double().flatten()
    return {"rel_l2": float((o - ref).norm() / ref.norm()), "shrink": float((o @ ref) / (ref @ ref) - 1.0)}


def one_layer(c, hw, out):
    torch.manual_seed(0)
    x = torch.randn(4, c, hw, hw, device=DEV).abs_() + 0.1           # post-activation like: positive, so a truncation bias shows
    w = torch.randn(c, c, 3, 3, device=DEV) / (3.0 * c ** 0.5)
    b = torch.zeros(c, device=DEV)
    ref = F.conv2d(x.double(), w.double(), None, padding=1)
    res = {}
    for xa, xn in ((x, "x_raw"), (rna(x), "x_rna"), (trunc(x), "x_trunc")):
        for wa, wn in ((w, "w_raw"), (rna(w), "w_rna"), (trunc(w), "w_trunc")):
            y = ops.conv(ops.cl(xa), wa.clone(), b, CONV_K3)
            res["tcgen05 %s %s" % (xn, wn)] = stats(y, ref)
    ops.set_conv_precision(ops.PRECISE_ALL)
    res["tcgen05 3xTF32 (split operands)"] = stats(ops.conv(ops.cl(x), w.clone(), b, CONV_K3), ref)
    ops.set_conv_precision(0)
    # emulations on exact arithmetic: what "hardware truncates" / "hardware rounds" would give
    res["emul trunc both"] = stats(F.conv2d(trunc(x).double(), trunc(w).double(), None, padding=1), ref)
    res["emul rna both"] = stats(F.conv2d(rna(x).double(), rna(w).double(), None, padding=1), ref)
    res["emul rna x, trunc w"] = stats(F.conv2d(rna(x).double(), trunc(w).double(), None, padding=1), ref)
    torch.backends.cudnn.allow_tf32 = True
    res["cudnn tf32"] = stats(F.conv2d(x, w, None, padding=1), ref)
    torch.backends.cudnn.allow_tf32 = False
    res["cudnn fp32"] = stats(F.conv2d(x, w, None, padding=1), ref)
    out["layer c%d %dx%d" % (c, hw, hw)] = res


def network(dims, out):
    torch.manual_seed(1337)
    if dims == 2:
        model = networks.DualDecoder(1, 4, {"decoder_type": "mcnet"})
        x = torch.rand(4, 1, 128, 128)
    else:
        model = networks.DualDecoder3d(1, 2, normalization="batchnorm", has_dropout=False)
        x = torch.randn(2, 1, 48, 48, 32)
    for m in model.modules():
        if hasattr(m, "dropout_p"):
            m.dropout_p = 0.0
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in onets.clone_state_dict(model.state_dict(), device=DEV).items()}
    sd32 = onets.clone_state_dict(model.state_dict(), device=DEV)
    fwd = onets.dualdecoder2d_forward if dims == 2 else onets.dualdecoder3d_forward
    xd = x.to(DEV)
    with torch.no_grad():
        ref = fwd(sd64, xd.double(), True, False)
        torch.backends.cudnn.allow_tf32 = True
        cud = fwd(sd32, xd, True, False)
        torch.backends.cudnn.allow_tf32 = False
        c32 = fwd(sd32, xd, True, False)
    model = model.to(DEV).train()
    res = {"cudnn tf32": [stats(a, r) for a, r in zip(cud, ref)], "cudnn fp32": [stats(a, r) for a, r in zip(c32, ref)]}
    for c in (0, 16, 32, 64, 128, ops.PRECISE_ALL):
        ops.set_conv_precision(c)
        with torch.no_grad(), ops.bn_tracking(False):
            o = model(xd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model(xd)
            e1.record()
            torch.cuda.synchronize()
        res["tcgen05 3xTF32 for C <= %d" % c] = [stats(a, r) for a, r in zip(o, ref)] + [{"fwd_ms": e0.elapsed_time(e1) / 5}]
    ops.set_conv_precision(0)
    out["network %dD" % dims] = res


def main():
    out = {}
    for c, hw in ((16, 64), (64, 32), (128, 16)):
        one_layer(c, hw, out)
    network(2, out)
    network(3, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tf32_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k)
        for kk, vv in v.items():
            print("   %-34s %s" % (kk, vv))


if __name__ == "__main__":
    main()

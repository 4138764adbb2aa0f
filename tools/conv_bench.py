"""Per-layer convolution micro-benchmark (developer tool, not part of the product path).

    python tools/conv_bench.py "k3 16 16 12 256 256" "k3 32 16 12 256 256" ...      # kind cin cout n h w [d]
    CHAP_TC_DEBUG=... python tools/conv_bench.py --env CHAP_NO_PERSIST=1 "k3 16 16 12 256 256"

For every layer: forward (with the BatchNorm-statistics epilogue), data gradient and weight gradient are launched
`--reps` times through the C ABI and timed by the library's own CUDA-event timers (chap_timing_report), so the numbers
are per kernel launch.  Inputs are larger than L2 for the level-0/1 layers; for the small ones the rotation over
`--bufs` input buffers keeps them from being served from L2.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from chap_b200 import _lib, ops  # noqa: E402

KINDS = {"k3": _lib.CONV_K3, "k1": _lib.CONV_K1, "down2": _lib.CONV_DOWN2, "up2": _lib.CONV_UP2}


def run_layer(spec, reps, bufs, dgrad=True):
    f = spec.split()
    kind, cin, cout, n, h, w = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5])
    d = int(f[6]) if len(f) > 6 else None
    nd = 3 if d else 2
    k = {"k3": 3, "k1": 1, "down2": 2, "up2": 2}[kind]
    dev = torch.device("cuda:0")
    fmt = torch.channels_last_3d if nd == 3 else torch.channels_last
    shape = (n, cin, d, h, w) if d else (n, cin, h, w)
    xs = [torch.randn(shape, device=dev).contiguous(memory_format=fmt).requires_grad_(dgrad) for _ in range(bufs)]
    wshape = ((cin, cout) if kind == "up2" else (cout, cin)) + (k,) * nd
    weight = torch.nn.Parameter(torch.randn(wshape, device=dev) * 0.05)
    bias = torch.nn.Parameter(torch.zeros(cout, device=dev))
    y, _ = ops.conv_stats(xs[0], weight, bias, KINDS[kind], True)
    gs = [torch.randn_like(y) for _ in range(bufs)]
    y.backward(gs[0])                       # warm-up: the first launch of a kernel pays its module load
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    for i in range(reps):
        x = xs[i % bufs]
        x.grad = None
        y, _ = ops.conv_stats(x, weight, bias, KINDS[kind], True)
        y.backward(gs[i % bufs])
    rep = _lib.timing_report()
    _lib.timing_enable(False)
    out = []
    for name, v in sorted(rep.items()):
        out.append("%s %.1fus" % (name.split(":")[0], 1e3 * v["ms"] / max(v["launches"], 1)))
    in_mb = 4e-6 * n * cin * h * w * (d or 1)
    out_mb = in_mb / cin * cout * (4 if kind == "up2" else 0.25 if kind == "down2" else 1)
    print("%-28s in %.0f MB out %.0f MB | %s" % (spec, in_mb, out_mb, " | ".join(out)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("layers", nargs="+")
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--bufs", type=int, default=3)
    ap.add_argument("--no-dgrad", action="store_true", help="the input needs no gradient (network stems)")
    ap.add_argument("--env", action="append", default=[], help="NAME=VALUE set before the run (repeatable)")
    args = ap.parse_args()
    for kv in args.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    for spec in args.layers:
        run_layer(spec, args.reps, args.bufs, not args.no_dgrad)


if __name__ == "__main__":
    main()

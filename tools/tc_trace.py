"""Timeline of CTA 0 of conv_tc_kernel (developer tool; needs the -DCHAP_TC_DEBUG_HOOKS build of the library:
CHAP_B200_LIB=chap_b200/lib/libchap_b200_dbg.so python tools/tc_trace.py "k3 16 16 12 256 256" ...)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from chap_b200 import _lib, ops  # noqa: E402

NAMES = ["entry", "prologue done", "first stage full", "first accumulator full", "first tile stored", "last tile stored",
         "statistics written", "before exit"]
for spec in sys.argv[1:]:
    f = spec.split()
    kind, cin, cout, n, h, w = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5])
    k = {"k3": 3, "k1": 1}[kind]
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    kc = {"k3": _lib.CONV_K3, "k1": _lib.CONV_K1}[kind]
    for _ in range(3):
        ops.conv_stats(x, wt, b, kc, True)
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 8)()
    assert _lib.load().chap_debug_tc_trace(out) == 0
    t = list(out)
    mhz = 1965.0
    print(spec, " | ".join("%s +%.2fus" % (NAMES[i], (t[i] - t[0]) / mhz) for i in range(1, 8)))

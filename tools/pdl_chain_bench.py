"""Per-launch cost of a dependent chain of small kernels, eager and as a replayed CUDA graph, with programmatic dependent launch on / off.
Usage: python tools/pdl_chain_bench.py            (prints us per launch for a few tensor sizes)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chap_b200 import ops


def chain(x, n):
    for _ in range(n):
        x = ops.axpy(x, x, 0.5)
    return x


def run(elems, n=200, reps=20):
    dev = torch.device("cuda:0")
    a = torch.randn(1, 16, 64, max(1, elems // 1024), device=dev).contiguous(memory_format=torch.channels_last)
    out = {}
    for pdl in (1, 0):
        ops.set_pdl(pdl)
        for mode in ("eager", "graph"):
            if mode == "graph":
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    chain(a, 3)
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, stream=s):
                        y = chain(a, n)
                fn = g.replay
            else:
                fn = lambda: chain(a, n)
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            out[(pdl, mode)] = e0.elapsed_time(e1) * 1e3 / (reps * n)
    print(f"{elems * 4 / 1e6:8.2f} MB: " + "  ".join(f"pdl={k[0]} {k[1]} {v:6.2f} us" for k, v in out.items()))


if __name__ == "__main__":
    for e in (1 << 12, 1 << 18, 1 << 22, 1 << 24):
        run(e)

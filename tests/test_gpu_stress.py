"""GPU: repeated-run checks of the kernels that synchronise by hand -- the lock-free union-find of the largest-CC filter, the
ticket-counter BatchNorm finalize inside the convolution kernel, fp32 / vector `red.global.add` weight gradients, and the
mbarrier / tcgen05 pipelines across persistent tiles.  compute-sanitizer is closed on this GPU pool ("runs under it have left GPUs
needing a reset", profiles/r02_sanitizer.md), so races are hunted the other way round: the same launch many times, every result
compared with the CPU oracle or with an independently computed value; a race shows up as a run that differs."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import chap_losses as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_largest_cc_union_find_is_deterministic_under_contention():
    from chap_b200 import ops
    g = torch.Generator().manual_seed(0)
    noise = torch.rand((6, 256, 256), generator=g)
    seg = (noise * 4 * 1.3).long().clamp_(0, 3)                 # speckle: thousands of small components, heavy root contention
    seg[1, 64:192, 64:192] = 2                                  # one huge component
    seg[2, ::2, :] = 1                                          # long horizontal runs
    seg[3, :, ::2] = 3                                          # long vertical runs (unions only at run boundaries)
    want = L.largest_cc_labels(seg, 4)
    dseg = seg.to(DEV)
    for rep in range(40):
        got = ops.largest_cc(dseg, 4).cpu()
        assert torch.equal(got, want), rep


@pytest.mark.parametrize("shape", [(2, 16, 64, 64, 32), (3, 64, 24, 40, 64), (2, 16, 12, 28, 20, 16)])
def test_conv_bn_ticket_finalize_matches_two_pass_statistics_every_time(shape):
    """chap_conv_bn_fwd: statistics accumulated by all CTAs with double atomics, finalize by whichever CTA draws the last ticket.
    200 launches; mean / invstd must equal statistics recomputed from the stored output by a separate pass, every time."""
    from chap_b200 import ops
    from chap_b200._lib import CONV_K3
    import torch.nn as nn
    n, cin = shape[0], shape[1]
    sp, cout = shape[2:-1], shape[-1]
    torch.manual_seed(1)
    x = ops.cl(torch.randn((n, cin) + tuple(sp), device=DEV))
    nd = len(sp)
    conv = (nn.Conv2d if nd == 2 else nn.Conv3d)(cin, cout, 3, padding=1).to(DEV)
    bn = (nn.BatchNorm2d if nd == 2 else nn.BatchNorm3d)(cout).to(DEV).train()
    dims = (0,) + tuple(range(2, 2 + nd))
    for rep in range(200):
        with torch.no_grad(), ops.bn_tracking(False):
            y, st = ops.conv_stats(x, conv.weight, conv.bias, CONV_K3, True, feeds_train_bn=True, bn=bn)
        assert isinstance(st, ops.BnStats)
        mi = st[1]
        mean, var = y.double().mean(dims), y.double().var(dims, unbiased=False)
        assert rel_err(mi[:cout], mean) < 1e-4 or float((mi[:cout].double() - mean).abs().max()) < 1e-5, rep
        assert rel_err(mi[cout:], 1.0 / torch.sqrt(var + bn.eps)) < 1e-5, rep


def test_weight_gradient_atomics_reproduce_across_runs():
    """The tensor-core weight gradient reduces pixel splits with fp32 atomics: 30 runs of the same layer must agree with an
    fp64 evaluation within TF32 rounding and with each other within fp32 summation-order noise."""
    from chap_b200 import ops
    from chap_b200._lib import CONV_K3
    torch.manual_seed(2)
    x = ops.cl(torch.randn(4, 32, 96, 96, device=DEV)).requires_grad_(True)
    w = (torch.randn(32, 32, 3, 3, device=DEV) / 17.0).requires_grad_(True)
    b = torch.zeros(32, device=DEV, requires_grad=True)
    dy = ops.cl(torch.randn(4, 32, 96, 96, device=DEV))
    ref = torch.nn.grad.conv2d_weight(x.detach().double(), w.shape, dy.double(), padding=1)
    first = None
    for rep in range(30):
        y = ops.conv(x, w, b, CONV_K3)
        (gw,) = torch.autograd.grad(y, (w,), dy)
        assert rel_err(gw, ref) < 2e-3, rep
        if first is None:
            first = gw.clone()
        assert rel_err(gw, first) < 1e-5, rep


def test_graph_replays_of_the_whole_iteration_are_reproducible():
    """Two trainers, same seeds, no stochastic inputs: 8 graph replays each; losses agree step by step (a pipeline race in the
    persistent conv kernels or a stale statistics slot would make runs drift apart immediately)."""
    from conftest import seeded_model
    from chap_b200.train_step import ChapTrainer
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    w = dict(bench.WORKLOADS["unet2d"], shape=(128, 128), batch=8, labeled=4)
    runs = []
    for _ in range(2):
        m = seeded_model("dualdecoder2d", seed=8).to(DEV)
        t = ChapTrainer(m, 4, 4, max_iterations=100, adv_noise=False, use_graph=True, graph_warmup=1)
        seen = []
        for it in range(8):
            vol, lab = bench.synth_batch(w, it)
            seen.append(float(t.step(vol.to(DEV), lab.to(DEV), mask_offsets=(5, 9))["loss"]))
        t.close()
        runs.append(seen)
    assert np.allclose(runs[0], runs[1], rtol=2e-3), runs        # atomics reorder fp32 sums; nothing more
    assert runs[0][-1] < runs[0][0]

"""Performance guard (B200): the CUDA path must beat the SAME iteration executed by torch-eager + cuDNN on the same GPU.

SURVEY.md 8(d): "time PyTorch-eager on the B200 running the same oracle -- that, not the CPU, is the bar the kernels must
beat".  The oracle restatement (oracle/train_step.py: reference nets + frozen losses, host largest-CC like the reference)
is moved to the device and timed as the checker; the product path is ChapTrainer (CUDA graph).  Prints both numbers."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _synthetic(n, size, classes, seed):
    g = torch.Generator().manual_seed(seed)
    vol = torch.rand((n, 1, size, size), generator=g)
    lab = torch.zeros((n, size, size), dtype=torch.int64)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    for i in range(n):
        for c in range(1, classes):
            cy, cx = [int(torch.randint(size // 4, 3 * size // 4, (1,), generator=g)) for _ in range(2)]
            r = size // (6 + 2 * c)
            lab[i][(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = c
    return vol, lab


def test_chap_iteration_beats_torch_eager_cudnn_on_the_same_gpu():
    from chap_b200 import networks
    from chap_b200.train_step import ChapTrainer
    from oracle import chap_losses as L
    from oracle import nets
    from oracle import train_step as ost
    n, labeled, size, classes = 24, 12, 256, 4
    torch.manual_seed(1337)
    model = networks.DualDecoder(1, classes, {"decoder_type": "mcnet"})
    # ---- checker: the oracle restatement as torch-eager / cuDNN on the device (cuDNN TF32 allowed: torch's default)
    sd = nets.clone_state_dict(model.state_dict(), requires_grad=True, device=DEV)
    om = ost.OracleModel(sd, dims=2, has_dropout=False, drop="torch")
    bufs = [None] * len(om.params())
    vat = L.VAT(10.0, 6.0, classes)
    batches = [_synthetic(n, size, classes, s) for s in range(3)]
    dev_batches = [(v.to(DEV), l.to(DEV)) for v, l in batches]
    eager = []
    for it in range(4):
        v, l = dev_batches[it % 3]
        offs = L.draw_mask_offsets((size, size), np.random.RandomState(it))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ost.chap_train_step(om, bufs, v, l, labeled, classes, offs, it, vat=vat, topk=0.1)
        torch.cuda.synchronize()
        if it >= 1:
            eager.append(time.perf_counter() - t0)
    eager_ms = 1e3 * float(np.median(eager))
    # ---- product path
    trainer = ChapTrainer(model.to(DEV), n_classes=classes, labeled_bs=labeled, use_graph=True)
    for it in range(5):
        trainer.step(*dev_batches[it % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 10
    for it in range(steps):
        out = trainer.step(*dev_batches[it % 3])
    e1.record()
    torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / steps
    print("\n2D CHAP iteration b24 256^2 on this GPU: torch-eager/cuDNN oracle %.1f ms, chap_b200 %.2f ms (%.1fx)"
          % (eager_ms, ours_ms, eager_ms / ours_ms))
    assert torch.isfinite(out["loss"]).item()
    assert ours_ms < eager_ms, "the CUDA path is slower than torch-eager on the same GPU"

"""CPU, build container only: the oracle restatement and the product modules' parameter layout are
checked against the UNMODIFIED reference imported from /root/reference (skipped where it is absent,
e.g. on the GPU box -- the committed golden fixtures cover that case)."""
import numpy as np
import pytest
import torch

from conftest import seeded_model, zero_dropout
from oracle import nets, ref_import, sliding_window

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


def _same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), k


def test_state_dict_layout_and_seeded_init_match_reference(ref):
    torch.manual_seed(1337)
    _same_state(ref.DualDecoder(1, 4, {"decoder_type": "mcnet"}), seeded_model("dualdecoder2d"))
    torch.manual_seed(1337)
    _same_state(ref.UNet(1, 4), seeded_model("unet2d"))
    torch.manual_seed(1337)
    _same_state(ref.DualDecoder3d(1, 2, normalization='batchnorm', has_dropout=False), seeded_model("dualdecoder3d"))
    torch.manual_seed(1337)
    _same_state(ref.VNet(1, 2, normalization='batchnorm', has_dropout=False), seeded_model("vnet"))
    assert len(seeded_model("dualdecoder2d").state_dict()) == 202 and len(seeded_model("dualdecoder3d").state_dict()) == 298


def test_reference_checkpoint_loads_into_product_modules(ref):
    torch.manual_seed(7)
    r = ref.DualDecoder(1, 4, {"decoder_type": "mcnet"})
    m = seeded_model("dualdecoder2d", seed=99)
    missing, unexpected = m.load_state_dict(r.state_dict(), strict=True)
    assert not missing and not unexpected
    _same_state(r, m)


def test_oracle_forward_backward_bit_exact_2d(ref):
    torch.manual_seed(3)
    r = ref.DualDecoder(1, 4, {"decoder_type": "mcnet"})
    zero_dropout(r)
    r.train()
    x = torch.rand(3, 1, 32, 32)
    sd = nets.clone_state_dict(r.state_dict(), requires_grad=True)
    o1, o2 = r(x)
    p1, p2 = nets.dualdecoder2d_forward(sd, x, True, True, None)
    assert torch.equal(o1, p1) and torch.equal(o2, p2)
    gr = torch.autograd.grad((o1 ** 2).sum() + o2.sum(), list(r.parameters()))
    names = [n for n, _ in r.named_parameters()]
    go = torch.autograd.grad((p1 ** 2).sum() + p2.sum(), [sd[n] for n in names])
    for n, a, b in zip(names, gr, go):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), n


def test_oracle_forward_bit_exact_3d_with_dropout_masks(ref):
    torch.manual_seed(5)
    r = ref.DualDecoder3d(1, 2, normalization='batchnorm', has_dropout=True)
    r.train()
    x = torch.randn(2, 1, 16, 16, 16)
    sd = nets.clone_state_dict(r.state_dict())
    torch.manual_seed(11)
    o1, o2 = r(x)                                   # draws 3 Dropout3d masks: encoder x5, decoder1 x9, decoder2 x9
    torch.manual_seed(11)
    p1, p2 = nets.dualdecoder3d_forward(sd, x, True, True, True, "torch")
    assert torch.equal(o1, p1) and torch.equal(o2, p2)


def _dropout_case(branch):
    """(features, level, scores, comp_drop) for one branch of perform_dropout (code/networks/FilterDropout.py:54-80)."""
    g = torch.Generator().manual_seed(31)
    feats = [torch.randn(4, c, s, s, generator=g) for c, s in ((16, 8), (32, 4), (64, 2))]
    level = [0, 2]                                                   # level 1 stays untouched (:82-84)
    if branch == "binomial_comp":
        return feats, level, None, True
    if branch == "dropout2d":
        return feats, level, None, False
    if branch == "zero_scores":
        return feats, level, [torch.zeros(c) for c in (16, 32, 64)], False
    scores = [torch.rand(c, generator=g) + 0.1 for c in (16, 32, 64)]
    return feats, level, scores, branch == "scores_comp"


@pytest.mark.parametrize("branch", ["none", "binomial_comp", "dropout2d", "zero_scores", "scores", "scores_comp"])
def test_perform_dropout_every_branch_matches_reference(ref, branch):
    """Under the same seeds the product's mask draws (draw_dropout_masks: same generator calls in the same order) equal the
    factors the reference applied (recovered from its outputs), and the oracle's apply step given those factors reproduces the
    reference's outputs bit for bit.  (The product's apply step is a CUDA kernel: tests/test_gpu_ops.py against the fixture.)"""
    import random
    from chap_b200.networks import FilterDropout as fd
    from oracle import filter_dropout as ofd
    if branch == "none":
        feats, level, scores, comp = _dropout_case("dropout2d")
        level = []
    else:
        feats, level, scores, comp = _dropout_case(branch)
    torch.manual_seed(77); random.seed(5)
    r1, r2 = ref.perform_dropout(feats, level=level, scores=scores, comp_drop=comp)
    torch.manual_seed(77); random.seed(5)
    masks = fd.draw_dropout_masks(feats, level, scores, comp)
    rec1, rec2 = ofd.recover_masks(feats, r1, level), ofd.recover_masks(feats, r2, level)
    for idx, m in enumerate(masks):
        if idx not in level:
            assert m is None
            continue
        assert torch.allclose(m[0].reshape(rec1[idx].shape), rec1[idx], rtol=1e-6, atol=0), (branch, idx)
        assert torch.allclose(m[1].reshape(rec2[idx].shape), rec2[idx], rtol=1e-6, atol=0), (branch, idx)
    o1, o2 = ofd.perform_dropout(feats, masks)
    for u, v in zip(r1 + r2, o1 + o2):
        assert torch.equal(u, v) and u.shape[0] == 6


def test_sliding_window_oracle_equals_reference_function(ref):
    r_sw = ref_import.load_sliding_window()
    torch.manual_seed(2)
    conv = torch.nn.Conv3d(1, 3, 3, padding=1)
    net_t = lambda p: conv(p)                                    # noqa: E731
    net_np = lambda p: conv(torch.from_numpy(p)).detach().numpy()   # noqa: E731
    image = np.random.RandomState(0).randn(37, 30, 21).astype(np.float32)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        want = r_sw.test_single_case(net_t, image, 7, 5, (24, 24, 16), num_classes=3)
    finally:
        torch.Tensor.cuda = orig
    got = sliding_window.test_single_case(net_np, image, 7, 5, (24, 24, 16), 3,
                                          softmax_fn=lambda y: torch.softmax(torch.from_numpy(y), 1).numpy())
    assert np.array_equal(want, got)

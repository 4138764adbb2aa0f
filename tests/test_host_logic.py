"""CPU: host-side logic of the product package that needs no GPU (schedules, masks, metrics, window geometry) and
its agreement with the oracle restatement."""
import numpy as np
import torch

from oracle import chap_losses as L
from oracle import sliding_window as osw


def test_schedules_match_oracle():
    from chap_b200.train_step import consistency_weight
    from chap_b200.utils import ramps
    for t in (0, 1, 149, 150, 3000, 7500, 29999):
        assert consistency_weight(t, 1.0, 50.0) == L.consistency_weight(t, 1.0, 50.0)
    assert ramps.sigmoid_rampup(0, 50.0) == float(np.exp(-5.0)) and ramps.sigmoid_rampup(50, 50.0) == 1.0
    assert ramps.sigmoid_rampup(3, 0) == 1.0


def test_generate_mask_matches_reference_semantics():
    from chap_b200.train_step import generate_mask
    img = torch.zeros(6, 1, 256, 256)
    mask, loss_mask = generate_mask(img, (10, 20))
    assert mask.dtype == torch.int64 and loss_mask.shape == (6, 256, 256)
    assert int((mask == 0).sum()) == 170 * 170 and mask[10, 20] == 0 and mask[9, 20] == 1 and mask[179, 189] == 0 and mask[180, 189] == 1
    assert torch.equal(mask, L.generate_mask((256, 256), (10, 20)))
    m3, _ = generate_mask(torch.zeros(1, 1, 112, 112, 80), (1, 2, 3))
    assert int((m3 == 0).sum()) == 74 * 74 * 53
    np.random.seed(0)
    a, _ = generate_mask(img)
    np.random.seed(0)
    assert L.draw_mask_offsets((256, 256)) == tuple(int(torch.nonzero(a == 0)[0][i]) for i in range(2))


def test_host_largest_cc_and_metrics():
    from chap_b200.test_3D_util import asd, cal_dice, dice_coefficient, hd95, jc
    seg = torch.zeros(1, 8, 8, dtype=torch.int64)
    seg[0, 0:2, 0:2] = 1; seg[0, 4:7, 4:7] = 1; seg[0, 0, 7] = 2
    out = L.largest_cc_labels(seg, 3)               # the oracle's filter (the product's is the device kernel: tests/test_gpu_step.py)
    assert out.sum() == 9 + 2 and out[0, 0, 0] == 0 and out[0, 5, 5] == 1 and out[0, 0, 7] == 2
    a = np.zeros((8, 8, 8), bool); a[2:6, 2:6, 2:6] = True
    b = np.zeros((8, 8, 8), bool); b[3:7, 2:6, 2:6] = True
    assert abs(dice_coefficient(a, b) - 0.75) < 1e-12 and abs(jc(a, b) - 0.6) < 1e-12
    assert dice_coefficient(a, a) == 1.0 and dice_coefficient(np.zeros(3), np.zeros(3)) == 0.0
    assert hd95(a, a) == 0.0 and hd95(a, b) == 1.0 and 0.0 < asd(a, b) < 1.0
    assert abs(cal_dice(a.astype(int), b.astype(int), 2)[0] - 0.75) < 1e-12
    assert L.dice_coefficient(a, b) == dice_coefficient(a, b)


def test_window_geometry_matches_oracle():
    from chap_b200.test_3D_util import _pad_amounts
    for shape in ((192, 192, 88), (177, 203, 88), (100, 120, 70)):
        assert _pad_amounts(shape, (112, 112, 80)) == osw.pad_amounts(shape, (112, 112, 80))
    assert len(osw.window_starts(192, 112, 18)) ** 2 * len(osw.window_starts(88, 80, 4)) == 108


def test_filter_dropout_mask_helpers():
    from chap_b200.networks import FilterDropout as fd
    torch.manual_seed(0)
    probs = torch.rand(6, 32)
    m1, m2 = fd.drop_based_on_prob(probs, False)
    assert m1.shape == (6, 32) and abs(float(m1.mean()) - 1.0) < 1e-5 and abs(float(m2.mean()) - 1.0) < 1e-5
    m1, m2 = fd.scores_dropoutV2(torch.rand(32), torch.rand(6, 32), True, 'sigmoid')
    assert set(torch.unique(m1 > 0).tolist()) <= {True, False} and m1.shape == (6, 32)


def test_zoom_index_tables_reproduce_scipy_zoom_order0():
    """ops.zoom_index is scipy.ndimage.zoom(order=0)'s index map (code/val_2D.py:60,91), including its constant-fill quirk on an
    overshooting last coordinate; the gather kernel only applies these tables."""
    import numpy as np
    from scipy.ndimage import zoom
    from chap_b200.ops import zoom_index
    rng = np.random.RandomState(0)
    for (h, w, hh, ww) in [(200, 180, 256, 256), (256, 256, 200, 180), (313, 257, 256, 256), (256, 256, 313, 257), (64, 64, 256, 256),
                           (5, 7, 256, 256), (256, 256, 5, 7), (256, 256, 256, 256), (224, 208, 256, 256), (256, 256, 224, 208)]:
        a = rng.rand(h, w).astype(np.float32) + 1.0
        ref = zoom(a, (hh / h, ww / w), order=0)
        oh, ow = int(round(h * (hh / h))), int(round(w * (ww / w)))
        assert ref.shape == (oh, ow)
        iy, ix = zoom_index(h, oh), zoom_index(w, ow)
        got = np.where((iy[:, None] < 0) | (ix[None, :] < 0), np.float32(0), a[np.maximum(iy, 0)][:, np.maximum(ix, 0)])
        assert np.array_equal(ref, got), (h, w, hh, ww)


def test_generated_dropout_description_is_host_side_bookkeeping_only():
    """ops.dropout_rng(p, C): the (p, seed, subsequence, epoch) description handed to chap_bn_act_{fwd,bwd}_rng -- None whenever the
    generated form does not apply (inactive dropout, channel counts outside the fixed-group kernels, switched off), a fresh
    subsequence per call, the torch seed as key, the registered device epoch tensor passed through."""
    from chap_b200 import ops
    torch.manual_seed(1234)
    assert ops.dropout_rng(0.0, 16) is None and ops.dropout_rng(1.0, 16) is None
    assert ops.dropout_rng(0.3, 6) is None and ops.dropout_rng(0.3, 40) is None            # c % 4 != 0 / 256 % (c / 4) != 0
    a, b = ops.dropout_rng(0.05, 16), ops.dropout_rng(0.5, 256)
    assert a[0] == 0.05 and b[0] == 0.5 and a[1] == b[1] == 1234 and b[2] == a[2] + 1 and a[3] is None
    marker = torch.zeros(1, dtype=torch.int64)
    ops.set_dropout_epoch(marker)
    try:
        assert ops.dropout_rng(0.1, 32)[3] is marker
    finally:
        ops.set_dropout_epoch(None)
    ops.set_generated_dropout(False)
    try:
        assert ops.dropout_rng(0.1, 32) is None
    finally:
        ops.set_generated_dropout(True)

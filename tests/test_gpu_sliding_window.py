"""GPU: sliding-window aggregation (chap_sw_extract / chap_sw_aggregate / test_single_case) against
the numpy oracle and the reference-generated fixtures.  Label and count maps must be bit-exact;
score maps are bit-exact when both sides are given the same per-window probabilities."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import sliding_window as osw

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class TinyNet(torch.nn.Module):
    """conv3^3 + tanh on torch ops: a stand-in network so that the test isolates the aggregator."""

    def __init__(self, w, b):
        super().__init__()
        self.conv = torch.nn.Conv3d(1, w.shape[0], 3, padding=1)
        with torch.no_grad():
            self.conv.weight.copy_(torch.as_tensor(w))
            self.conv.bias.copy_(torch.as_tensor(b))

    def forward(self, x):
        return torch.tanh(self.conv(x)) * 3.0


@pytest.mark.parametrize("name", ["clamped", "padded", "exact"])
def test_aggregator_bit_exact_given_window_probabilities(name):
    from chap_b200 import ops
    from chap_b200.test_3D_util import test_single_case
    g = golden("sliding_window.npz")
    cfg = g[name + "_cfg"]
    patch, sxy, sz = tuple(int(v) for v in cfg[:3]), int(cfg[3]), int(cfg[4])
    image = g[name + "_image"]
    net = TinyNet(g["conv_w"], g["conv_b"]).to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    captured = {}
    label, score, cnt = test_single_case(net, image, sxy, sz, patch, 2, batch_windows=3, return_maps=True,
                                         window_logits_hook=lambda first, lg: captured.__setitem__(first, lg.clone()))
    # replay the SAME per-window probabilities (the kernel's own softmax, copied to the host) through the numpy oracle
    logits = torch.cat([captured[k] for k in sorted(captured)]).contiguous(memory_format=torch.channels_last_3d)
    probs = ops.softmax(logits).cpu().numpy()
    it = iter(range(probs.shape[0]))
    want_label, want_score, want_cnt = osw.test_single_case(lambda p: probs[next(it)][None], image, sxy, sz, patch, 2,
                                                            softmax_fn=lambda y: y, return_maps=True)
    assert label.dtype == np.int64 and label.shape == image.shape
    assert np.array_equal(cnt, want_cnt)
    assert np.array_equal(score, want_score)            # same summation order -> bit-identical fp32
    assert np.array_equal(label, want_label)
    # against the reference-generated label map: differences only where the two class scores are within float noise
    diff = label != g[name + "_label"]
    assert diff.mean() < 5e-3
    if diff.any():
        assert np.abs(score[0][diff] - score[1][diff]).max() < 1e-4


def test_extract_matches_numpy_slicing():
    from chap_b200 import ops
    vol = np.random.RandomState(0).randn(20, 18, 14).astype(np.float32)
    desc = ops.sw_desc(vol.shape, (8, 8, 6), (3, 3, 3), (6, 6, 4), 2)
    got = ops.sw_extract(desc, torch.from_numpy(vol).to(DEV), 0, 27).cpu().numpy()
    i = 0
    for xs in osw.window_starts(20, 8, 6):
        for ys in osw.window_starts(18, 8, 6):
            for zs in osw.window_starts(14, 6, 4):
                assert np.array_equal(got[i, 0], vol[xs:xs + 8, ys:ys + 8, zs:zs + 6]); i += 1
    assert i == 27


def test_vnet_sliding_window_end_to_end_vs_oracle():
    """product VNet through the product test_single_case vs the functional oracle net + numpy oracle."""
    from conftest import seeded_model
    from chap_b200.test_3D_util import test_single_case
    from oracle import nets
    m = seeded_model("vnet", seed=6).to(DEV).eval()
    sd = nets.clone_state_dict(m.state_dict())
    image = np.random.RandomState(3).randn(40, 36, 20).astype(np.float32)

    def oracle_net(p):
        with torch.no_grad():
            return nets.vnet_forward(sd, torch.from_numpy(p), False).numpy()
    want, wscore, wcnt = osw.test_single_case(oracle_net, image, 8, 4, (32, 32, 16), 2,
                                              softmax_fn=lambda y: torch.softmax(torch.from_numpy(y), 1).numpy(), return_maps=True)
    got, score, cnt = test_single_case(m, image, 8, 4, (32, 32, 16), 2, batch_windows=2, return_maps=True)
    assert np.array_equal(cnt, wcnt)
    np.testing.assert_allclose(score, wscore, atol=2e-3)
    diff = got != want
    assert diff.mean() < 0.02
    if diff.any():
        assert np.abs(wscore[0][diff] - wscore[1][diff]).max() < 5e-3     # ties only


def test_val_2d_test_single_volume_batched_equals_slice_loop():
    """chap_b200.val_2D.test_single_volume (all slices in one forward) vs the reference's per-slice protocol
    (code/val_2D.py:57-92) run with the same product net: identical label volume, Dice interface preserved."""
    from scipy.ndimage import zoom
    from conftest import seeded_model
    from chap_b200 import ops
    from chap_b200.val_2D import predict_volume, test_single_volume
    m = seeded_model("dualdecoder2d", seed=9).to(DEV).eval()
    rng = np.random.RandomState(1)
    image = rng.rand(5, 40, 36).astype(np.float32)
    label = (rng.rand(5, 40, 36) * 4).astype(np.int64)
    got = predict_volume(image, m, (64, 64), 'logit_ensemble', DEV)
    want = np.zeros_like(label)
    with torch.no_grad():
        for i in range(5):
            sl = zoom(image[i], (64 / 40, 64 / 36), order=0)
            o1, o2 = m(torch.from_numpy(sl)[None, None].float().to(DEV))
            out = ops.argmax(o1, o2)[0].cpu().numpy()
            want[i] = zoom(out, (40 / 64, 36 / 64), order=0)
    assert (got != want).mean() < 1e-3          # batch-1 vs batch-5 convolutions may differ in the last bit of a tie
    res = test_single_volume(torch.from_numpy(image)[None], torch.from_numpy(label)[None], m, classes=4,
                             patch_size=[64, 64], model_type='logit_ensemble', device=DEV)
    assert len(res) == 3 and all(len(r) == 2 for r in res)

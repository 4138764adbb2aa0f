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


def test_val_2d_device_zoom_and_dice_equal_host_protocol():
    """On-device nearest zoom (incl. scipy's constant-fill quirk on 256 -> 200) and on-device Dice counts against the host
    protocol of code/val_2D.py:43-51,57-92; `validate` shards volumes over ranks and the partial means add up."""
    from scipy.ndimage import zoom
    from conftest import seeded_model
    from chap_b200 import ops
    from chap_b200 import val_2D
    rng = np.random.RandomState(3)
    for shape, out in (((3, 200, 180), (256, 256)), ((2, 256, 256), (200, 180)), ((4, 64, 72), (64, 64))):
        a = rng.rand(*shape).astype(np.float32) + 1.0
        lab = (rng.rand(*shape) * 5).astype(np.int64)
        want_f = np.stack([zoom(a[i], (out[0] / shape[1], out[1] / shape[2]), order=0) for i in range(shape[0])])
        want_l = np.stack([zoom(lab[i], (out[0] / shape[1], out[1] / shape[2]), order=0) for i in range(shape[0])])
        assert np.array_equal(ops.zoom_nearest(torch.from_numpy(a).to(DEV), *want_f.shape[1:]).cpu().numpy(), want_f)
        assert np.array_equal(ops.zoom_nearest(torch.from_numpy(lab).to(DEV), *want_l.shape[1:]).cpu().numpy(), want_l)
    pred = (rng.rand(6, 50, 40) * 4).astype(np.int64)
    gt = (rng.rand(6, 50, 40) * 4).astype(np.int64)
    pred[pred == 3] = 0                                           # class 3 never predicted -> (0, 0) like :50-51
    counts = ops.label_overlap(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV), 4).cpu().numpy()
    for c in range(4):
        assert tuple(counts[c]) == (int(((pred == c) & (gt == c)).sum()), int((pred == c).sum()), int((gt == c).sum()))
    m = seeded_model("dualdecoder2d", seed=9).to(DEV).eval()
    vols = [dict(image=torch.from_numpy(rng.rand(1, 4, 48, 40).astype(np.float32)),
                 label=torch.from_numpy((rng.rand(1, 4, 48, 40) * 4).astype(np.int64))) for _ in range(3)]
    full = val_2D.validate(vols, m, 4, [64, 64], 'logit_ensemble', DEV)
    for b in vols:                                                # Dice / HD95 of every volume equal the host metric on the same prediction
        p = val_2D.predict_volume(b["image"][0].numpy(), m, (64, 64), 'logit_ensemble', DEV)
        res = val_2D.test_single_volume(b["image"], b["label"], m, 4, [64, 64], 'logit_ensemble', DEV)
        for i in range(1, 4):
            d, h = val_2D.calculate_metric_percase(p == i, b["label"][0].numpy() == i)
            assert abs(res[i - 1][0] - d) < 1e-12 and abs(res[i - 1][1] - h) < 1e-9
    parts = [val_2D.validate(vols, m, 4, [64, 64], 'logit_ensemble', DEV, rank=r, world_size=2) for r in range(2)]
    np.testing.assert_allclose(parts[0] + parts[1], full, rtol=1e-12)
    assert full.shape == (3, 2)


def test_test_all_case_reference_signature_reads_case_files(tmp_path):
    """test_all_case(net, base_dir, method, test_list, ...) / val_3D.test_all_case(net, base_dir, test_list, ...) with the
    reference's positional signatures (code/test_3D_util.py:91, code/val_3D.py:91): case list + per-case files on disk, metric
    text file, and the in-memory / sharded extensions give the same numbers."""
    from conftest import seeded_model
    from chap_b200 import test_3D_util, val_3D
    rng = np.random.RandomState(5)
    base = tmp_path
    (base / "data").mkdir()
    cases = []
    for i in range(3):
        img = rng.randn(40, 36, 24).astype(np.float32)
        lab = (rng.rand(40, 36, 24) > 0.5).astype(np.uint8)
        np.savez(str(base / "data" / ("case%d.npz" % i)), image=img, label=lab)
        cases.append((img, lab))
    (base / "test.list").write_text("case0\ncase1,extra\ncase2\n")
    net = seeded_model("vnet", seed=3).to(DEV).eval()
    out_dir = tmp_path / "pred"
    out_dir.mkdir()
    got = test_3D_util.test_all_case(net, str(base), "vnet", "test.list", 2, (32, 32, 16), 16, 8, str(out_dir))
    assert got.shape == (1, 4)
    mem = test_3D_util.test_all_case(net, None, num_classes=2, patch_size=(32, 32, 16), stride_xy=16, stride_z=8, cases=cases)
    np.testing.assert_allclose(got, mem, rtol=1e-12)
    parts = [test_3D_util.test_all_case(net, None, num_classes=2, patch_size=(32, 32, 16), stride_xy=16, stride_z=8, cases=cases,
                                        rank=r, world_size=2) for r in range(2)]
    np.testing.assert_allclose(parts[0] + parts[1], got, rtol=1e-12)
    text = (out_dir / "vnet.txt").read_text()
    assert text.count("\n") == 3 and "Mean metrics" in text and (out_dir / "case1_pred.npy").exists()
    v = val_3D.test_all_case(net, str(base), "test.list", 2, (32, 32, 16), 16, 8)
    assert v.shape == (1, 2) and abs(v[0, 0] - got[0, 0]) < 1e-12          # same Dice through cal_metric
    from chap_b200 import networks
    net1 = networks.VNet(1, 1, normalization='batchnorm', has_dropout=False).to(DEV).eval()
    lab1 = test_3D_util.test_single_case(net1, cases[0][0], 16, 8, (32, 32, 16))    # the reference's default num_classes=1
    assert lab1.shape == cases[0][0].shape and not lab1.any()                        # softmax over one class: label 0 everywhere


def test_eval_mode_fused_conv_bn_act_equals_separate_kernels_and_oracle():
    """Inference path: conv + eval BatchNorm + (Leaky)ReLU + additive skip in ONE kernel (chap_conv_bn_act_fwd) vs the same
    network run with the separate BatchNorm pass (autograd enabled -> not fused) and vs the oracle; running statistics that
    change (a train-mode pass) invalidate the cached scale / shift."""
    from conftest import rel_err, seeded_model
    from chap_b200 import _lib
    from oracle import nets
    # (W = 48 in 3D: the top level runs the kx-in-N tile -- 30 output columns per tile, partial second tile -- with the inference epilogue)
    for kind, shape in (("vnet", (2, 1, 16, 32, 48)), ("dualdecoder2d", (3, 1, 64, 64))):
        m = seeded_model(kind, seed=5).to(DEV)
        x = torch.randn(*shape, device=DEV)
        m.train()
        with torch.no_grad():
            m(x)                                             # moves the running statistics away from (0, 1)
        m.eval()
        sd = nets.clone_state_dict(m.state_dict())
        _lib.timing_enable(True)
        with torch.no_grad():
            fused = m(x)
        rep = _lib.timing_report()
        _lib.timing_enable(False)
        n_bn = sum(v["launches"] for k, v in rep.items() if k.split(":")[0] == "bn_act_fwd")
        assert n_bn == 1, rep.keys()                         # only the Cin = 1 stem (CUDA-core conv) keeps its separate BatchNorm pass
        with torch.enable_grad():
            plain = m(x)                                     # autograd on: conv, BatchNorm pass, activation as separate kernels
        want = nets.vnet_forward(sd, x.cpu(), False, False) if kind == "vnet" else nets.dualdecoder2d_forward(sd, x.cpu(), False, False)
        for a, b, c in zip(fused if isinstance(fused, tuple) else (fused,), plain if isinstance(plain, tuple) else (plain,),
                           want if isinstance(want, tuple) else (want,)):
            assert rel_err(a, b) < 1e-5, kind                # same arithmetic, fused
            assert rel_err(a, c) < 5e-3, kind                # TF32 convs vs the fp32 oracle
        m.train()
        with torch.no_grad():
            m(x * 2.0)                                       # running statistics change -> the cached eval parameters must not be reused
        m.eval()
        with torch.no_grad():
            again = m(x)
        sd2 = nets.clone_state_dict(m.state_dict())
        want2 = nets.vnet_forward(sd2, x.cpu(), False, False) if kind == "vnet" else nets.dualdecoder2d_forward(sd2, x.cpu(), False, False)
        a = again[0] if isinstance(again, tuple) else again
        c = want2[0] if isinstance(want2, tuple) else want2
        assert rel_err(a, c) < 5e-3, kind

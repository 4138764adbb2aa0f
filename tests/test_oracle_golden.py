"""CPU: the oracle restatement (oracle/*.py) reproduces the golden vectors generated from the
UNMODIFIED reference (oracle/make_golden.py) -- runs anywhere, /root/reference not needed."""
import numpy as np
import torch

from conftest import golden, rel_err, seeded_model, weights_checksum
from oracle import chap_losses as L
from oracle import nets, sliding_window


def test_unet2d_oracle_matches_reference_fixture():
    g = golden("unet2d.npz")
    m = seeded_model("dualdecoder2d")
    assert abs(weights_checksum(m.state_dict()) - float(g["weights_checksum"])) < 1e-6   # same init as the reference
    sd = nets.clone_state_dict(m.state_dict(), requires_grad=True)
    x = torch.from_numpy(g["x"])
    o1, o2, feats = nets.dualdecoder2d_forward(sd, x, True, True, None, True)
    assert np.array_equal(o1.detach().numpy(), g["o1"]) and np.array_equal(o2.detach().numpy(), g["o2"])
    np.testing.assert_allclose([f.double().sum().item() for f in feats], g["feat_sums"], rtol=1e-12)
    w = torch.linspace(-1.0, 1.0, o1.numel()).reshape(o1.shape)
    loss = (o1 * w).sum() + (o2 * w.flip(0)).sum()
    names = [str(s) for s in g["grad_names"]]
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    np.testing.assert_allclose([t.double().norm().item() for t in grads], g["grad_norms"], rtol=1e-5, atol=5e-3)   # conv biases feeding a BatchNorm have analytically ZERO gradient: pure rounding noise
    assert rel_err(grads[names.index("encoder.in_conv.conv_conv.0.weight")], g["grad_in_conv"]) < 1e-5   # fp32 reduction-order noise only
    assert rel_err(grads[names.index("decoder2.up4.up.weight")], g["grad_up4_t"]) < 1e-5
    assert np.array_equal(sd["encoder.in_conv.conv_conv.1.running_mean"].numpy(), g["running_mean0"])
    assert np.array_equal(sd["encoder.in_conv.conv_conv.1.running_var"].numpy(), g["running_var0"])
    with torch.no_grad():
        e1, e2 = nets.dualdecoder2d_forward(sd, x, False)
    assert np.array_equal(e1.numpy(), g["eval_o1"]) and np.array_equal(e2.numpy(), g["eval_o2"])
    u = seeded_model("unet2d")
    assert abs(weights_checksum(u.state_dict()) - float(g["unet_weights_checksum"])) < 1e-6
    uo, uf = nets.unet2d_forward(nets.clone_state_dict(u.state_dict()), x, True, True, None, True)
    assert np.array_equal(uo.numpy(), g["unet_o"])


def test_vnet3d_oracle_matches_reference_fixture():
    g = golden("vnet3d.npz")
    m = seeded_model("dualdecoder3d")
    assert abs(weights_checksum(m.state_dict()) - float(g["weights_checksum"])) < 1e-6
    sd = nets.clone_state_dict(m.state_dict(), requires_grad=True)
    x = torch.from_numpy(g["x"])
    o1, o2 = nets.dualdecoder3d_forward(sd, x, True, True, False)
    assert np.array_equal(o1.detach().numpy(), g["o1"]) and np.array_equal(o2.detach().numpy(), g["o2"])
    w = torch.linspace(-1.0, 1.0, o1.numel()).reshape(o1.shape)
    loss = (o1 * w).sum() + (o2 * w.flip(0)).sum()
    names = [str(s) for s in g["grad_names"]]
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    np.testing.assert_allclose([t.double().norm().item() for t in grads], g["grad_norms"], rtol=1e-5, atol=5e-2)   # conv biases feeding a BatchNorm have analytically ZERO gradient: pure rounding noise
    assert rel_err(grads[names.index("encoder.block_one_dw.conv.0.weight")], g["grad_dw"]) < 1e-5
    v = seeded_model("vnet")
    assert abs(weights_checksum(v.state_dict()) - float(g["vnet_weights_checksum"])) < 1e-6
    with torch.no_grad():
        out = nets.vnet_forward(nets.clone_state_dict(v.state_dict()), x, False)
    assert np.array_equal(out.numpy(), g["vnet_eval"])


def _sw_net(g):
    conv = torch.nn.Conv3d(1, 2, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g["conv_w"]))
        conv.bias.copy_(torch.from_numpy(g["conv_b"]))

    def net(patch):
        with torch.no_grad():
            return (torch.tanh(conv(torch.from_numpy(patch))) * 3.0).numpy()
    return net


def test_sliding_window_oracle_matches_reference_fixture():
    g = golden("sliding_window.npz")
    net = _sw_net(g)
    torch_softmax = lambda y: torch.softmax(torch.from_numpy(y), dim=1).numpy()   # noqa: E731 (reference: torch.softmax, :64)
    for name in ("clamped", "padded", "exact"):
        cfg = g[name + "_cfg"]
        patch, sxy, sz = tuple(int(v) for v in cfg[:3]), int(cfg[3]), int(cfg[4])
        label, score, cnt = sliding_window.test_single_case(net, g[name + "_image"], sxy, sz, patch, 2,
                                                            softmax_fn=torch_softmax, return_maps=True)
        assert label.dtype == np.int64 and np.array_equal(label, g[name + "_label"]), name
        assert cnt.min() >= 1 and float(cnt.max()) == cnt.max().round()
        np.testing.assert_allclose(score.sum(axis=0), 1.0, atol=1e-5)


def test_sliding_window_helpers():
    assert sliding_window.window_starts(40, 32, 8) == [0, 8]
    assert sliding_window.window_starts(44, 32, 8) == [0, 8, 12]           # last window clamped to the border
    assert sliding_window.window_starts(32, 32, 8) == [0]
    assert sliding_window.window_starts(192, 112, 18) == [0, 18, 36, 54, 72, 80]
    assert len(sliding_window.window_starts(88, 80, 4)) == 3               # 192x192x88 -> 6*6*3 = 108 windows
    assert sliding_window.pad_amounts((20, 40, 12), (32, 32, 16)) == [(6, 6), (0, 0), (2, 2)]
    assert sliding_window.pad_amounts((21, 40, 13), (32, 32, 16)) == [(5, 6), (0, 0), (1, 2)]


def test_frozen_losses_have_not_drifted():
    g = golden("frozen_losses.npz")
    logits, logits2 = torch.from_numpy(g["logits"]), torch.from_numpy(g["logits2"])
    lab_a, lab_b, mask = torch.from_numpy(g["lab_a"]), torch.from_numpy(g["lab_b"]), torch.from_numpy(g["mask"])
    li, lp, tot = L.mix_loss(logits, lab_a, lab_b, mask, 4, unlab=True)
    np.testing.assert_allclose([li.item(), lp.item(), tot.item()], g["mix"], rtol=1e-6)
    soft1, soft2, ps1, ps2, know = L.pseudo_label_block(logits, logits2)
    np.testing.assert_allclose(know.numpy(), g["knowledge"], rtol=1e-6, atol=1e-6)
    dm = L.create_mask_v1(ps1, ps2, know, 4, 0.25)
    assert np.array_equal(dm.numpy(), g["diff_mask"])
    np.testing.assert_allclose([L.kl_consistency(logits, soft2, dm).item(), L.kl_consistency(logits, soft2, None).item()], g["kl"], rtol=1e-6)
    np.testing.assert_allclose([L.kl_consistency(logits, soft2, dm, "batchmean").item(),
                                L.kl_consistency(logits, soft2, None, "batchmean").item()], g["kl_batchmean"], rtol=1e-6)
    # the two reductions differ by exactly the number of positions per sample
    np.testing.assert_allclose(g["kl_batchmean"], g["kl"] * 16 * 16, rtol=1e-6)
    np.testing.assert_allclose([L.dice_consistency(logits, soft2, dm).item(), L.dice_consistency(logits, soft2, None).item()], g["dice"], rtol=1e-6)
    gf = torch.from_numpy(g["gfield"])
    for key, mode in (("r_cs", "channel_spatial"), ("r_c", "channel"), ("r_s", "spatial"), ("r_n", "sample")):
        r = L.perturbation(gf, 6.0, mode)
        np.testing.assert_allclose(r.numpy(), g[key], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r.reshape(2, -1).norm(dim=1).numpy(), 6.0, rtol=1e-5)   # ||r||_2 = eps per sample
    np.testing.assert_allclose([L.sigmoid_rampup(t, 50.0) for t in (0, 10, 25, 50, 80)], g["rampup"], rtol=1e-12)
    np.testing.assert_allclose([L.poly_lr(0.01, t, 30000) for t in (0, 1, 15000, 29999)], g["poly"], rtol=1e-12)


def test_dice_and_mix_loss_properties():
    torch.manual_seed(0)
    lab = torch.randint(0, 4, (2, 8, 8))
    perfect = torch.nn.functional.one_hot(lab, 4).permute(0, 3, 1, 2).float() * 50.0
    ones = torch.ones(2, 1, 8, 8, dtype=torch.int64)
    d = L.dice_loss_bcp(torch.softmax(perfect, 1), lab.unsqueeze(1), ones, 4)
    assert abs(d.item()) < 1e-5                                         # perfect prediction -> dice loss 0
    zero = L.dice_loss_bcp(torch.softmax(perfect, 1), lab.unsqueeze(1), ones * 0, 4)
    assert abs(zero.item()) < 1e-6                                      # empty mask -> (0+e)/(0+e) -> loss 0


def test_oracle_perform_dropout_reproduces_reference_fixture():
    """tests/golden/filter_dropout.npz holds the reference's perform_dropout outputs and the factors it applied."""
    from oracle import filter_dropout as ofd
    g = golden("filter_dropout.npz")
    feats = [torch.from_numpy(g["feat%d" % i]) for i in range(3)]
    for name in ("binomial_comp", "dropout2d", "scores", "scores_comp"):
        masks = [(torch.from_numpy(g["%s_m1_%d" % (name, i)]), torch.from_numpy(g["%s_m2_%d" % (name, i)])) if i in (0, 2) else None
                 for i in range(3)]
        o1, o2 = ofd.perform_dropout(feats, masks)
        for i in range(3):
            np.testing.assert_allclose(o1[i].numpy(), g["%s_fp1_%d" % (name, i)], rtol=2e-7, atol=0)
            np.testing.assert_allclose(o2[i].numpy(), g["%s_fp2_%d" % (name, i)], rtol=2e-7, atol=0)
        for i in (0, 2):                                  # Dropout2d / Binomial factors are 0 or 2; score masks are rescaled to mean 1
            m = g["%s_m1_%d" % (name, i)]
            assert m.shape == (2, feats[i].shape[1]) and (m >= 0).all()

"""GPU: the product networks (chap_b200.networks) against the reference -- through the committed golden
fixtures (generated from the unmodified reference) and through the functional oracle on the same weights and
inputs.

Tolerances (relative L2 against the fp32 CPU oracle / fixture), BASELINE.json north_star asks <= 1e-3 on logits:
  * fp32 CUDA-core mode (ops.set_force_simt(True)):                       logits <= 1e-4 (measured ~1e-6)
  * 3xTF32 tensor-core mode (ops.set_conv_precision(ops.PRECISE_ALL)):    logits <= 1e-3 (measured ~5e-6 2D, ~2e-4 3D)
  * TF32 tensor-core mode (default, = the reference's default cuDNN setting `allow_tf32=True`): no constant --
        <= 1.25x the deviation of the IDEAL TF32 evaluation of the same network (tests/tf32_emul.py: float64 arithmetic,
        conv operands rounded to TF32).  Measured at BASELINE shapes (tests/test_gpu_baseline_shapes.py): product 1.684e-3
        = ideal 1.684e-3 (2D), 1.49e-2 vs 1.47e-2 (3D); torch-eager + cuDNN allow_tf32 on the same box: 2.2e-3 / 1.5e-2.
"""
import numpy as np
import pytest
import torch

from conftest import golden, max_err, rel_err, seeded_model, weights_checksum
from oracle import nets
from tf32_emul import ideal_tf32

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = ["fp32-cuda-core", "tf32-tensor-core", "3xtf32-tensor-core"]


def _oracle64(kind, state_dict, x, loss_fn=None, names=None, emulate=False, drop=None, update_running=False):
    """float64 evaluation of the oracle on the GPU (optionally as ideal TF32): (o1, o2), grads of loss_fn w.r.t. `names`."""
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in nets.clone_state_dict(state_dict, device=DEV).items()}
    for n in names or []:
        sd[n].requires_grad_(True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ctx = ideal_tf32() if emulate else torch.enable_grad()
        with ctx:
            if kind == "3d":
                o1, o2 = nets.dualdecoder3d_forward(sd, x.to(DEV).double(), True, update_running, drop is not None, drop)
            else:
                o1, o2 = nets.dualdecoder2d_forward(sd, x.to(DEV).double(), True, update_running, drop)
            grads = torch.autograd.grad(loss_fn(o1, o2), [sd[n] for n in names]) if names else None
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return (o1.detach(), o2.detach()), grads


def ideal_tf32_devs(kind, state_dict, x, loss_fn=None, names=None, drop=None):
    """Deviation of the ideal-TF32 evaluation from the unrounded float64 one: (logit deviation, [per-tensor gradient deviation])."""
    if drop is not None:
        drop = {k: v.to(DEV).double() for k, v in drop.items()}
    o, g = _oracle64(kind, state_dict, x, loss_fn, names, False, drop)
    oi, gi = _oracle64(kind, state_dict, x, loss_fn, names, True, drop)
    return max(rel_err(a, b) for a, b in zip(oi, o)), ([rel_err(a, b) for a, b in zip(gi, g)] if names else None)


def logit_tol(mode, kind, state_dict=None, x=None, drop=None):
    if mode == "fp32-cuda-core":
        return 1e-4
    if mode == "3xtf32-tensor-core":
        return 1e-3
    return 1.25 * ideal_tf32_devs(kind, state_dict, x, drop=drop)[0] + 1e-5


@pytest.fixture(params=MODES)
def mode(request):
    from chap_b200 import ops
    ops.set_force_simt(request.param == "fp32-cuda-core")
    ops.set_conv_precision(ops.PRECISE_ALL if request.param == "3xtf32-tensor-core" else 0)
    yield request.param
    ops.set_force_simt(False)
    ops.set_conv_precision(0)


def assert_grads(mode, kind, mine, want, ideal_dev):
    """fp32 mode: fixed tolerance; 3xTF32: every conv op is at 1e-6 forward and data gradient (tools/_diag / replay), the
    weight-gradient GEMM is one TF32 rounding deep (3e-4), and a handful of (Leaky)ReLU units whose pre-activation lies within
    the 5e-6 forward error of zero flip their derivative -- each flip moves a small tensor like the 16x1x3x3 stem gradient by
    ~1e-3 (DESIGN.md "conditioning"); TF32: no worse than 2x the ideal-TF32 deviation of that tensor (+ a small floor)."""
    for i, (a, b) in enumerate(zip(mine, want)):
        if mode == "fp32-cuda-core":
            lim = 2e-4 if kind == "2d" else 5e-2
        elif mode == "3xtf32-tensor-core":
            # weight gradients: the wgrad GEMM stays plain TF32 in this mode -> the TF32 yardstick of that tensor, with a floor for
            # the derivative flips described above
            lim = max(2.0 * ideal_dev[i] + 2e-3, 1e-2 if kind == "2d" else 5e-2)
        else:
            lim = 2.0 * ideal_dev[i] + 2e-3
        assert rel_err(a, b) < lim, (i, rel_err(a, b), lim)


def test_dualdecoder2d_matches_reference_fixture(mode):
    g = golden("unet2d.npz")
    m = seeded_model("dualdecoder2d")
    assert abs(weights_checksum(m.state_dict()) - float(g["weights_checksum"])) < 1e-6
    sd0 = seeded_model("dualdecoder2d").state_dict()
    m = m.to(DEV).train()
    x = torch.from_numpy(g["x"]).to(DEV)
    simt = mode == "fp32-cuda-core"
    tol = logit_tol(mode, "2d", sd0, x)
    o1, o2, feats = m(x, with_feat=True)
    assert o1.shape == (2, 4, 48, 48)
    assert rel_err(o1, g["o1"]) < tol and rel_err(o2, g["o2"]) < tol
    np.testing.assert_allclose([f.double().sum().item() for f in feats], g["feat_sums"], rtol=5 * tol)
    w = torch.linspace(-1.0, 1.0, o1.numel()).reshape(o1.shape).to(DEV)
    loss = (o1 * w).sum() + (o2 * w.flip(0)).sum()
    names = [str(s) for s in g["grad_names"]]
    params = dict(m.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    sel = ["encoder.in_conv.conv_conv.0.weight", "decoder1.out_conv.weight", "decoder2.up4.up.weight"]
    want = [g["grad_in_conv"], g["grad_out1"], g["grad_up4_t"]]
    lin = lambda a, b: (a * w.double()).sum() + (b * w.double().flip(0)).sum()      # noqa: E731
    dev = ideal_tf32_devs("2d", sd0, x, lin, sel)[1] if mode != "fp32-cuda-core" else None
    assert_grads(mode, "2d", [grads[names.index(n)] for n in sel], want, dev)
    norms = np.array([t.double().norm().item() for t in grads])
    big = g["grad_norms"] > 1.0                      # pre-BN conv biases have analytically zero gradient
    if mode != "fp32-cuda-core":                     # norms of ALL parameter gradients: as close as the ideal TF32 evaluation (x1.5);
        # the weight-gradient GEMM is plain TF32 in the 3xTF32 mode too, and this linear zero-mean functional cancels heavily
        all_names = [n for n in names if big[names.index(n)]]
        (_, g64) = _oracle64("2d", sd0, x, lin, all_names)
        (_, gid) = _oracle64("2d", sd0, x, lin, all_names, emulate=True)
        ideal_norm_dev = max(abs(float(a.norm()) / float(b.norm()) - 1.0) for a, b in zip(gid, g64))
        np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=1.5 * ideal_norm_dev + 1e-3)
    else:
        np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=1e-3)
    assert rel_err(m.encoder.in_conv.conv_conv[1].running_mean, g["running_mean0"]) < 1e-4
    assert rel_err(m.encoder.in_conv.conv_conv[1].running_var, g["running_var0"]) < 1e-4
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x)
    assert rel_err(e1, g["eval_o1"]) < tol and rel_err(e2, g["eval_o2"]) < tol
    u = seeded_model("unet2d").to(DEV).train()
    uo, uf = u(x, with_feats=True)
    assert rel_err(uo, g["unet_o"]) < tol and uf.shape == (2, 16, 48, 48)


def test_dualdecoder3d_and_vnet_match_reference_fixture(mode):
    g = golden("vnet3d.npz")
    sd0 = seeded_model("dualdecoder3d").state_dict()
    m = seeded_model("dualdecoder3d").to(DEV).train()
    x = torch.from_numpy(g["x"]).to(DEV)
    tol = logit_tol(mode, "3d", sd0, x)
    o1, o2 = m(x)
    assert rel_err(o1, g["o1"]) < tol and rel_err(o2, g["o2"]) < tol
    w = torch.linspace(-1.0, 1.0, o1.numel()).reshape(o1.shape).to(DEV)
    loss = (o1 * w).sum() + (o2 * w.flip(0)).sum()
    names = [str(s) for s in g["grad_names"]]
    params = dict(m.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    # The V-Net backward is ill-conditioned in fp32: the reference's OWN fp32 CPU gradients differ from an fp64
    # evaluation by 1.7e-2 (rel. L2) on this fixture (measured, DESIGN.md "conditioning"); the sharp check is the
    # per-op replay test below.
    sel = ["encoder.block_one.conv.0.weight", "encoder.block_one_dw.conv.0.weight", "decoder2.block_eight_up.conv.0.weight"]
    want = [g["grad_block_one"], g["grad_dw"], g["grad_up_t"]]
    lin = lambda a, b: (a * w.double()).sum() + (b * w.double().flip(0)).sum()      # noqa: E731
    dev = ideal_tf32_devs("3d", sd0, x, lin, sel)[1] if mode != "fp32-cuda-core" else None
    if dev is not None:
        dev = [d + 2.5e-2 for d in dev]               # the fp32 fixture itself is 1.7e-2 from an fp64 evaluation (see above)
    assert_grads(mode, "3d", [grads[names.index(n)] for n in sel], want, dev)
    v = seeded_model("vnet").to(DEV).eval()
    with torch.no_grad():
        out = v(x)
    assert rel_err(out, g["vnet_eval"]) < tol                        # eval-mode BN: better conditioned than the train-mode yardstick


def _oracle_grads(kind, sd_src, x, dtype, loss_fn):
    sd = {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in nets.clone_state_dict(sd_src).items()}
    names = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    for k in names:
        sd[k].requires_grad_(True)
    if kind == "3d":
        o1, o2 = nets.dualdecoder3d_forward(sd, x.to(dtype), True, True, False)
    else:
        o1, o2 = nets.dualdecoder2d_forward(sd, x.to(dtype), True, True, None)
    return names, torch.autograd.grad(loss_fn(o1, o2), [sd[n] for n in names]), (o1.detach(), o2.detach())


@pytest.mark.parametrize("kind,force_simt", [("2d", True), ("2d", False), ("3d", False)])
def test_every_op_of_a_network_pass_matches_torch_fp64_on_identical_inputs(kind, force_simt):
    """Sharp check (tests/replay.py): each recorded op, re-run alone, must match a float64 torch restatement on
    the same inputs and upstream gradient.  One outlier op is tolerated (a pre-activation within float rounding
    of the (Leaky)ReLU kink inside that very op)."""
    import replay
    from chap_b200 import ops
    m = seeded_model("dualdecoder3d" if kind == "3d" else "dualdecoder2d", seed=17).to(DEV).train()
    x = (torch.randn(2, 1, 16, 16, 16) if kind == "3d" else torch.rand(2, 1, 48, 48)).to(DEV)
    ops.set_force_simt(force_simt)
    try:
        def run():
            o1, o2 = m(x)
            (torch.softmax(o1, 1)[:, 0].mean() + (torch.softmax(o2, 1)[:, 1] ** 2).mean()).backward()
        rows = replay.check(replay.record(run))
    finally:
        ops.set_force_simt(False)
    assert len(rows) > 60          # (the skip-connection concats are part of the conv op)
    tol = 2e-5 if force_simt else 3e-3            # fp32 CUDA-core kernels / TF32 tensor-core kernels (one layer)
    bad = [r for r in rows if r[2] > tol or r[3] > tol]
    assert len(bad) <= 1, bad[:5]


@pytest.mark.parametrize("kind", ["2d", "3d"])
def test_end_to_end_gradients_vs_fp64_oracle(kind):
    """End-to-end parameter gradients (fp32 mode) against an fp64 evaluation of the oracle.  The derivative of the
    network is discontinuous at activation kinks, so two correct fp32 implementations can differ by ~1/sqrt(N) in
    one channel of a BatchNorm gradient, which everything upstream inherits (measured: 4e-2 in one channel of the
    last decoder block -> 5e-3..9e-3 on every upstream tensor, while every op matched torch fp64 to 1e-7 on
    identical inputs).  Bound: median <= 2e-2, max <= 5e-2; the sharp per-op check is the replay test above."""
    from chap_b200 import ops
    m = seeded_model("dualdecoder3d" if kind == "3d" else "dualdecoder2d", seed=17)
    x = torch.randn(2, 1, 16, 32, 16) if kind == "3d" else torch.rand(2, 1, 48, 48)
    loss_fn = lambda o1, o2: torch.softmax(o1, 1)[:, 0].mean() + (torch.softmax(o2, 1)[:, 1] ** 2).mean()   # noqa: E731
    names, g64, o64 = _oracle_grads(kind, m.state_dict(), x, torch.float64, loss_fn)
    m = m.to(DEV).train()
    ops.set_force_simt(True)
    try:
        o1, o2 = m(x.to(DEV))
        assert rel_err(o1, o64[0]) < 1e-3 and rel_err(o2, o64[1]) < 1e-3
        params = dict(m.named_parameters())
        gg = torch.autograd.grad(loss_fn(o1, o2), [params[n] for n in names])
    finally:
        ops.set_force_simt(False)
    scale = max(float(t.norm()) for t in g64)
    e_gpu = np.array([rel_err(a, b64) for a, b64 in zip(gg, g64) if float(b64.norm()) >= 1e-3 * scale])
    assert np.median(e_gpu) < 2e-2, np.median(e_gpu)
    assert e_gpu.max() < 5e-2, e_gpu.max()


def test_dualdecoder2d_vs_oracle_with_dropout_masks_and_odd_batch(mode):
    """same weights / inputs / explicit dropout masks on both sides, batch 3, 64x64."""
    torch.manual_seed(5)
    m = seeded_model("dualdecoder2d", seed=21).to(DEV).train()
    sd = nets.clone_state_dict(m.state_dict(), requires_grad=True)
    x = torch.rand(3, 1, 64, 64)
    ps = (0.05, 0.1, 0.2, 0.3, 0.5)
    chans, sizes = (16, 32, 64, 128, 256), (64, 32, 16, 8, 4)
    masks = [(torch.rand(3, c, s, s) > p).float() / (1 - p) for c, s, p in zip(chans, sizes, ps)]
    keys = ["encoder.in_conv.conv_conv.3"] + ["encoder.down%d.maxpool_conv.1.conv_conv.3" % i for i in range(1, 5)]
    o1r, o2r = nets.dualdecoder2d_forward(sd, x, True, True, dict(zip(keys, masks)))
    feats = m.encoder(x.to(DEV), [t.to(DEV) for t in masks])
    o1, o2 = m.decoder1(feats), m.decoder2(feats)
    sd0 = {k: v.detach() for k, v in sd.items()}
    drop = dict(zip(keys, masks))
    quad = lambda a, b: (a ** 2).sum() + (b ** 2).sum()      # noqa: E731
    names = [n for n, _ in m.named_parameters()]
    tol = logit_tol(mode, "2d", sd0, x, drop)
    assert rel_err(o1, o1r) < tol and rel_err(o2, o2r) < tol
    gr = torch.autograd.grad(quad(o1r, o2r), [sd[n] for n in names])
    gg = torch.autograd.grad(quad(o1, o2), list(m.parameters()))
    sel = [i for i, b in enumerate(gr) if b.norm() > 1e-2]
    errs = np.array([rel_err(gg[i], gr[i]) for i in sel])
    if mode == "fp32-cuda-core":
        lim_max, lim_med = 1e-2, 1e-3
    elif mode == "3xtf32-tensor-core":
        lim_max, lim_med = 2e-2, 2e-3                 # plain-TF32 weight-gradient GEMM on top of split-operand fwd / dgrad
    else:
        # TF32: the ideal TF32 evaluation of this very network is the bar (derivative-kink flips make its gradients
        # percent-level too): no more than 1.5x its median / maximum deviation
        ideal = np.array(ideal_tf32_devs("2d", sd0, x, quad, [names[i] for i in sel], drop)[1])
        lim_max, lim_med = 1.5 * ideal.max() + 1e-3, 1.5 * np.median(ideal) + 1e-4
    assert errs.max() < lim_max and np.median(errs) < lim_med, (errs.max(), lim_max, np.median(errs), lim_med)


def test_vnet_train_dropout3d_masks_vs_oracle(mode):
    m = seeded_model("dualdecoder3d", seed=4)
    for mod in m.modules():
        if hasattr(mod, "has_dropout"):
            mod.has_dropout = True
    m = m.to(DEV).train()
    sd = nets.clone_state_dict(m.state_dict())
    x = torch.randn(2, 1, 32, 32, 32)             # 16 values per channel at the deepest level (BN conditioning)
    d5 = (torch.rand(2, 256) > 0.5).float() * 2
    d9a, d9b = (torch.rand(2, 16) > 0.5).float() * 2, (torch.rand(2, 16) > 0.5).float() * 2
    drop = {"encoder.dropout": d5.reshape(2, 256, 1, 1, 1), "decoder1.dropout": d9a.reshape(2, 16, 1, 1, 1),
            "decoder2.dropout": d9b.reshape(2, 16, 1, 1, 1)}
    o1r, o2r = nets.dualdecoder3d_forward(sd, x, True, True, True, drop)
    feats = m.encoder(x.to(DEV), d5.to(DEV))
    o1, o2 = m.decoder1(feats, d9a.to(DEV)), m.decoder2(feats, d9b.to(DEV))
    tol = logit_tol(mode, "3d", sd, x, drop)
    assert rel_err(o1, o1r) < tol and rel_err(o2, o2r) < tol


def test_outputs_are_logically_nchw_and_checkpoint_roundtrip(tmp_path):
    m = seeded_model("dualdecoder2d").to(DEV).eval()
    x = torch.rand(1, 1, 32, 32, device=DEV)
    with torch.no_grad():
        o1, o2 = m(x)
    assert o1.shape == (1, 4, 32, 32) and o1.dtype == torch.float32
    path = str(tmp_path / "latest.pth")
    torch.save(m.state_dict(), path)                       # code/train_ours_2D.py:428-429
    m2 = seeded_model("dualdecoder2d", seed=3).to(DEV).eval()
    m2.load_state_dict(torch.load(path))                   # code/test_2D_fully.py:115-117
    with torch.no_grad():
        p1, _ = m2(x)
    assert max_err(o1, p1) == 0.0

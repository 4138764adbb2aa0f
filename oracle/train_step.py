"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product package.

CPU restatement of ONE CHAP training iteration, transcribed from
code/train_ours_2D.py:304-389 (flags: --adv_noise, --adv_losstype kl|dice, no --dropout),
dimension-generic so that the same function is the 3D (LA-shaped) oracle -- the
reference ships no 3D trainer, the 3D step is "by analogy" and
therefore frozen here.

`OracleModel` adapts the functional nets of oracle/nets.py (driven by a state dict) to
the small protocol the step and VAT need; `ModuleModel` adapts a real nn.Module (the
imported reference classes) to the same protocol, so the step can be run on either.
"""
import contextlib

import torch

from . import chap_losses as L
from . import filter_dropout
from . import nets


class OracleModel:
    """state-dict driven DualDecoder (2D) / DualDecoder3d (3D) in train mode."""

    def __init__(self, sd, dims=2, has_dropout=False, drop=None):
        self.sd, self.dims, self.has_dropout, self.drop = sd, dims, has_dropout, drop
        self._track = True

    @contextlib.contextmanager
    def bn_tracking(self, flag):
        old, self._track = self._track, flag
        try:
            yield
        finally:
            self._track = old

    def params(self):
        return [v for v in self.sd.values() if v.requires_grad]

    def __call__(self, x):
        if self.dims == 2:
            return nets.dualdecoder2d_forward(self.sd, x, True, self._track, self.drop)
        return nets.dualdecoder3d_forward(self.sd, x, True, self._track, self.has_dropout, self.drop)

    def encoder(self, x):
        mode = nets.BNMode(True, self._track)
        if self.dims == 2:
            return nets.unet_encoder(self.sd, x, mode, self.drop)
        return nets.vnet_encoder(self.sd, x, mode, self.has_dropout, self.drop)

    def _dec(self, feats, which):
        if self.dims == 2:
            return nets.dualdecoder2d_decode(self.sd, feats, which, True, self._track)
        return nets.dualdecoder3d_decode(self.sd, feats, which, True, self._track, self.has_dropout, self.drop)

    def decoder1(self, feats):
        return self._dec(feats, 1)

    def decoder2(self, feats):
        return self._dec(feats, 2)


class ModuleModel:
    """nn.Module (reference DualDecoder / DualDecoder3d) behind the same protocol."""

    def __init__(self, module):
        self.m = module

    @contextlib.contextmanager
    def bn_tracking(self, flag):
        bns = [m for m in self.m.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        old = [b.track_running_stats for b in bns]
        saved = [(b.running_mean, b.running_var, b.num_batches_tracked) for b in bns]
        try:
            if not flag:   # batch statistics, buffers untouched
                for b in bns:
                    b.track_running_stats = False
                    b.running_mean = b.running_var = b.num_batches_tracked = None
            yield
        finally:
            for b, o, s in zip(bns, old, saved):
                b.track_running_stats = o
                b.running_mean, b.running_var, b.num_batches_tracked = s

    def params(self):
        return [p for p in self.m.parameters() if p.requires_grad]

    def __call__(self, x):
        return self.m(x)

    def encoder(self, x):
        return self.m.encoder(x)

    def decoder1(self, feats):
        return self.m.decoder1(feats)

    def decoder2(self, feats):
        return self.m.decoder2(feats)


def feature_dropout_loss(model, uimg_ab, ps1, ps2, dropout_masks):
    """The --dropout branch, code/train_ours_2D.py:359-364: model(uimg_ab, False, True, [0..4], sim_score, False) -- encoder,
    perform_dropout (FilterDropout.py:45-89: the second half of uimg_ab gets a channel-masked copy appended per decoder), both
    decoders -- then CE against the other decoder's pseudo-labels.
    FROZEN FIX: as shipped the reference calls F.cross_entropy(outputs1_fp [1.5 N], pseudo_outputs2 [N]) (:362-363), which
    raises on the batch-size mismatch; the appended rows are perturbed copies of samples N/2.., so their targets are the
    pseudo-labels of those samples: target = cat(pseudo, pseudo[N/2:])."""
    import torch.nn.functional as F
    feats = model.encoder(uimg_ab)
    f1, f2 = filter_dropout.perform_dropout(feats, dropout_masks)
    o1, o2 = model.decoder1(f1), model.decoder2(f2)
    half = ps1.shape[0] // 2
    t1, t2 = torch.cat((ps1, ps1[half:])), torch.cat((ps2, ps2[half:]))
    return F.cross_entropy(o1, t2.long()) + F.cross_entropy(o2, t1.long())           # :362-364


def chap_losses_forward(model, volume, label, labeled_bs, n_classes, mask_offsets, iter_num,
                        vat=None, adv_losstype="kl", topk=0.1, use_diff_mask=True,
                        consistency=1.0, rampup=50.0, d_init=None, trace=None, dropout_masks=None):
    """Forward part of one iteration: returns (loss, aux dict).  Lines refer to
    code/train_ours_2D.py."""
    n = volume.shape[0]
    sub_l, sub_u = labeled_bs // 2, (n - labeled_bs) // 2                       # :295
    img_a, img_b = volume[:sub_l], volume[sub_l:labeled_bs]                     # :307
    uimg_a, uimg_b = volume[labeled_bs:labeled_bs + sub_u], volume[labeled_bs + sub_u:]   # :308
    ulab_a, ulab_b = label[labeled_bs:labeled_bs + sub_u], label[labeled_bs + sub_u:]     # :309
    lab_a, lab_b = label[:sub_l], label[sub_l:labeled_bs]                       # :310
    uimg_ab = torch.cat((uimg_a, uimg_b))                                       # :312

    with torch.no_grad():                                                       # :314-333
        pre1, pre2 = model(uimg_ab)
        soft1, soft2, ps1, ps2, knowledge = L.pseudo_label_block(pre1, pre2)
        pre_a1, pre_b1 = pre1.chunk(2)
        pre_a2, pre_b2 = pre2.chunk(2)
        plab_a1 = L.get_masks(pre_a1, n_classes)
        plab_b1 = L.get_masks(pre_b1, n_classes)
        plab_a2 = L.get_masks(pre_a2, n_classes)
        plab_b2 = L.get_masks(pre_b2, n_classes)
        img_mask = L.generate_mask(volume.shape[2:], mask_offsets, volume.device)
        loss_mask = img_mask.unsqueeze(0).expand((sub_l,) + tuple(img_mask.shape))

    net_input_unl = uimg_a * img_mask + img_a * (1 - img_mask)                  # :335
    net_input_l = img_b * img_mask + uimg_b * (1 - img_mask)                    # :336
    out1, out2 = model(torch.cat((net_input_l, net_input_unl)))                 # :338-339
    out_l1, out_unl1 = out1.chunk(2)
    out_l2, out_unl2 = out2.chunk(2)
    kw = dict(n_classes=n_classes, u_weight=0.5)
    lu_o1, ll_i1, m1 = L.mix_loss(out_unl1, plab_a2, lab_a, loss_mask, unlab=True, **kw)   # :345
    lu_o2, ll_i2, m2 = L.mix_loss(out_unl2, plab_a1, lab_a, loss_mask, unlab=True, **kw)   # :346
    ll_o1, lu_i1, m3 = L.mix_loss(out_l1, lab_b, plab_b2, loss_mask, **kw)                 # :348
    ll_o2, lu_i2, m4 = L.mix_loss(out_l2, lab_b, plab_b1, loss_mask, **kw)                 # :349
    bcp_loss = m1 + m2 + m3 + m4                                                # :351
    loss_l = ll_i1 + ll_i2 + ll_o1 + ll_o2                                      # :353
    loss_u = lu_i1 + lu_i2 + lu_o1 + lu_o2                                      # :354
    cw = L.consistency_weight(iter_num, consistency, rampup)                   # :356

    # :359-367 -- dropout_masks given <=> args["dropout"]; explicit masks (list of 5: None or (m1, m2)) replace the RNG draws
    fp_loss = feature_dropout_loss(model, uimg_ab, ps1, ps2, dropout_masks) if dropout_masks is not None else torch.zeros((), device=volume.device)
    if vat is not None:                                                         # :369-372
        diff_mask = L.create_mask_v1(ps1, ps2, knowledge, 4, topk) if use_diff_mask else None
        vat_loss = vat(model, volume, soft1, soft2, diff_mask, adv_losstype, d_init=d_init, trace=trace)
    else:
        vat_loss = torch.zeros((), device=volume.device)
    loss = bcp_loss + cw * (fp_loss + vat_loss)                                 # :378
    aux = dict(bcp_loss=bcp_loss.detach(), vat_loss=vat_loss.detach(), fp_loss=fp_loss.detach(), loss_l=loss_l.detach(),
               loss_u=loss_u.detach(), cw=cw, soft1=soft1, soft2=soft2, knowledge=knowledge,
               plab=(plab_a1, plab_b1, plab_a2, plab_b2), out_mix=(out1.detach(), out2.detach()))
    return loss, aux


def chap_train_step(model, bufs, volume, label, labeled_bs, n_classes, mask_offsets, iter_num,
                    base_lr=0.01, max_iterations=30000, **kw):
    """loss.backward(); SGD(momentum .9, wd 1e-4).step(); poly LR  (:381-389).
    `bufs` is the list of momentum buffers (None entries before the first step).
    The learning rate used at iteration `iter_num` (0-based) is poly_lr(iter_num)."""
    params = model.params()
    loss, aux = chap_losses_forward(model, volume, label, labeled_bs, n_classes, mask_offsets, iter_num, **kw)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
    lr = L.poly_lr(base_lr, iter_num, max_iterations)
    L.sgd_momentum_step(params, grads, bufs, lr)
    aux["loss"] = loss.detach()
    aux["grads"] = grads
    return aux

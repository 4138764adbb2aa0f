"""Loss / perturbation entry points of the CHAP hot path on the sm_100a kernels.

Mirrors the call sites of the reference's (absent) `utils.losses` module and of `mix_loss`:
  * `VAT2d(xi, epi, num_classes)` -> `adv_loss(model, volume_batch, soft1, soft2, diff_mask, losstype)`
    (ctor code/train_ours_2D.py:290, call :372; code/train_ablation_2D.py:148,234)
  * `DiceLoss_bcp(n_classes)` -> `dice_loss(soft, target[N,1,*], mask[N,1,*])`  (:197,206-207)
  * `mix_loss(output, img_l, patch_l, mask, l_weight, u_weight, unlab)`          (:198-216)
The arithmetic follows the frozen specification in oracle/chap_losses.py (the reference ships no
source for VAT2d / DiceLoss_bcp: "parity unpinned").  All reductions over pixels run
in the fused kernels of libchap_b200; only the final combination of a handful of per-class sums is
done with torch scalars (float64), which also carries the autograd link back to the kernels.
"""
import torch

from .. import ops
from .._lib import DIST_DICE, DIST_KL


def _dice_from_sums(inter, a, b):
    return (1.0 - (2.0 * inter + 1e-10) / (a + b + 1e-10)).mean()


def masked_dice_ce(logits, labels, mask, invert=False):
    """(dice, ce): DiceLoss_bcp(softmax(logits), labels, m) and sum(CE * m) / (sum(m) + 1e-16) with
    m = mask (or 1 - mask), in one fused pass over the logits."""
    c = logits.shape[1]
    s = ops.dice_ce_sums(logits, labels, mask, invert)
    dice = _dice_from_sums(s[:c], s[c:2 * c], s[2 * c:3 * c])
    ce = s[3 * c] / (s[3 * c + 1] + 1e-16)
    return dice, ce


def _spatial_mask(mask, logits):
    """The reference's loss_mask is the same spatial mask repeated over the batch (generate_mask,
    code/train_ours_2D.py:91-101); the kernel broadcasts one spatial mask [*spatial]."""
    return mask[0] if mask.dim() == logits.dim() - 1 else mask


def mix_loss(output, img_l, patch_l, mask, l_weight=1.0, u_weight=0.5, unlab=False):
    """code/train_ours_2D.py:198-216.  Returns (loss_image, loss_patch, (dice + ce) / 2) as fp32 scalars."""
    image_weight, patch_weight = (u_weight, l_weight) if unlab else (l_weight, u_weight)
    m = _spatial_mask(mask, output)
    out = ops.mix_loss_fused(output, img_l, patch_l, m, image_weight, patch_weight)      # one autograd node, two tiny scalar kernels
    return out[0], out[1], out[2]


class DiceLoss_bcp:
    """dice_loss(soft, target[N,1,*], mask[N,1,*]) on logits-derived probabilities.

    The fused kernel works from logits; `soft` must therefore be given as logits via `from_logits`
    (what mix_loss does internally) -- passing probabilities is accepted and converted with log()
    (softmax(log p) == p), which keeps the call-site signature of the reference."""

    def __init__(self, n_classes):
        self.n_classes = n_classes

    def from_logits(self, logits, target, mask):
        m = mask[:, 0] if mask.dim() == logits.dim() else mask
        t = target[:, 0] if target.dim() == logits.dim() else target
        return masked_dice_ce(logits, t, _spatial_mask(m, logits))[0].float()

    def __call__(self, soft, target, mask):
        return self.from_logits(torch.log(soft.clamp_min(1e-30)), target, mask)


KL_REDUCTIONS = ("mean", "batchmean")


def consistency_distance(logits, target_soft, mask, losstype, reduction="mean"):
    """'kl': sum(mask * KL(target || softmax(logits))) / (N * positions) ('mean', the frozen default) or / N
    ('batchmean', the classic-VAT scaling that made the round-1 step diverge -- see oracle/chap_losses.py
    kl_consistency);  'dice': masked soft Dice, mean over classes (scale free)."""
    c = logits.shape[1]
    if losstype == "kl":
        if reduction not in KL_REDUCTIONS:
            raise ValueError("reduction must be one of %r" % (KL_REDUCTIONS,))
        s = ops.consistency_sums(logits, target_soft, mask, DIST_KL)
        denom = logits.shape[0] if reduction == "batchmean" else logits.numel() // c
        return (s[0] / denom).float()
    if losstype == "dice":
        s = ops.consistency_sums(logits, target_soft, mask, DIST_DICE)
        return _dice_from_sums(s[:c], s[c:2 * c], s[2 * c:3 * c]).float()
    raise ValueError("losstype must be 'kl' or 'dice'")


class VAT2d:
    """Channel-spatial hierarchical adversarial perturbation at the 5 encoder levels (2D or 3D nets).

    adv_loss = VAT2d(xi, epi, num_classes); loss = adv_loss(model, x, soft1, soft2, mask, 'kl'|'dice').
    `model` exposes .encoder, .decoder1, .decoder2 (DualDecoder / DualDecoder3d).  Steps (frozen
    spec, oracle/chap_losses.py VAT): encoder forward on the unlabelled rows; one probing pass of both
    decoders on f + xi*l2n(d) recording data gradients only; the fused perturbation generator
    (channel-wise + spatial-wise L2 normalisation, eps scaling, injection) for all levels in one
    library call; decoder re-forward on f + r with gradients flowing to encoder and decoders.
    BatchNorm running statistics are not updated inside VAT.  reduction: scaling of the returned 'kl' loss
    ('mean' per pixel, default; 'batchmean' per sample).
    """

    def __init__(self, xi=10.0, epi=6.0, num_classes=4, mode="channel_spatial", reduction="mean"):
        if reduction not in KL_REDUCTIONS:
            raise ValueError("reduction must be one of %r" % (KL_REDUCTIONS,))
        self.xi, self.epi, self.num_classes, self.mode, self.reduction = xi, epi, num_classes, mode, reduction

    def __call__(self, model, x, soft1, soft2, mask=None, losstype="kl", d_init=None, trace=None):
        x_u = x[x.shape[0] - soft1.shape[0]:]
        with ops.bn_tracking(False):
            feats = model.encoder(x_u)
            if d_init is None:
                d_init = [torch.rand_like(f) - 0.5 for f in feats]
            hat = [h.requires_grad_(True) for h in ops.l2n_sample_axpy_all(d_init, feats, self.xi)]
            with ops.no_weight_grad():
                # the probe distance is always 'batchmean' (direction only; keeps |g| far above the 1e-8 of the norms)
                dist = consistency_distance(model.decoder1(hat), soft2, mask, losstype, "batchmean") + \
                    consistency_distance(model.decoder2(hat), soft1, mask, losstype, "batchmean")
            g = torch.autograd.grad(dist, hat)
            adv_values = ops.perturb(g, feats, self.epi, self.mode, g_scale=self.xi)
            adv = [ops.attach_identity_grad(f, v) for f, v in zip(feats, adv_values)]
            loss = consistency_distance(model.decoder1(adv), soft2, mask, losstype, self.reduction) + \
                consistency_distance(model.decoder2(adv), soft1, mask, losstype, self.reduction)
        if trace is not None:
            trace.update(feats=feats, g=g, adv=adv_values, dist=dist)
        return loss

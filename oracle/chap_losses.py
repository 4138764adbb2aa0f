"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product package.

Frozen pure-PyTorch (CPU, fp32) restatement of the loss / perturbation pieces of the
CHAP hot path.

PARITY STATUS
  * mix_loss, the pseudo-label block, generate_mask, the largest-CC filter, the SGD/poly-LR
    step and the step orchestration are transcriptions of reference code that IS present
    (code/train_ours_2D.py:91-144,198-216,304-389) -> pinned by reading; the reference
    ships no golden vectors for them.
  * `sigmoid_rampup`, `DiceLoss_bcp`, `VAT2d`, `create_maskV1` have NO SOURCE in the
    reference drop (the `utils/` package is absent; only constructor + call sites exist:
    code/train_ours_2D.py:36,197,206-207,290,371-372, code/train_ablation_2D.py:148,232-234).
    For those this file is the *frozen specification*: "parity unpinned".  The
    definitions follow BASELINE.json's north_star wording and SURVEY.md Appendix B; every
    free choice is a named parameter so a later-recovered true spec is a config change.

Everything is dimension-generic: tensors are [N, C, *spatial] with 2 (ACDC) or 3 (LA)
spatial dims.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- ramps
def sigmoid_rampup(current, rampup_length):
    """utils.ramps.sigmoid_rampup (absent; call site code/train_ours_2D.py:36, comment
    cites arXiv:1610.02242): exp(-5 (1 - t)^2), t = clip(current, 0, L) / L."""
    if rampup_length == 0:
        return 1.0
    t = min(max(float(current), 0.0), float(rampup_length)) / float(rampup_length)
    return float(math.exp(-5.0 * (1.0 - t) ** 2))


def consistency_weight(iter_num, consistency=1.0, rampup=50.0):
    """code/train_ours_2D.py:34-36,356: consistency * sigmoid_rampup(iter // 150, rampup)."""
    return consistency * sigmoid_rampup(iter_num // 150, rampup)


def poly_lr(base_lr, iter_num, max_iterations):
    """code/train_ours_2D.py:387: lr after `iter_num` completed iterations."""
    return base_lr * (1.0 - iter_num / max_iterations) ** 0.9


# ----------------------------------------------------------------------------- dice / CE
def dice_loss_bcp(soft, target, mask, n_classes):
    """losses.DiceLoss_bcp(n_classes)(soft, target[N,1,*], mask[N,1,*]) (absent; call site
    code/train_ours_2D.py:197,206-207).  Frozen: one-hot target; per class
    1 - (2 sum(s t m) + 1e-10) / (sum(s s m) + sum(t t m) + 1e-10); mean over classes."""
    mask = mask.to(soft.dtype)
    loss = soft.new_zeros(())
    for c in range(n_classes):
        s = soft[:, c]
        t = (target[:, 0] == c).to(soft.dtype)
        m = mask[:, 0]
        inter = (s * t * m).sum()
        den = (s * s * m).sum() + (t * t * m).sum()
        loss = loss + (1.0 - (2.0 * inter + 1e-10) / (den + 1e-10))
    return loss / n_classes


def mix_loss(output, img_l, patch_l, mask, n_classes, l_weight=1.0, u_weight=0.5, unlab=False):
    """Transcription of mix_loss, code/train_ours_2D.py:198-216.  Returns
    (loss_image, loss_patch, (dice+ce)/2)."""
    img_l, patch_l = img_l.long(), patch_l.long()
    soft = F.softmax(output, dim=1)
    image_weight, patch_weight = (u_weight, l_weight) if unlab else (l_weight, u_weight)
    patch_mask = 1 - mask
    d1 = dice_loss_bcp(soft, img_l.unsqueeze(1), mask.unsqueeze(1), n_classes) * image_weight
    d2 = dice_loss_bcp(soft, patch_l.unsqueeze(1), patch_mask.unsqueeze(1), n_classes) * patch_weight
    ce_i = F.cross_entropy(output, img_l, reduction="none")
    ce_p = F.cross_entropy(output, patch_l, reduction="none")
    c1 = image_weight * (ce_i * mask).sum() / (mask.sum() + 1e-16)
    c2 = patch_weight * (ce_p * patch_mask).sum() / (patch_mask.sum() + 1e-16)
    return (d1 + c1) / 2.0, (d2 + c2) / 2.0, ((d1 + d2) + (c1 + c2)) / 2.0


def pseudo_label_block(pre1, pre2):
    """code/train_ours_2D.py:319-325: softmax, argmax, cross CE(none), knowledge."""
    soft1, soft2 = F.softmax(pre1, dim=1), F.softmax(pre2, dim=1)
    ps1, ps2 = torch.argmax(soft1, dim=1), torch.argmax(soft2, dim=1)
    know = F.cross_entropy(pre1, ps2, reduction="none") + F.cross_entropy(pre2, ps1, reduction="none")
    return soft1, soft2, ps1, ps2, know


# ----------------------------------------------------------------------------- largest CC
def largest_cc_labels(seg, n_classes):
    """get_ACDC_2DLargestCC, code/train_ours_2D.py:123-144: per sample, per foreground
    class keep the largest connected component (skimage.measure.label default = full
    connectivity; first-labelled component wins ties through np.argmax), sum of c * CC.
    scipy.ndimage.label with an all-ones structure is the same labelling (scan order).
    Returns float32 like the reference (`torch.Tensor(batch_list)`, :144)."""
    from scipy import ndimage
    seg_np = seg.detach().cpu().numpy()
    nd = seg_np.ndim - 1
    structure = np.ones((3,) * nd, dtype=bool)
    out = np.zeros(seg_np.shape, dtype=np.float32)
    for i in range(seg_np.shape[0]):
        for c in range(1, n_classes):
            binary = seg_np[i] == c
            labels, n = ndimage.label(binary, structure=structure)
            if n != 0:
                keep = labels == (np.argmax(np.bincount(labels.ravel())[1:]) + 1)
                out[i] += keep.astype(np.float32) * c
    return torch.from_numpy(out).to(seg.device)


def get_masks(output, n_classes, nms=1):
    """get_ACDC_masks, code/train_ours_2D.py:103-108."""
    probs = torch.argmax(F.softmax(output, dim=1), dim=1)
    return largest_cc_labels(probs, n_classes) if nms == 1 else probs


def generate_mask(shape_spatial, offsets, device="cpu"):
    """generate_mask, code/train_ours_2D.py:91-101, with the np.random.randint draws made
    explicit (`offsets`).  Zero box of side int(dim*2/3) per spatial dim.  Returns int64
    mask[*spatial] (the reference's img_mask and loss_mask hold the same values; loss_mask
    is only broadcast over the batch)."""
    mask = torch.ones(tuple(shape_spatial), dtype=torch.int64, device=device)
    sl = tuple(slice(o, o + int(s * 2 / 3)) for o, s in zip(offsets, shape_spatial))
    mask[sl] = 0
    return mask


def draw_mask_offsets(shape_spatial, rng=np.random):
    """The np.random.randint(0, dim - patch) draws of code/train_ours_2D.py:97-98."""
    return tuple(int(rng.randint(0, s - int(s * 2 / 3))) for s in shape_spatial)


# ----------------------------------------------------------------------------- VAT pieces
def l2n_sample(d):
    """d / (||d||_2 per sample over all non-batch dims + 1e-8)  (SURVEY.md App. B.3 lineage)."""
    n = d.reshape(d.shape[0], -1).norm(dim=1).reshape((-1,) + (1,) * (d.dim() - 1))
    return d / (n + 1e-8)


def l2n_channel(g):
    """channel-wise: one norm per (sample, channel), reduced over space."""
    sp = tuple(range(2, g.dim()))
    return g / (g.pow(2).sum(dim=sp, keepdim=True).sqrt() + 1e-8)


def l2n_spatial(g):
    """spatial-wise: one norm per (sample, position), reduced over channels."""
    return g / (g.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-8)


def perturbation(g, eps, mode="channel_spatial"):
    """The CHAP perturbation generator for one level: r = eps * normalise(g).
    FROZEN combination rule for 'channel_spatial': r = eps * l2n_sample(0.5*(n_c + n_s)),
    i.e. the channel-wise and the spatial-wise unit fields are averaged and the result is
    rescaled so that every sample's perturbation has L2 norm eps at every level."""
    if mode == "sample":
        return eps * l2n_sample(g)
    if mode == "channel":
        return eps * l2n_sample(l2n_channel(g))
    if mode == "spatial":
        return eps * l2n_sample(l2n_spatial(g))
    assert mode == "channel_spatial"
    return eps * l2n_sample(0.5 * (l2n_channel(g) + l2n_spatial(g)))


KL_REDUCTIONS = ("mean", "batchmean")


def kl_consistency(logits, target_soft, mask=None, reduction="mean"):
    """'kl': sum_c t (log t - log_softmax(logits)) per position (0 log 0 = 0), optionally
    multiplied by mask[N,*spatial], summed and divided by
      reduction='mean'       N * positions  -- a PER-PIXEL mean, the frozen default: the loss stays O(1) and
                             comparable to bcp_loss, which is also a per-pixel mean (code/train_ours_2D.py:208-209)
      reduction='batchmean'  N              -- F.kl_div(..., reduction='batchmean') of the classic VAT lineage;
                             for a segmentation map this is H*W (*D) times the per-pixel loss.  Round 1 froze this
                             and the CHAP step diverged (vat_loss 1.8e3 -> 5e5 within 25 iterations at 256^2,
                             1e15 / NaN in 3D) because cw * vat_loss swamped bcp_loss; kept selectable only.
    The VAT2d source is absent from the reference, so the scaling is a frozen CHOICE ("parity unpinned")."""
    assert reduction in KL_REDUCTIONS
    logp = F.log_softmax(logits, dim=1)
    kl = (torch.xlogy(target_soft, target_soft) - target_soft * logp).sum(dim=1)
    if mask is not None:
        kl = kl * mask.to(kl.dtype)
    denom = logits.shape[0] if reduction == "batchmean" else kl.numel()
    return kl.sum() / denom


def dice_consistency(logits, target_soft, mask=None):
    """'dice': soft Dice between softmax(logits) and the soft target, masked, mean over classes."""
    p = F.softmax(logits, dim=1)
    m = torch.ones_like(p[:, 0]) if mask is None else mask.to(p.dtype)
    loss = p.new_zeros(())
    for c in range(p.shape[1]):
        inter = (p[:, c] * target_soft[:, c] * m).sum()
        den = (p[:, c] * p[:, c] * m).sum() + (target_soft[:, c] * target_soft[:, c] * m).sum()
        loss = loss + (1.0 - (2.0 * inter + 1e-10) / (den + 1e-10))
    return loss / p.shape[1]


def consistency_distance(logits, target_soft, mask, losstype, reduction="mean"):
    """`reduction` only concerns 'kl' (the soft Dice is a ratio of sums, scale free)."""
    if losstype == "kl":
        return kl_consistency(logits, target_soft, mask, reduction)
    assert losstype == "dice"
    return dice_consistency(logits, target_soft, mask)


def create_mask_v1(pseudo1, pseudo2, knowledge, scale_factor=4, topk=0.1):
    """patch.create_maskV1 (absent; call site code/train_ours_2D.py:371).  FROZEN: patches of
    scale_factor^nd positions; patch score = mean(knowledge) + mean(pseudo1 != pseudo2);
    per sample keep patches whose score >= the k-th largest score, k = max(1, int(topk * P))
    (ties are all kept, so the result does not depend on a sort's tie order); nearest
    upsample back.  Returns float mask [N, *spatial]."""
    nd = knowledge.dim() - 1
    pool = F.avg_pool2d if nd == 2 else F.avg_pool3d
    dis = (pseudo1 != pseudo2).to(knowledge.dtype)
    score = pool(knowledge.unsqueeze(1), scale_factor) + pool(dis.unsqueeze(1), scale_factor)
    n = score.shape[0]
    flat = score.reshape(n, -1)
    k = max(1, int(topk * flat.shape[1]))
    kth = torch.topk(flat, k, dim=1).values[:, -1:]
    keep = (flat >= kth).to(knowledge.dtype).reshape(score.shape)
    for d in range(nd):
        keep = keep.repeat_interleave(scale_factor, dim=2 + d)
    return keep[:, 0]


class VAT:
    """losses.VAT2d(xi, epi, num_classes) (absent; ctor code/train_ours_2D.py:290, call :372).

    FROZEN feature-level, hierarchical spec:
      rows   x_u = the last soft1.shape[0] rows of x (the unlabelled half, or all rows)
      1      feats = encoder(x_u)                                    (with grad)
      2      d_l = l2n_sample(rand_like(f_l) - 0.5)  for the 5 levels (or `d_init` given)
             o1, o2 = decoder1/2([f_l.detach() + xi d_l]); dist = D(o1, soft2) + D(o2, soft1)
             g_l = d dist / d d_l                                     (decoder data-grad only)
      3      r_l = epi * normalise(g_l)   (perturbation(), channel+spatial)
      4      o1, o2 = decoder1/2([f_l + r_l]); loss = D(o1, soft2) + D(o2, soft1)
    Reduction of the 'kl' distance: the returned loss (step 4) uses `reduction` (default 'mean', a per-pixel mean;
    see kl_consistency).  The PROBE distance of step 2 always uses 'batchmean': the direction normalise(g) is
    invariant to a positive rescaling of dist except through the absolute 1e-8 in the norms, and a per-pixel mean
    would shrink |g| by H*W*D (1e6 in 3D) down to that epsilon.
    BatchNorm uses batch statistics in all VAT passes when the model is in train mode, and
    running statistics are NOT updated inside VAT (`model.bn_tracking(False)`).
    `model` must offer encoder(x), decoder1(feats), decoder2(feats), bn_tracking(flag) ctx.
    """

    def __init__(self, xi=10.0, epi=6.0, num_classes=4, mode="channel_spatial", reduction="mean"):
        assert reduction in KL_REDUCTIONS
        self.xi, self.epi, self.num_classes, self.mode, self.reduction = xi, epi, num_classes, mode, reduction

    def __call__(self, model, x, soft1, soft2, mask=None, losstype="kl", d_init=None, trace=None):
        x_u = x[x.shape[0] - soft1.shape[0]:]
        with model.bn_tracking(False):
            feats = model.encoder(x_u)
            if d_init is None:
                d_init = [torch.rand_like(f) - 0.5 for f in feats]
            d = [l2n_sample(t).detach().requires_grad_(True) for t in d_init]
            hat = [f.detach() + self.xi * di for f, di in zip(feats, d)]
            dist = consistency_distance(model.decoder1(hat), soft2, mask, losstype, "batchmean") + \
                consistency_distance(model.decoder2(hat), soft1, mask, losstype, "batchmean")
            g = torch.autograd.grad(dist, d)
            r = [perturbation(gi.detach(), self.epi, self.mode) for gi in g]
            adv = [f + ri for f, ri in zip(feats, r)]
            loss = consistency_distance(model.decoder1(adv), soft2, mask, losstype, self.reduction) + \
                consistency_distance(model.decoder2(adv), soft1, mask, losstype, self.reduction)
        if trace is not None:
            trace.update(feats=feats, d=d, g=g, r=r, dist=dist)
        return loss


# ----------------------------------------------------------------------------- optimiser
def sgd_momentum_step(params, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD semantics used at code/train_ours_2D.py:278,383 (dampening 0, no
    nesterov): g += wd p; buf = g (first step) or mom buf + g; p -= lr buf."""
    with torch.no_grad():
        for i, (p, g) in enumerate(zip(params, grads)):
            g = g + weight_decay * p
            if bufs[i] is None:
                bufs[i] = g.clone()
            else:
                bufs[i].mul_(momentum).add_(g)
            p.add_(bufs[i], alpha=-lr)


# ----------------------------------------------------------------------------- metrics
def dice_coefficient(pred, gt):
    """medpy.metric.binary.dc (third party, absent): 2|A&B| / (|A|+|B|), 0.0 when both empty."""
    pred = np.asarray(pred).astype(bool)
    gt = np.asarray(gt).astype(bool)
    inter = np.count_nonzero(pred & gt)
    size = np.count_nonzero(pred) + np.count_nonzero(gt)
    return 2.0 * inter / float(size) if size > 0 else 0.0

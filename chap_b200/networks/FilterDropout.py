"""Feature-level dropout perturbation between the encoder and the two decoders.

Same entry points and argument meaning as the reference's code/networks/FilterDropout.py
(`perform_dropout` :45-89, `scores_dropoutV2` :116-138, `drop_based_on_prob` :140-160): for every
pyramid level the unlabelled half of the batch gets two channel-masked copies that are appended
to the batch (one list per decoder).  The [N, C] masks are tiny host-side bookkeeping drawn with
torch's generator (like the reference); applying them AND building the two concatenated decoder
inputs is one fused kernel per level (`chap_feature_dropout_fwd`).
"""
import random

import torch

from .. import ops


def drop_based_on_prob(drop_probs, if_comp):
    """Bernoulli keep masks from per-(sample, channel) drop probabilities, rescaled so that the mean
    of each mask is 1 (reference :140-160)."""
    if if_comp:
        if random.randint(0, 1) == 0:
            mask1, mask2 = torch.bernoulli(1 - drop_probs), torch.bernoulli(drop_probs)
        else:
            mask1, mask2 = torch.bernoulli(drop_probs), torch.bernoulli(1 - drop_probs)
    else:
        mask1, mask2 = torch.bernoulli(1 - drop_probs), torch.bernoulli(1 - drop_probs)
    mask1 = mask1.float() * mask1.numel() / mask1.sum()
    mask2 = mask2.float() * mask2.numel() / mask2.sum()
    return mask1, mask2


def scores_dropoutV2(grad_sim, activation, if_comp, type):
    """Score-driven drop probabilities (reference :116-138): scores = grad_sim[c] * mean activation,
    standardised per sample, squashed with a Gaussian CDF ('gauss') or sigmoid(-2 z) ('sigmoid')."""
    scores = grad_sim.unsqueeze(0).expand(activation.size(0), activation.size(1)) * activation
    sigma = torch.std(scores, dim=1, keepdim=True)
    mean = torch.mean(scores, dim=1, keepdim=True)
    if type == 'gauss':
        z = (scores - mean) / (sigma * 2.0 + 1e-8)
        probs = torch.clamp(0.5 * (1 + torch.erf(z / (2.0 ** 0.5))), 0.0, 1.0)
    elif type == 'sigmoid':
        z = (scores - mean) / (sigma + 1e-8)
        probs = torch.sigmoid(-z * 2.0)
    else:
        raise ValueError(type)
    return drop_based_on_prob(probs, if_comp)


def _dropout2d_mask(n, c, device):
    """The factor nn.Dropout2d(0.5) applies, as [n, c]: F.dropout2d on a ones tensor [n, c, 1, 1] consumes the generator
    exactly like the reference's `nn.Dropout2d(0.5)(unlab_feat)` (the noise tensor is [n, c, 1, 1] whatever the spatial size)."""
    return torch.nn.functional.dropout2d(torch.ones(n, c, 1, 1, device=device), 0.5, True).reshape(n, c)


def draw_dropout_masks(x, level=None, scores=None, comp_drop=False):
    """The mask draws of perform_dropout (reference :54-80), level by level and in the reference's order of generator calls:
    returns a list with None (level not perturbed) or (m1, m2) [nu, C] per level.  Pure torch on tiny tensors."""
    out = []
    for idx, feat in enumerate(x):
        bs, dim = feat.shape[0], feat.shape[1]
        labeled_bs = bs // 2
        nu = bs - labeled_bs
        if level is None or idx not in level:
            out.append(None)
        elif scores is None:
            if comp_drop:
                binomial = torch.distributions.binomial.Binomial(probs=0.5)           # sampled on the host like the reference (:47,58)
                m1 = binomial.sample((labeled_bs, dim)).to(feat.device) * 2.0
                out.append((m1, 2.0 - m1))
            else:
                out.append((_dropout2d_mask(nu, dim, feat.device), _dropout2d_mask(nu, dim, feat.device)))
        elif torch.all(scores[idx].eq(0)):
            out.append((_dropout2d_mask(nu, dim, feat.device), _dropout2d_mask(nu, dim, feat.device)))
        else:
            activation = feat[labeled_bs:].detach().mean(dim=tuple(range(2, feat.dim())))     # adaptive_avg_pool2d(.., (1, 1))
            out.append(scores_dropoutV2(scores[idx], activation, comp_drop, 'sigmoid'))
    return out


def perform_dropout(x, level=None, scores=None, comp_drop=False, masks=None):
    """Returns (features_for_decoder1, features_for_decoder2); each level is cat(feat, perturbed
    unlabelled half) along the batch (reference :45-89).  Mask draws follow the reference branch by branch
    (draw_dropout_masks); applying them and building the two concatenated tensors is ONE fused kernel per level
    (`chap_feature_dropout_fwd`).
    masks (extension, the parity-test / CUDA-graph protocol): explicit list with, per level, None or a pair (m1, m2) of
    per-(sample, channel) factors [nu, C] used instead of drawing."""
    if masks is None:
        masks = draw_dropout_masks(x, level, scores, comp_drop)
    feature_fp1, feature_fp2 = [], []
    for feat, m in zip(x, masks):
        nu = feat.shape[0] - feat.shape[0] // 2
        m1, m2 = (None, None) if m is None else m
        p1, p2 = ops.feature_dropout(feat, m1, m2, nu)
        feature_fp1.append(p1)
        feature_fp2.append(p2)
    return feature_fp1, feature_fp2

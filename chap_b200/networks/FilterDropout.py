"""Feature-level dropout perturbation between the encoder and the two decoders.

Same entry points and argument meaning as the reference's code/networks/FilterDropout.py
(`perform_dropout` :45-89, `scores_dropoutV2` :116-138, `drop_based_on_prob` :140-160): for every
pyramid level the unlabelled half of the batch gets two channel-masked copies that are appended
to the batch (one list per decoder).  The [N, C] masks are tiny host-side bookkeeping drawn with
torch's generator (like the reference); applying them to the feature maps is the
`chap_channel_scale` kernel.
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
    """nn.Dropout2d(0.5) as an [n, c] factor."""
    return torch.empty(n, c, device=device).bernoulli_(0.5).mul_(2.0)


def perform_dropout(x, level=None, scores=None, comp_drop=False):
    """Returns (features_for_decoder1, features_for_decoder2); each level is cat(feat, perturbed
    unlabelled half) along the batch (reference :45-89)."""
    feature_fp1, feature_fp2 = [], []
    for idx, feat in enumerate(x):
        bs, dim = feat.shape[0], feat.shape[1]
        labeled_bs = bs // 2
        unlab = feat[labeled_bs:]
        nu = unlab.shape[0]
        if level is not None and idx in level:
            if scores is None:
                if comp_drop:
                    m1 = torch.empty(nu, dim, device=feat.device).bernoulli_(0.5).mul_(2.0)
                    m2 = 2.0 - m1
                else:
                    m1, m2 = _dropout2d_mask(nu, dim, feat.device), _dropout2d_mask(nu, dim, feat.device)
            elif torch.all(scores[idx].eq(0)):
                m1, m2 = _dropout2d_mask(nu, dim, feat.device), _dropout2d_mask(nu, dim, feat.device)
            else:
                activation = unlab.detach().mean(dim=tuple(range(2, unlab.dim())))
                m1, m2 = scores_dropoutV2(scores[idx], activation, comp_drop, 'sigmoid')
            p1, p2 = ops.channel_scale(unlab, m1), ops.channel_scale(unlab, m2)
        else:
            p1 = p2 = unlab
        feature_fp1.append(torch.cat((feat, p1)))
        feature_fp2.append(torch.cat((feat, p2)))
    return feature_fp1, feature_fp2
